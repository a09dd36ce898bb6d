#!/bin/bash
# A/B of library builds on ONE box: nbodylib_b200/libnbk_<v>.so for v in $VARIANTS (built by the caller), alternating
mkdir -p gpurun_out
for rep in 1 2; do
  for v in ${VARIANTS:-a b}; do
    echo "== variant $v (rep $rep)"
    NBK_LIB_FILE=libnbk_$v.so PROBE_REPS=3 python scripts/gpu_knn_sweep.py ${1:-512} 64 "" 2>&1 | tail -1
    if [ -n "$WITH_FOF" ]; then NBK_LIB_FILE=libnbk_$v.so python scripts/gpu_probe_fof_build.py ${1:-512} 2>&1 | tail -3; fi
  done
done
