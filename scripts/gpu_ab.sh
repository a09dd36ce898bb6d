#!/bin/bash
# A/B of density-kernel builds on ONE box: nbodylib_b200/libnbk_a.so vs libnbk_b.so (built by the caller), alternating
mkdir -p gpurun_out
for rep in 1 2; do
  for v in a b; do
    echo "== variant $v (rep $rep)"
    NBK_LIB_FILE=libnbk_$v.so PROBE_REPS=3 python scripts/gpu_knn_sweep.py ${1:-512} 64 "" 2>&1 | tail -1
  done
done
