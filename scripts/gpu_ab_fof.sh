#!/bin/bash
# A/B of library builds on ONE box, FOF / build probe only
for rep in 1 2; do
  for v in ${VARIANTS:-a b}; do
    echo "== variant $v (rep $rep)"
    NBK_LIB_FILE=libnbk_$v.so python scripts/gpu_probe_fof_build.py ${1:-512} 2>&1 | tail -3
  done
done
