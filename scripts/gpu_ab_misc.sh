#!/bin/bash
# transposed-screen threshold on the final density kernel; CTA size of the fp64-heap kernel (a: 4 warps, b: 2 warps)
NBK_LIB_FILE=libnbk_a.so PROBE_REPS=3 python scripts/gpu_knn_sweep.py 512 64 "" "knn_transpose=8" "knn_transpose=16" "knn_transpose=20" 2>&1 | tail -4
for rep in 1 2; do for v in a b; do echo "== exact kernel, variant $v"; NBK_LIB_FILE=libnbk_$v.so PROBE_REPS=2 python scripts/gpu_knn_sweep.py 256 64 "knn_exact=1" 2>&1 | tail -1; done; done
