#!/bin/bash
# density-kernel change check: parity subset, counters of the stats build, then kernel time at 256^3 and 512^3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "density or golden or port_parity or ties or properties or halo or odd_k or aligned or smoothed" > gpurun_out/l_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/l_parity.log; tail -4 gpurun_out/l_parity.log
NBK_LIB_FILE=libnbk_stats.so PROBE_REPS=1 python scripts/gpu_knn_sweep.py 256 64 "" 2>&1 | tail -2
PROBE_REPS=3 python scripts/gpu_knn_sweep.py 256 64 "" "knn_transpose=0" 2>&1 | tail -3
PROBE_REPS=3 python scripts/gpu_knn_sweep.py 256 32 "" 2>&1 | tail -2
PROBE_REPS=3 python scripts/gpu_knn_sweep.py 512 64 "" 2>&1 | tail -2
