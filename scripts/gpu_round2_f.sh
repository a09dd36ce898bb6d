#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -x -k "density or golden or port_parity or fp32_key or halo or config1 or world1 or bucket" > gpurun_out/f_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/f_parity.log
tail -5 gpurun_out/f_parity.log
timeout 600 python scripts/gpu_knn_sweep.py 512 64 "" > gpurun_out/f_sweep512.log 2>&1
PROBE_N=16790000 timeout 300 python scripts/gpu_knn_sweep.py 256 64 "" > gpurun_out/f_sweep_odd1.log 2>&1
PROBE_N=16000000 timeout 300 python scripts/gpu_knn_sweep.py 256 64 "" > gpurun_out/f_sweep_odd2.log 2>&1
timeout 300 python scripts/gpu_knn_sweep.py 256 64 "" > gpurun_out/f_sweep256.log 2>&1
cat gpurun_out/f_sweep512.log gpurun_out/f_sweep_odd1.log gpurun_out/f_sweep_odd2.log gpurun_out/f_sweep256.log | grep -v "^$" | tail -12
