#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "density or golden or port_parity or fp32_key or halo or config1" > gpurun_out/c_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/c_parity.log
tail -12 gpurun_out/c_parity.log
NBK_LIB_FILE=libnbk_stats.so timeout 300 python scripts/gpu_knn_sweep.py 256 64 knn_mode=1 > gpurun_out/c_stats.log 2>&1
timeout 600 python scripts/gpu_knn_sweep.py 256 64 "" knn_mode=1 knn_mode=1,knn_leaf=16 knn_mode=1,knn_leaf=64 > gpurun_out/c_sweep256.log 2>&1
timeout 600 python scripts/gpu_knn_sweep.py 512 64 knn_mode=1 > gpurun_out/c_sweep512.log 2>&1
timeout 300 python scripts/gpu_knn_sweep.py 256 32 "" knn_mode=1 > gpurun_out/c_sweep256_k32.log 2>&1
cat gpurun_out/c_stats.log gpurun_out/c_sweep256.log gpurun_out/c_sweep512.log gpurun_out/c_sweep256_k32.log | grep -v "^$" | tail -30
