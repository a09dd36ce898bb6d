#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "density or golden or port_parity or fp32_key or halo" > gpurun_out/e_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/e_parity.log
tail -5 gpurun_out/e_parity.log
timeout 600 python scripts/gpu_knn_sweep.py 512 64 knn_transpose=0 knn_transpose=8 knn_transpose=12 knn_transpose=16 knn_transpose=20 knn_transpose=24 knn_transpose=32 > gpurun_out/e_sweep512.log 2>&1
timeout 300 python scripts/gpu_knn_sweep.py 256 32 knn_transpose=0 knn_transpose=16 knn_transpose=24 > gpurun_out/e_sweep256_k32.log 2>&1
cat gpurun_out/e_sweep512.log gpurun_out/e_sweep256_k32.log | grep -v "^$" | tail -12
