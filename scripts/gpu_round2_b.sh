#!/bin/bash
# gpurun call B: parity of density kernel v2 (packed words), timing sweep, then an ncu full capture at 256^3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "density or golden or port_parity or fp32_key or halo or config1" > gpurun_out/b_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/b_parity.log
tail -12 gpurun_out/b_parity.log
NBK_LIB_FILE=libnbk_stats.so timeout 300 python scripts/gpu_knn_sweep.py 256 64 "" > gpurun_out/b_stats.log 2>&1
timeout 600 python scripts/gpu_knn_sweep.py 256 64 "" knn_cap=96 knn_cap=160 knn_leaf=16 > gpurun_out/b_sweep256.log 2>&1
timeout 600 python scripts/gpu_knn_sweep.py 512 64 "" > gpurun_out/b_sweep512.log 2>&1
timeout 300 python scripts/gpu_knn_sweep.py 256 32 "" > gpurun_out/b_sweep256_k32.log 2>&1
cat gpurun_out/b_stats.log gpurun_out/b_sweep256.log gpurun_out/b_sweep512.log gpurun_out/b_sweep256_k32.log | grep -v "^$" | tail -30
bash scripts/gpu_ncu_knn.sh 256 r2_ap_v2
