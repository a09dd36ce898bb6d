#!/bin/bash
# gpurun call A of round 2: parity suite on the new density kernel, scale parity at 256^3, kernel timing sweep, stats build.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
free -g >> gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/a_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/a_parity.log
tail -15 gpurun_out/a_parity.log
NBK_LIB_FILE=libnbk_stats.so timeout 300 python scripts/gpu_knn_sweep.py 256 64 "" knn_cap=96 knn_cap=160 knn_cap=192 > gpurun_out/a_stats.log 2>&1
timeout 600 python scripts/gpu_knn_sweep.py 256 64 "" knn_cap=96 knn_cap=112 knn_cap=160 knn_cap=192 knn_leaf=16 knn_leaf=64 > gpurun_out/a_sweep256.log 2>&1
timeout 600 python scripts/gpu_knn_sweep.py 512 64 "" knn_cap=160 > gpurun_out/a_sweep512.log 2>&1
timeout 300 python scripts/gpu_knn_sweep.py 256 32 "" > gpurun_out/a_sweep256_k32.log 2>&1
cat gpurun_out/a_stats.log gpurun_out/a_sweep256.log gpurun_out/a_sweep512.log gpurun_out/a_sweep256_k32.log | grep -v "^$" | tail -30
NBK_SKIP_512_PARITY=1 timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -s > gpurun_out/a_scale.log 2>&1
echo "scale exit $?" >> gpurun_out/a_scale.log
tail -8 gpurun_out/a_scale.log
