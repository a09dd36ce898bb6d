#!/bin/bash
# full -m gpu suite (512^3 scale parity skipped here), then timing of the density kernel, then the bench line
mkdir -p gpurun_out
NBK_SKIP_512_PARITY=1 timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/d_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/d_suite.log
tail -15 gpurun_out/d_suite.log
NBK_LIB_FILE=libnbk_stats.so timeout 300 python scripts/gpu_knn_sweep.py 256 64 "" > gpurun_out/d_stats.log 2>&1
timeout 600 python scripts/gpu_knn_sweep.py 512 64 "" > gpurun_out/d_sweep512.log 2>&1
timeout 300 python scripts/gpu_knn_sweep.py 256 32 "" > gpurun_out/d_sweep256_k32.log 2>&1
cat gpurun_out/d_stats.log gpurun_out/d_sweep512.log gpurun_out/d_sweep256_k32.log | grep -v "^$" | tail
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench exit $?" >> gpurun_out/d_bench.err
tail -5 gpurun_out/d_bench.err
tail -c 3000 gpurun_out/d_bench.json
