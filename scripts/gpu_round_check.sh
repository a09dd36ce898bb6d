#!/bin/bash
# One gpurun call: new parity tests first (all failures reported), then the whole -m gpu suite, then the default bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
NEW="single_target or criterion_search or dense_search or node_mirror or filtered_knn or smoothed_velocity or cxx_shim"
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$NEW" > gpurun_out/new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/new_tests.log
timeout 600 python -m pytest tests -m gpu -q -x -k "not ($NEW)" > gpurun_out/all_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/all_tests.log
BENCH_VERBOSE=1 timeout 420 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/new_tests.log gpurun_out/all_tests.log
tail -c 600 gpurun_out/bench.json
