#!/bin/bash
# whole -m gpu suite INCLUDING the 512^3 parity case, then the default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x -s > gpurun_out/g_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/g_suite.log
grep -E "passed|failed|parity at|HARNESS|ratio|exit" gpurun_out/g_suite.log | tail -12
BENCH_VERBOSE=1 timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench exit $?" >> gpurun_out/g_bench.err
tail -3 gpurun_out/g_bench.err
