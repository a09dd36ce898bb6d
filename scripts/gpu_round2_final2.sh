#!/bin/bash
# after the traversal-stack change: full -m gpu suite (512^3 reference parity skipped: it ran in final1), racecheck subset, bench
mkdir -p gpurun_out
NBK_SKIP_512_PARITY=1 timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/f2_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/f2_suite.log; tail -4 gpurun_out/f2_suite.log
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool racecheck --error-exitcode 7 --print-limit 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "golden_knn or golden_density or golden_fof or golden_criterion or duplicates or attached_halo or fof_linked" > gpurun_out/f2_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/f2_racecheck.log; tail -4 gpurun_out/f2_racecheck.log
timeout 900 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err
echo "bench exit $?"; tail -c 300 gpurun_out/f2_bench.json
