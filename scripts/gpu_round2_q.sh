#!/bin/bash
# last GPU call of round 2 (80 s of box time left): the phase-space kNN tests, the reference harness on all five tree types,
# a timing probe.  Each step has its own timeout; logs under gpurun_out/.
mkdir -p gpurun_out
( time timeout 40 python -m pytest tests/test_gpu_z_phase.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/q_phase_tests.log 2>&1
echo "phase tests exit $?"; tail -4 gpurun_out/q_phase_tests.log
( time timeout 20 oracle/_ref/test_kdtree_shim 20000 ) > gpurun_out/q_harness.log 2>&1
echo "harness exit $?"; grep -c "CHECK FAILED" gpurun_out/q_harness.log; tail -2 gpurun_out/q_harness.log | cut -c1-200
timeout 12 python scripts/gpu_phase_probe.py 1000000 32 > gpurun_out/q_phase_probe.log 2>&1
echo "probe exit $?"; cat gpurun_out/q_phase_probe.log | cut -c1-200
