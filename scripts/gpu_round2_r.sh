#!/bin/bash
# the round's remaining ~55 s of box time: the final library + shim (new host-side permutation, byte relocation of the
# reference's Particle in the harness build) through the C++ programs first, then as much of the parity file as fits.
mkdir -p gpurun_out
( time timeout 10 oracle/_ref/test_kdtree_shim 20000 ) > gpurun_out/r_harness.log 2>&1
echo "harness exit $?"; grep -c "CHECK FAILED" gpurun_out/r_harness.log; grep "HARNESS" gpurun_out/r_harness.log
( time timeout 22 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "shim or harness" ) > gpurun_out/r_shim_tests.log 2>&1
echo "shim tests exit $?"; tail -3 gpurun_out/r_shim_tests.log | cut -c1-300
( time timeout 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_phase.py -x -q -m gpu -p no:cacheprovider -k "not shim and not harness" ) > gpurun_out/r_parity_tests.log 2>&1
echo "parity tests exit $? (124 = cut by the time limit)"; tail -3 gpurun_out/r_parity_tests.log | cut -c1-300
