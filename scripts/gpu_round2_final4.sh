#!/bin/bash
# after the CTA-size change of the fp64-heap kernel: every test that reaches it
mkdir -p gpurun_out
NBK_SKIP_512_PARITY=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or port_parity or knn or filtered or single_target or smoothed or config1 or tvel or duplicates or empty or ties or odd_k or input_layouts or density or cxx_shim or harness" > gpurun_out/f4_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/f4_tests.log; tail -4 gpurun_out/f4_tests.log
