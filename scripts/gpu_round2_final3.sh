#!/bin/bash
# after the FOF / ball CTA-size change: the FOF-, ball- and shim-related GPU tests, then the bench
mkdir -p gpurun_out
NBK_SKIP_512_PARITY=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_sharded.py -m gpu -q -x -k "fof or ball or golden or port_parity or duplicates or properties or cxx_shim or harness or criterion or dense or scale or world1 or demo or tphs" > gpurun_out/f3_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/f3_tests.log; tail -4 gpurun_out/f3_tests.log
timeout 900 python bench.py > gpurun_out/f3_bench.json 2> gpurun_out/f3_bench.err
echo "bench exit $?"; tail -c 200 gpurun_out/f3_bench.json
