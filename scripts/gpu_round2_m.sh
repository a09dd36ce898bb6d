#!/bin/bash
# 4-GPU call: the sharded tests only (incl. the C-ABI variants)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/m_sharded_tests.log 2>&1
echo "sharded tests exit $?" >> gpurun_out/m_sharded_tests.log; tail -30 gpurun_out/m_sharded_tests.log
