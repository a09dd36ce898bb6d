#!/bin/bash
# last call of the round: the complete -m gpu suite on the final code (512^3 reference parity skipped: it ran in final1), smoke()
mkdir -p gpurun_out
NBK_SKIP_512_PARITY=1 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/f5_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/f5_suite.log; tail -4 gpurun_out/f5_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
