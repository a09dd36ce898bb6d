#!/bin/bash
# 2-GPU call: sharded parity tests (torch driver + C ABI), bench at N=2 through the C ABI and through the torch driver
mkdir -p gpurun_out
NG=${1:-512}
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/h_sharded_tests.log 2>&1
echo "sharded tests exit $?" >> gpurun_out/h_sharded_tests.log
tail -15 gpurun_out/h_sharded_tests.log
for DRV in native torch; do
  NBK_SHARDED_DRIVER=$DRV BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --warmup 3 --ng $NG > gpurun_out/h_bench_$DRV.json 2> gpurun_out/h_bench_$DRV.err
  echo "bench $DRV exit $?" >> gpurun_out/h_bench_$DRV.err
  tail -4 gpurun_out/h_bench_$DRV.err
  tail -c 3000 gpurun_out/h_bench_$DRV.json
done
