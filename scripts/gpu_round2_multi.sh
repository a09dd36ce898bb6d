#!/bin/bash
# multi-GPU call: sharded parity tests (world 2 and 4), then the bench at N = 2 and N = 4 (ng from $1, default 512)
mkdir -p gpurun_out
NG=${1:-512}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/m_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/m_sharded_tests.log 2>&1
echo "sharded tests exit $?" >> gpurun_out/m_sharded_tests.log
tail -12 gpurun_out/m_sharded_tests.log
for N in ${NLIST:-2 4}; do
  BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 4 --warmup 3 --ng $NG > gpurun_out/m_bench_$N.json 2> gpurun_out/m_bench_$N.err
  echo "bench $N exit $?" >> gpurun_out/m_bench_$N.err
  tail -4 gpurun_out/m_bench_$N.err
  tail -c 2500 gpurun_out/m_bench_$N.json
done
