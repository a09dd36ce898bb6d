#!/bin/bash
# 2-GPU bench only (refresh of the N=2 row after the FOF CTA-size change)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/p_bench_4.json 2> gpurun_out/p_bench_4.err
echo "bench 4 exit $?"; tail -c 300 gpurun_out/p_bench_4.json
