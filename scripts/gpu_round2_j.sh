#!/bin/bash
# 4-GPU call: warp-aligned tree tests, sharded parity (world 2 and 4, torch driver + C ABI), bench at N = 2 and 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_aligned" > gpurun_out/j_aligned_tests.log 2>&1
echo "aligned tests exit $?" >> gpurun_out/j_aligned_tests.log; tail -5 gpurun_out/j_aligned_tests.log
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/j_sharded_tests.log 2>&1
echo "sharded tests exit $?" >> gpurun_out/j_sharded_tests.log; tail -5 gpurun_out/j_sharded_tests.log
for N in ${NLIST:-2 4}; do
  BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/j_bench_$N.json 2> gpurun_out/j_bench_$N.err
  echo "bench $N exit $?" >> gpurun_out/j_bench_$N.err
  tail -3 gpurun_out/j_bench_$N.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/j_bench_$N.json").read().strip().splitlines()[-1])
print("N=$N value %.1f Mpart/s step %.1f ms lib %.1f ms e2e %.1f ms" % (j["value"]/1e6, j["ms_per_step"], j["library_ms_per_step"], j["e2e"]["ms_per_step"]))
print({k:(v["value"]/1e6, v["ms_per_step"]) for k,v in j["rows"].items()})
print(j["extra"].get("sharded_rank0"))
PY
done
