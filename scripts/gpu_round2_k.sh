#!/bin/bash
# 8-GPU call: the config-5 geometry (one 1024^3 cube, 8 slabs), bench only
mkdir -p gpurun_out
N=8
BENCH_VERBOSE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/k_bench_$N.json 2> gpurun_out/k_bench_$N.err
echo "bench $N exit $?" >> gpurun_out/k_bench_$N.err
tail -3 gpurun_out/k_bench_$N.err
python - <<PY
import json
j=json.loads(open("gpurun_out/k_bench_$N.json").read().strip().splitlines()[-1])
print("N=$N value %.1f Mpart/s step %.1f ms lib %.1f ms e2e %.1f ms" % (j["value"]/1e6, j["ms_per_step"], j["library_ms_per_step"], j["e2e"]["ms_per_step"]))
print({k:(v["value"]/1e6, v["ms_per_step"]) for k,v in j["rows"].items()})
print(j["extra"].get("sharded_rank0"))
PY
