#!/bin/bash
# r02p / r02s (2 GPUs, strict time limits): the bench line of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/r02s_bench_n$N.json 2> gpurun_out/r02s_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02s_bench_n$N.json').read().strip().splitlines()[-1])
print("value %.4g (%.3f ms) e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"])); print(json.dumps(d["krylov"]))
s=d.get("side_workload"); print(json.dumps(s["parity"])[:900]); print(json.dumps(s["sparse"])[:400])
PY
grep -i "phb200\|error" gpurun_out/r02s_bench_n$N.err | head -5
