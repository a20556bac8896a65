#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02h_bench_n$N.json').read().strip().splitlines()[-1])
print("value %.4g (%.3f ms)" % (d["value"], d["ms_per_step"])); print(json.dumps(d["krylov"]))
s=d.get("side_workload"); print(json.dumps(s["sparse"])[:700]); print(json.dumps(s["solgmre"]))
print(json.dumps(d["sparse"])[:300])
PY
grep -i "phb200\|error" gpurun_out/r02h_bench_n$N.err | head -5
