#!/bin/bash
# N-GPU run (one box): NCCL-transport parity at N ranks in both dot-product transports, then the weak-scaling bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
for p2p in 1 0; do
  PHB200_P2P=$p2p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port 2953$p2p tests/nccl_worker.py 2>&1 | grep -a "NCCL_PARITY\|Error\|error\|phb200" | head -5 | sed "s/^/[peer-memory dots=$p2p N=$N] /" | tee -a gpurun_out/r02_nccl_parity_n$N.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_scale_n$N.json').read().strip().splitlines()[-1])
print("N=$N value %.4g (%.3f ms) e2e %.4g (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"])); print(json.dumps(d["krylov"]))
print("parity", d["parity"]["ok"], d["parity"]["res"], d["parity"]["Dy"])
s=d.get("side_workload"); print("side", s["workload"], "%.4g" % s["value"], s["parity"]["ok"], json.dumps(s["sparse"])[:500]); print(json.dumps(s["solgmre"]))
print(json.dumps(d["e2e"].get("solgmrs")), d["config"]["numa"])
PY
grep -i "phb200\|error" gpurun_out/r02_scale_n$N.err | head -5
