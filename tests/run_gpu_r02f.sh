#!/bin/bash
# r02f (N GPUs): NCCL-transport parity in the three transport modes with the device-side Krylov loop, then the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
for mode in "0 0" "1 0" "1 1"; do
  set -- $mode
  PHB200_P2P=$1 PHB200_P2P_HALO=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port 2953$1 tests/nccl_worker.py 2>&1 | grep -a "NCCL_PARITY\|Error\|error\|phb200" | head -5 | sed "s/^/[dots_p2p=$1 halo_p2p=$2 N=$N] /" | tee -a gpurun_out/r02f_nccl_parity_n$N.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/r02f_bench_n$N.json 2> gpurun_out/r02f_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_bench_n$N.json').read().strip().splitlines()[-1])
print("value %.4g (%.3f ms)" % (d["value"], d["ms_per_step"])); print(json.dumps(d["krylov"])); print(json.dumps(d["parity"])[:700])
print(json.dumps(d["e2e"])[:600]); print(json.dumps(d.get("side_workload"))[:1500]); print(d["config"]["numa"])
PY
grep -i "phb200\|error" gpurun_out/r02f_bench_n$N.err | head -5
