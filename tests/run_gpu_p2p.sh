#!/bin/bash
# N-GPU checks of the NVLink peer-memory all-reduce: parity over NCCL+P2P, then the bench with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_nccl.py -x -q -m gpu 2>&1 | tail -5
for p2p in 1 0; do
  PHB200_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$p2p \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/p2p${p2p}_n$N.json 2> gpurun_out/p2p${p2p}_n$N.err
  echo "== P2P=$p2p N=$N rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/p2p${p2p}_n$N.json').read().strip().splitlines()[-1])
print("value %.4g  ap %.1f/s  solgmre %.2f ms (%d its)  sparse solve %.2f ms (%d its) sparseap %.3f ms" % (d["value"], d["ap"]["value"], d["solgmre"]["solve_ms"], d["solgmre"]["gmres_iterations"], d["sparse"]["solve_ms"], d["sparse"]["gmres_iterations"], d["sparse"]["sparseap_ms"]))
PY
  tail -2 gpurun_out/p2p${p2p}_n$N.err
done
