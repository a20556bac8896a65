#!/bin/bash
# N-GPU checks of the NVLink peer-memory transport (dots + halos): parity, then the bench with (1 1), dots only (1 0), NCCL (0 0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_nccl.py -x -q -m gpu 2>&1 | tail -5
for mode in "1 1" "1 0" ${ALSO_NCCL:+"0 0"}; do
  set -- $mode
  tag="d$1h$2"
  PHB200_P2P=$1 PHB200_P2P_HALO=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$1 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/p2p_${tag}_n$N.json 2> gpurun_out/p2p_${tag}_n$N.err
  echo "== dots_p2p=$1 halo_p2p=$2 N=$N rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/p2p_${tag}_n$N.json').read().strip().splitlines()[-1])
print("value %.4g (%.3f ms) ap %.1f/s  solgmre %.2f ms (%d its)  sparse solve %.2f ms (%d its) sparseap %.3f ms halo_ms %.3f" % (d["value"], d["ms_per_step"], d["ap"]["value"], d["solgmre"]["solve_ms"], d["solgmre"]["gmres_iterations"], d["sparse"]["solve_ms"], d["sparse"]["gmres_iterations"], d["sparse"]["sparseap_ms"], d["kernel_class_ms"]["halo"]))
PY
  grep -i "phb200\|error" gpurun_out/p2p_${tag}_n$N.err | head -5
done
