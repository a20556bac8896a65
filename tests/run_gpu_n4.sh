#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m pytest tests/test_gpu_nccl.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_n$N.json').read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value %.4g e2e %.4g ap %.1f solgmre %.2f (%d) sparse %.2f (%d) incomp %.2f mfg %.2f clocks %s cpu %s" % (d["value"], d["e2e"]["value"], d["ap"]["value"], d["solgmre"]["solve_ms"], d["solgmre"]["gmres_iterations"], d["sparse"]["solve_ms"], d["sparse"]["gmres_iterations"], d["incomp"]["assembly_ms"], d["mfg"]["solve_ms"], d["clocks"], d["cpu_baseline"] and d["cpu_baseline"]["value"]))
PY
tail -3 gpurun_out/scale_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/scale_ref_n$N.json 2> gpurun_out/scale_ref_n$N.err
echo "ref rc=$?"; wc -l gpurun_out/scale_ref_n$N.json; tail -c 300 gpurun_out/scale_ref_n$N.json
