#!/bin/bash
# r02w: BASELINE.json configs[2] shape at single-GPU size (3.69 M tets + wedges), with its parity leg
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 90 python bench.py --workload c3_plate_mixed_4M --steps 5 --warmup 3 --no-cpu --no-mfg --no-incomp --no-side > gpurun_out/r02w_bench_c3.json 2> gpurun_out/r02w_bench_c3.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench_c3.json').read().strip().splitlines()[-1])
print("value %.4g (%.3f ms)" % (d["value"], d["ms_per_step"])); print(json.dumps(d["krylov"])); print(json.dumps(d["parity"])[:600])
PY
tail -2 gpurun_out/r02w_bench_c3.err
