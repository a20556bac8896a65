#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-mfg > gpurun_out/bench_inc.json 2> gpurun_out/bench_inc.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_inc.json').read().strip().splitlines()[-1])
print(json.dumps(d.get("incomp"), indent=1)); print("asm value", d["value"], "sparse", d["sparse"]["assembly_ms"])
PY
tail -3 gpurun_out/bench_inc.err
