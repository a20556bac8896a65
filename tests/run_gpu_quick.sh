#!/bin/bash
# all GPU parity tests + the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print("value %.4g  e2e %.4g (%.2f ms)  roofline %.3f  ap %.1f  solve %.2f  sparse solve %.2f  incomp asm %.2f ms ap %.3f ms" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["ap"]["value"], d["solgmre"]["solve_ms"], d["sparse"]["solve_ms"], d["incomp"]["assembly_ms"], d["incomp"]["apfull_ms"]))
PY
tail -3 gpurun_out/bench.err
