#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse.py -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-solve > gpurun_out/bench_clk.json 2> gpurun_out/bench_clk.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_clk.json').read().strip().splitlines()[-1])
print(d["clocks"], d["value"], d["e2e"]["value"])
PY
tail -2 gpurun_out/bench_clk.err
