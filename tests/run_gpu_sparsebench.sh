#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_topology.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/bench_sp.json 2> gpurun_out/bench_sp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_sp.json').read().strip().splitlines()[-1])
print("value %.4g  sparse assembly %.2f ms  solve %.2f ms  sparseap %.3f" % (d["value"], d["sparse"]["assembly_ms"], d["sparse"]["solve_ms"], d["sparse"]["sparseap_ms"]))
PY
tail -2 gpurun_out/bench_sp.err
