#!/bin/bash
# r02e: SparseAp v2 + spsi3pre v2 timing, the scatter-add microbenchmark, sparse tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py tests/test_golden_f77.py tests/test_gpu_at_size.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r02e_red_peak.log
import sys; sys.path.insert(0, ".")
import bench
from phasta_b200 import SolverParams, make_tables
from phasta_b200.solver import PhastaGPU
part, y, ac = bench.build_part("small", 0, 1)
g = PhastaGPU(part, SolverParams(), make_tables(2, 2), device=0)
for nblk in (1000, 100000, 10244535, 80786922):
    print("red_peak nblk=%d: %.1f G adds/s" % (nblk, g.red_peak(nblk)))
print("fp64 peak %.2f TF" % g.fp64_peak())
g.close()
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d["krylov"])); print(json.dumps(d["sparse"])[:900])
PY
tail -3 gpurun_out/r02e_bench.err
