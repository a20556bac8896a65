#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r02l_red_peak.log
import sys; sys.path.insert(0, ".")
import bench
from phasta_b200 import SolverParams, make_tables
from phasta_b200.solver import PhastaGPU
part, y, ac = bench.build_part("small", 0, 1)
g = PhastaGPU(part, SolverParams(), make_tables(2, 2), device=0)
for nblk in (100000, 400000, 10244535):
    print("nblk=%d: warp-wide red.f64 %.1f G adds/s   cp.reduce.async.bulk.add.f64 (208 B per lane) %.1f G adds/s" % (nblk, g.red_peak(nblk), g.red_peak(-nblk)))
g.close()
PY
