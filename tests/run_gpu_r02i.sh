#!/bin/bash
# r02i: second-generation warp-specialised assembly kernel: parity suite, then A/B against the first generation; DMMA peak
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PHB200_SKIP_32M=1 timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02i_pytest_gpu.log
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/r02i_peaks.log
import sys; sys.path.insert(0, ".")
import bench
from phasta_b200 import SolverParams, make_tables
from phasta_b200.solver import PhastaGPU
part, y, ac = bench.build_part("small", 0, 1)
g = PhastaGPU(part, SolverParams(), make_tables(2, 2), device=0)
print("DFMA chain peak %.2f TFLOP/s   DMMA m8n8k4 chain peak %.2f TFLOP/s" % (g.fp64_peak(), g.dmma_peak()))
g.close()
PY
for gen in 2 1; do
  PHB200_ASM_WS=$gen timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side > gpurun_out/r02i_bench_ws$gen.json 2> gpurun_out/r02i_bench_ws$gen.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02i_bench_ws$gen.json').read().strip().splitlines()[-1])
print("ws gen $gen: value %.4g (%.3f ms) kernel %.3f ms  ElmGMRs %.4g (%.3f ms, kernel %.3f)  res-only %.3f ms" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["sparse"]["elements_assembled_per_s"], d["sparse"]["assembly_ms"], d["sparse"]["assembly_kernel_ms"], d["residual_only"]["ms"]))
print(json.dumps(d["krylov"]))
PY
done
