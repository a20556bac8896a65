#!/bin/bash
# r02o: SparseAp with only the indices staged against blocks + indices staged
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py tests/test_golden_f77.py tests/test_gpu_at_size.py tests/test_timestep.py -x -q -m gpu 2>&1 | tail -3
for sa in 0 1; do
  PHB200_AP_STAGEA=$sa timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side > gpurun_out/r02o_bench_sa$sa.json 2> gpurun_out/r02o_bench_sa$sa.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02o_bench_sa$sa.json').read().strip().splitlines()[-1])
s=d["sparse"]
print("stage blocks=$sa: SparseAp %.4f ms (kernel %.4f ms, %.0f GB/s = %.3f of HBM)  SolGMRs %.3f ms, %.4f ms/iteration" % (s["sparseap_ms"], s["sparseap_kernel_ms"], s["roofline_sparseap"]["achieved"], s["roofline_sparseap"]["frac"], s["solve_ms"], s["ms_per_iteration"]))
PY
done
