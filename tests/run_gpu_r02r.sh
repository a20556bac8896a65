#!/bin/bash
# r02r: SparseAp gathering p from a node-major copy against the [5][nshg] vector
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PHB200_SKIP_32M=1 timeout 200 python -m pytest tests/test_gpu_sparse.py tests/test_golden_f77.py tests/test_gpu_at_size.py tests/test_timestep.py tests/test_gpu_multipart.py -x -q -m gpu 2>&1 | tail -3
for nm in 1 0; do
  PHB200_AP_NODEMAJOR=$nm timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side > gpurun_out/r02r_bench_nm$nm.json 2> gpurun_out/r02r_bench_nm$nm.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02r_bench_nm$nm.json').read().strip().splitlines()[-1])
s=d["sparse"]
print("node-major p=$nm: SparseAp %.4f ms (kernel %.4f ms, %.0f GB/s = %.3f of HBM)  SolGMRs %.3f ms, %.4f ms/iteration" % (s["sparseap_ms"], s["sparseap_kernel_ms"], s["roofline_sparseap"]["achieved"], s["roofline_sparseap"]["frac"], s["solve_ms"], s["ms_per_iteration"]))
PY
done
