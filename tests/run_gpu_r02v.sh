#!/bin/bash
# r02v: BASELINE.json configs[3] size (16.1 M tets): incompressible ElmGMR into CSR + fLesSparseApFull, compressible legs beside it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 110 python bench.py --workload c4_incomp_16M --steps 5 --warmup 3 --no-cpu --no-mfg --no-check --no-side > gpurun_out/r02v_bench_16M.json 2> gpurun_out/r02v_bench_16M.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v_bench_16M.json').read().strip().splitlines()[-1])
print("value %.4g (%.3f ms)" % (d["value"], d["ms_per_step"])); print(json.dumps(d["krylov"])); print(json.dumps(d["incomp"])[:900]); print(d["sparse"]["genadj_s"])
PY
tail -2 gpurun_out/r02v_bench_16M.err
