#!/bin/bash
# r02m: the driver's sequence on the current build: full GPU suite (32M test included), smoke, both bench arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests/ -x -q -m gpu --durations=6 2>&1 | tail -14 | tee gpurun_out/r02m_pytest_gpu.log
echo "pytest wall $(( $(date +%s) - T0 )) s" | tee -a gpurun_out/r02m_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02m_smoke.log
T0=$(date +%s)
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench.json').read().strip().splitlines()[-1])
print("value %.4g (%.3f ms) e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"])); print(json.dumps(d["krylov"])); print(json.dumps(d["roofline"])[:900])
print("parity ok:", d["parity"]["ok"], "side:", d["side_workload"]["workload"], d["side_workload"]["value"], d["side_workload"]["parity"]["ok"], json.dumps(d["side_workload"]["sparse"])[:400])
print(json.dumps(d["e2e"]["solgmrs"])); print(json.dumps(d["incomp"])[:300]); print(json.dumps(d["mfg"])[:300])
r=json.loads(open('gpurun_out/r02m_bench_ref.json').read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
tail -3 gpurun_out/r02m_bench.err
