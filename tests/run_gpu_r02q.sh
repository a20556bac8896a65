#!/bin/bash
# r02q: final build: ncu --set full of the shipped SparseAp, the driver's sequence (GPU suite, smoke, bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_sparseap_tma -s 5 -c 1 -o gpurun_out/prof_sparseap_v3 -f $B > gpurun_out/b_ncu_q.log 2>&1
T0=$(date +%s)
timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02q_pytest_gpu.log
echo "pytest wall $(( $(date +%s) - T0 )) s" | tee -a gpurun_out/r02q_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02q_smoke.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print("value %.4g (%.3f ms) e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"])); print(json.dumps(d["krylov"]))
print("parity ok:", d["parity"]["ok"], "side:", d["side_workload"]["workload"], d["side_workload"]["value"], d["side_workload"]["parity"]["ok"])
print(json.dumps(d["sparse"]["roofline_sparseap"]))
PY
tail -2 gpurun_out/r02q_bench.err
