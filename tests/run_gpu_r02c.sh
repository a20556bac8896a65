#!/bin/bash
# r02c: the whole GPU suite with the device-side Krylov loop and the staged SparseAp, parity at size, the new bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
PHB200_SKIP_32M=1 timeout 1500 python -m pytest tests/ -x -q -m gpu --durations=8 -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/r02c_pytest_gpu.log
echo "pytest wall $(( $(date +%s) - T0 )) s" | tee -a gpurun_out/r02c_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02c_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench rc=$?"; head -c 1800 gpurun_out/r02c_bench.json; echo; tail -5 gpurun_out/r02c_bench.err
