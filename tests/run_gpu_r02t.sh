#!/bin/bash
# r02t: GPU suite + smoke on the final, cleaned-up build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
timeout 200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02t_pytest_gpu.log
echo "pytest wall $(( $(date +%s) - T0 )) s" | tee -a gpurun_out/r02t_pytest_gpu.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02t_smoke.log
