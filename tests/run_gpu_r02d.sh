#!/bin/bash
# r02d: per-kernel durations of the two solves (ncu launch list), then the GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_solve.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side > gpurun_out/b_ncu_d.log 2>&1
echo "ncu rc=$?"
T0=$(date +%s)
PHB200_SKIP_32M=1 timeout 1500 python -m pytest tests/ -x -q -m gpu --durations=8 -s 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/r02d_pytest_gpu.log
echo "pytest wall $(( $(date +%s) - T0 )) s" | tee -a gpurun_out/r02d_pytest_gpu.log
