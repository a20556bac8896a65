#!/bin/bash
# r02x: ncu launch list (gpu__time_duration) of the final build's bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side > gpurun_out/b_ncu_x.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_final.csv
