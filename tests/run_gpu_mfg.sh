#!/bin/bash
# bench with the matrix-free section + one full ncu capture of the AsIRes kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_asires -s 6 -c 1 -o gpurun_out/prof_asires -f \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu5.log 2>&1
ls -la gpurun_out | tail -5
