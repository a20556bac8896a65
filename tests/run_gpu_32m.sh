#!/bin/bash
# parity tests + default bench, then north_star's single-GPU size: 32 M tets (EGmass 102.6 GB) on one B200
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
free -g | head -2; nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 python bench.py --workload c5_tet_32M --steps 5 --warmup 3 --no-cpu --no-mfg > gpurun_out/bench_32M.json 2> gpurun_out/bench_32M.err
tail -c 4000 gpurun_out/bench_32M.json; tail -5 gpurun_out/bench_32M.err
