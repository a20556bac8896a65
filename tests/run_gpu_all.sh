#!/bin/bash
# full GPU round-trip: parity tests, bench, launch list, ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
which gfortran mpif90 mpirun >> gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | head -20 >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 2000 gpurun_out/bench_ref.json
bash tests/run_gpu_prof.sh > gpurun_out/prof.log 2>&1
ls -la gpurun_out
