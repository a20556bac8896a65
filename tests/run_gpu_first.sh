#!/bin/bash
# first GPU contact: parity tests, quick timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
which gfortran mpif90 mpirun >> gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -40
