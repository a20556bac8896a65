#!/bin/bash
# incompressible path on the GPU: parity tests (+ the phio GPU test), then the bench leg
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_incomp.py tests/test_phio.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_inc.log
bash tests/run_gpu_incbench.sh
