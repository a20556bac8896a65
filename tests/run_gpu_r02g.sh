#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python tests/diag_c3.py 2>&1 | grep -v "^$" | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tests/diag_c3.py 2>&1 | grep "rep\|rank\|unprof\|rror" | tail -12
timeout 600 python -m pytest tests/test_gpu_sparse.py tests/test_golden_f77.py tests/test_phio.py tests/test_gpu_at_size.py -x -q -m gpu 2>&1 | tail -3
