#!/bin/bash
# First thing next round: the device assertions that have never run on a B200 (tests/test_zz_gpu_late.py), without -x
# so that every failure shows, then the measured suite and the default bench line; 2 GPUs: the peer-store halo
# transport that is proven on the host but not re-measured (PHB200_P2P_HALO=1).
#   gpurun --timeout 900 -- 'bash tests/run_gpu_late.sh'
#   gpurun --gpus 2 --timeout 900 -- 'bash tests/run_gpu_late.sh p2p'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" = "p2p" ]; then
  for h in 0 1; do
    PHB200_P2P_HALO=$h timeout 600 python -m pytest tests/test_gpu_nccl.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_nccl_halo$h.log
    PHB200_P2P_HALO=$h timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_halo$h.json 2> gpurun_out/bench_n2_halo$h.err
  done
  exit 0
fi
timeout 900 python -m pytest tests/test_zz_gpu_late.py -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_late.log
timeout 1200 python -m pytest tests/ -x -q -m gpu --deselect tests/test_zz_gpu_late.py 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
