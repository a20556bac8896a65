#!/bin/bash
# weak-scaling bench under torchrun on N GPUs of one box (usage: run_gpu_scale.sh N [workload ...])
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; shift
WL=${@:-c2_channel_4M}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for w in $WL; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --workload $w > gpurun_out/scale_${w}_n$N.json 2> gpurun_out/scale_${w}_n$N.err
  echo "== $w N=$N rc=$?"; tail -c 3500 gpurun_out/scale_${w}_n$N.json; tail -3 gpurun_out/scale_${w}_n$N.err
done
