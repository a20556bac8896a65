#!/bin/bash
# r02b (N GPUs): NCCL-transport parity (tests/nccl_worker.py under torchrun) for NCCL-only, peer dots, peer dots + peer halos;
# then the weak-scaling bench in the same three modes; then (1 GPU) ncu --set full of the ElmGMRs assembly kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
for mode in "0 0" "1 0" "1 1"; do
  set -- $mode
  PHB200_P2P=$1 PHB200_P2P_HALO=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port 2953$1 tests/nccl_worker.py 2>&1 | grep -a "NCCL_PARITY\|Error\|error\|phb200" | head -5 | sed "s/^/[dots_p2p=$1 halo_p2p=$2 N=$N] /" | tee -a gpurun_out/r02b_nccl_parity_n$N.log
done
ALSO_NCCL=1 bash tests/run_gpu_p2p.sh $N 2>&1 | tee gpurun_out/r02b_p2p_n$N.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_asigmr_tet_ws<.int.2>' -s 3 -c 1 \
    -o gpurun_out/prof_asm_csr -f python bench.py --steps 2 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/b_ncu_a.log 2>&1
ls -la gpurun_out/*.ncu-rep
