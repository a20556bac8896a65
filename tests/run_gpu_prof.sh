#!/bin/bash
# launch list + full ncu captures of the dominant kernels (usage: run_gpu_prof.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_asigmr_tet -s 2 -c 1 -o gpurun_out/prof_asm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-solve > gpurun_out/b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sparseap -s 2 -c 1 -o gpurun_out/prof_sparseap -f \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ap_ebe -s 2 -c 1 -o gpurun_out/prof_ap -f \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu4.log 2>&1
ls -la gpurun_out
