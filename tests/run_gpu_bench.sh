#!/bin/bash
# bench + launch list + one full ncu capture of the assembly and Ap kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_asigmr_tet -s 2 -c 1 -o gpurun_out/prof_asm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-solve > gpurun_out/b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ap_ebe -s 2 -c 1 -o gpurun_out/prof_ap -f \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out
