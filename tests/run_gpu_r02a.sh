#!/bin/bash
# r02a: the driver's own sequence (pytest -m gpu -x, smoke), then ncu --set full of the ElmGMRs assembly kernel and the
# current SparseAp kernel, and the default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/ -x -q -m gpu --durations=15 2>&1 | tail -30 | tee gpurun_out/r02a_pytest_gpu.log
echo "pytest wall $(( $(date +%s) - T0 )) s" | tee -a gpurun_out/r02a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02a_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_asigmr_tet_ws<2>' -s 3 -c 1 \
    -o gpurun_out/prof_asm_csr -f python bench.py --steps 2 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/b_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sparseap -s 5 -c 1 \
    -o gpurun_out/prof_sparseap -f python bench.py --steps 2 --warmup 3 --no-cpu --no-mfg --no-incomp > gpurun_out/b_ncu_b.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 3000 gpurun_out/r02a_bench.json; tail -3 gpurun_out/r02a_bench.err
ls -la gpurun_out | tail -8
