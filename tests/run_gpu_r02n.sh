#!/bin/bash
# r02n: ncu evidence for the round-2 kernels: launch list of the default bench command, then --set full of the EBE
# assembly kernel, the CSR assembly kernel, the staged SparseAp and the EBE Ap
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/b_ncu0.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_asigmr_tet_ws2<.int.1' -s 3 -c 1 -o gpurun_out/prof_asm -f $B > gpurun_out/b_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_asigmr_tet_ws2<.int.2' -s 3 -c 1 -o gpurun_out/prof_asm_csr -f $B > gpurun_out/b_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sparseap_tma -s 5 -c 1 -o gpurun_out/prof_sparseap -f $B > gpurun_out/b_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ap_ebe_tet -s 5 -c 1 -o gpurun_out/prof_ap -f $B > gpurun_out/b_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mgs_pass -s 20 -c 1 -o gpurun_out/prof_mgs -f $B > gpurun_out/b_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
