#!/bin/bash
# full GPU round-trip (r01g): all parity tests, smoke, bench (both arms), configs[3] size, launch list, ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 600 gpurun_out/bench_ref.json
timeout 900 python bench.py --workload c4_incomp_16M --steps 5 --warmup 3 --no-cpu --no-mfg > gpurun_out/bench_16M.json 2> gpurun_out/bench_16M.err
tail -c 1800 gpurun_out/bench_16M.json; tail -3 gpurun_out/bench_16M.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-mfg > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inc_asigmr_tet -s 2 -c 1 -o gpurun_out/prof_incasm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-mfg > gpurun_out/b_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_les_ap -s 2 -c 1 -o gpurun_out/prof_lesap -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-mfg > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out | tail -12
