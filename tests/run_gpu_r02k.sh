#!/bin/bash
# r02k: asynchronous CSR scatter (scatter warps) against consumers scattering themselves
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py tests/test_gpu_deterministic.py tests/test_gpu_at_size.py tests/test_golden_f77.py -x -q -m gpu 2>&1 | tail -3
for cfg in "2 0" "4 0" "3 0"; do
  set -- $cfg
  PHB200_WS_PROD=$1 PHB200_WS_SCAT=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side > gpurun_out/r02k_bench_p$1s$2.json 2> gpurun_out/r02k_bench_p$1s$2.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02k_bench_p$1s$2.json').read().strip().splitlines()[-1])
print("ws2 CSR producers $1 scatter warps $2: ElmGMRs %.4g (%.3f ms, kernel %.3f)   [EBE kernel %.3f ms]" % (d["sparse"]["elements_assembled_per_s"], d["sparse"]["assembly_ms"], d["sparse"]["assembly_kernel_ms"], d["roofline"]["kernel_ms"]))
PY
done
