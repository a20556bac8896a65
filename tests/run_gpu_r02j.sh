#!/bin/bash
# r02j: producer / consumer split of the second-generation kernel: 2+10, 3+9, 4+8 warps, and the first generation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py tests/test_gpu_deterministic.py tests/test_gpu_at_size.py -x -q -m gpu 2>&1 | tail -3
for np in 3 2 4; do
  PHB200_WS_PROD=$np timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-mfg --no-incomp --no-check --no-side > gpurun_out/r02j_bench_np$np.json 2> gpurun_out/r02j_bench_np$np.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02j_bench_np$np.json').read().strip().splitlines()[-1])
print("ws2 producers $np: value %.4g (%.3f ms) kernel %.3f ms  ElmGMRs %.4g (%.3f ms, kernel %.3f)" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["sparse"]["elements_assembled_per_s"], d["sparse"]["assembly_ms"], d["sparse"]["assembly_kernel_ms"]))
PY
done
