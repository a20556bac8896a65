cd /root/repo
python bench.py --steps 10 --warmup 3 --no-cpu --no-solve > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print("value %.4g (%.2f ms) e2e %.4g (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
tail -2 gpurun_out/bench_e2e.err
