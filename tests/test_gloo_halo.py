"""N>1 host-side logic on CPU: torch.distributed (gloo), world_size 2 and 3, one mesh part per process.
tests/gloo_worker.py runs the reference's halo exchange (common/commu.f) over real messages from each rank's own
ilwork and compares with the oracle's in-process commu; bench.py's reference arm is checked under the driver's
torchrun launch (rank 0 prints the one JSON line, the other ranks exit 0 without work)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, script, *args, port=29541, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, *args]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


@pytest.mark.parametrize("world", [2, 3])
def test_commu_over_gloo_matches_the_oracle(world):
    r = _torchrun(world, os.path.join(ROOT, "tests", "gloo_worker.py"), port=29541 + world)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("GLOO_HALO")][-1]
    assert "in=1 out=1 consistent=1" in line and ("world=%d" % world) in line
    assert int(line.rsplit("=", 1)[1]) > 0


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    r = _torchrun(2, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                  "--warmup", "0", "--cpu-seconds", "1", port=29547, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["metric"] == "fp64_elements_assembled_per_s"
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
