"""Host-side pieces of bench.py that need no GPU: workload sizes (BASELINE.json configs), the reader of the committed
ncu summaries the roofline block quotes, the NUMA pinning fall-back, and the reference arm's JSON line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workload_sizes_are_the_configs_of_baseline_json():
    assert bench.workload_elements("c2_channel_4M") == 128 * 64 * 82 * 6 == 4030464          # configs[1]
    assert bench.workload_elements("c5_tet_32M") == 256 * 128 * 163 * 6 == 32047104          # north_star's 32 M tets
    assert bench.workload_elements("c4_incomp_16M") == 256 * 128 * 82 * 6                    # configs[3]
    # configs[2] shape per GPU: 4 wedge layers (2 wedges per hex) at each wall, tets in between
    assert bench.workload_elements("c3_plate_mixed_4M") == 128 * 82 * (2 * 4 * 2 + (64 - 8) * 6) == 3694592
    part, y, ac = bench.build_part("small", 0, 1)
    assert part.numel == bench.workload_elements("small") and y.shape == (part.nshg, 5)
    # weak scaling: rank r of N gets an x-slab of the same size, with a halo task per neighbour
    p1 = bench.build_part("small", 1, 3)[0]
    assert p1.numel == part.numel and int(p1.ilwork[0]) == 2 and p1.numpe == 3 and p1.rank == 1


def test_roofline_quotes_the_newest_committed_ncu_summary():
    t = bench.ncu_traffic("asm")
    m = bench.ncu_metric("asm", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    assert t and m and t["source"] == m["source"] and t["source"].startswith("profiles/r02")
    # DRAM traffic of the dominant kernel is within 10 % of its algorithmic 3 310 B/element (no wasted re-reads)
    assert 0.95 < t["bytes_per_launch"] / (4030464 * bench.BYTES_PER_ELEM_LHS) < 1.10
    assert 50.0 < m["value"] < 100.0
    assert bench.ncu_metric("asm", "no_such_metric") is None and bench.ncu_traffic("no_such_kernel") is None
    assert bench.FLOP_PER_ELEM_KERNEL == 49200.0


def test_numa_pinning_never_raises():
    out = bench.pin_to_gpu_numa(0)          # no GPU here: reports why and leaves the affinity alone
    assert out["pinned"] is False


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-seconds", "0.5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fp64_elements_assembled_per_s" and d["unit"] == "elements/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "c2_channel_4M" and d["higher_is_better"] is True


def test_ours_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
