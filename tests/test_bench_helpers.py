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


def _rank_dependent_collectives(path):
    """(if-line, call-line, name) of every communicating call inside an `if` whose test depends on the rank: it mentions
    `rank` itself or a variable that is assigned (as a plain name) inside a rank-conditioned block"""
    import ast
    collective = {"dev_elmgmre", "dev_elmgmrs", "dev_solve", "dev_solve_sparse", "dev_solve_mfg", "dev_ap", "dev_sparseap",
                  "dev_au1mfg", "dev_elmmfg", "dev_inc_elmgmr", "dev_inc_apfull", "comm_init", "barrier", "all_reduce",
                  "broadcast", "SolGMRe", "SolGMRs", "SolMFG", "ElmGMRe", "ElmGMRs", "commu", "sumgat", "TimeStep",
                  "time_workload", "side_workload", "e2e_legs", "bench_mfg", "bench_incomp", "parity_small", "init_comm"}
    tree = ast.parse(open(path).read())

    def names(test):
        return {getattr(n, "id", None) or getattr(n, "attr", None) for n in ast.walk(test)
                if isinstance(n, (ast.Name, ast.Attribute))}

    tainted = {"rank"}
    for _ in range(3):                       # propagate: assigned under a tainted condition -> tainted
        for node in ast.walk(tree):
            if isinstance(node, ast.If) and names(node.test) & tainted:
                for sub in node.body + node.orelse:
                    for n in ast.walk(sub):
                        if isinstance(n, ast.Assign):
                            tainted |= {t.id for t in n.targets if isinstance(t, ast.Name)}
    bad = []
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and names(node.test) & tainted:
            for sub in node.body + node.orelse:
                for n in ast.walk(sub):
                    if isinstance(n, ast.Call):
                        f = n.func
                        name = f.attr if isinstance(f, ast.Attribute) else getattr(f, "id", None)
                        if name in collective:
                            bad.append((node.lineno, n.lineno, name))
    return bad


def test_no_collective_call_under_a_rank_condition(tmp_path):
    """bench.py is SPMD: anything that exchanges halos or reduces across ranks must be called by EVERY rank.  Round 2
    lost an 8-GPU run to `g.dev_elmgmrs(st)` under `if isinstance(par, dict)` -- and `par` is a dict on rank 0 only:
    rank 0 waited for a halo nobody sent.  Static check with one level of data flow; the version that hung must be
    flagged when the history is at hand."""
    assert not _rank_dependent_collectives(os.path.join(ROOT, "bench.py"))
    r = subprocess.run(["git", "-C", ROOT, "show", "5c46fed:bench.py"], capture_output=True, text=True)
    if r.returncode == 0 and "def side_workload" in r.stdout:
        old = tmp_path / "bench_that_hung.py"
        old.write_text(r.stdout)
        assert any(name == "dev_elmgmrs" for _, _, name in _rank_dependent_collectives(str(old)))
