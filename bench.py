#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on B200: FP64 elements assembled / s
(ElmGMRe, lhs=1) with GMRES Ap / s and the full implicit solve beside it.

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

A "step" is one ElmGMRe pass (AsIq + qpbc + AsIGMR/e3/bc3LHS + halo + bc3Res/
bc3BDg, lhs=1, iprec=1) over the whole synthetic mesh, state resident in HBM.
At N=1 the workload is BASELINE.json configs[1]: compressible channel,
128x64x82 hexes x 6 = 4 030 464 linear tets; at N>1 each GPU
gets a slab of the same per-GPU size (weak scaling, ilwork halo over NCCL).
Before anything is timed the line checks itself: a small partitioned SolGMRe
against the oracle through the communicator that is timed next, and rank 0's
full-size part against the oracle on slabs (oracle/spot_check.py); a mismatch
ends the run with exit code 1.  The same JSON line carries Ap/s (EBE Au1GMR +
bc3per), the whole SolGMRe and SolGMRs, the roofline of the dominant kernel,
the end-to-end numbers through the C-ABI with host buffers (phb200_elmgmre and
the whole phb200_solgmrs call), the CPU baseline (oracle port, -O3
-march=native, one replica per host core) and a second BASELINE.json
configuration in its own context: c5_tet_32M (32 M tets, north_star's
single-GPU target) at N=1, c3_plate_mixed_4M per GPU (tets + wedges) at N>1.
DESIGN.md section 6 explains every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_ELEM_LHS = 52000.0     # hand count, 4-pt rule, lhs=1, incl. AsIq
FLOP_PER_ELEM_ASIQ = 2800.0     # of which the AsIq/e3q pre-pass (k_asiq_tet, not the dominant kernel)
FLOP_PER_ELEM_KERNEL = FLOP_PER_ELEM_LHS - FLOP_PER_ELEM_ASIQ   # what k_asigmr_tet_ws<1> replaces: 49.2 kflop/element
FLOP_PER_ELEM_INSTR = 51026.0   # counted while executing the reference Fortran (tests/golden/count_flops_f77.py)
BYTES_PER_ELEM_LHS = 3310.0     # EGmass 3200 + ien 16 + node data/6
BYTES_PER_ELEM_AP = 3215.0      # EBE Ap
FP64_NOMINAL_TF = 40.0          # BASELINE.json north_star

WORKLOADS = {
    # name: (nx_per_gpu, ny, nz)
    "c2_channel_4M": (128, 64, 82),
    "c1_cube_50k": (20, 20, 21),
    "small": (32, 24, 24),
    # BASELINE.json configs[2] shape at single-GPU size: flat-plate box, 4 wedge layers at each wall, tets above
    "c3_plate_mixed_4M": (128, 64, 82, "mixed", 4),
    "hex_1M": (128, 96, 82, "hex", 0),
    # north_star's single-GPU target size: 256x128x163 hexes x 6 = 32 047 104 tets (EGmass 102.6 GB in HBM)
    "c5_tet_32M": (256, 128, 163),
    # BASELINE.json configs[3] size: incompressible channel, 256x128x82 hexes x 6 = 16 121 856 tets
    "c4_incomp_16M": (256, 128, 82),
}


def workload_elements(name):
    w = WORKLOADS[name]
    nxg, ny, nz = w[:3]
    topo, wl = (w[3], w[4]) if len(w) > 3 else ("tet", 0)
    nh = nxg * ny * nz
    if topo == "hex":
        return nh
    if topo == "mixed":
        return nxg * nz * (2 * wl * 2 + (ny - 2 * wl) * 6)
    return nh * 6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def ncu_traffic(kind):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the newest committed
    `ncu --set full` summary under profiles/ (one launch on the c2 workload); None if there is none."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_prof_%s.txt" % kind)))
    if not files:
        return None
    tot, unit = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for ln in open(files[-1]):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+(\w+)\s+([0-9.,]+)", ln)
        if m:
            tot += float(m.group(3).replace(",", "")) * unit.get(m.group(2), 1.0)
    return {"bytes_per_launch": tot, "source": os.path.relpath(files[-1], ROOT)} if tot else None


def ncu_metric(kind, name):
    """one metric of the newest committed `ncu --set full` summary profiles/r*_prof_<kind>.txt, or None"""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_prof_%s.txt" % kind)))
    if not files:
        return None
    for ln in open(files[-1]):
        f = ln.split()
        if f and f[0] == name:
            try:
                return {"value": float(f[-1].replace(",", "")), "source": os.path.relpath(files[-1], ROOT)}
            except ValueError:
                return None
    return None


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region: NVML (pynvml, ~5 ms period) when importable,
    else `nvidia-smi` polling."""
    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.stop_evt = threading.Event()
        self.sm, self.smax, self.reasons = [], 0.0, set()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _nvml(self):
        nv = self.nv
        while not self.stop_evt.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in self.NAMES.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_evt.wait(0.005)

    def _smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                r = [c.strip() for c in out.split(",")]
                self.sm.append(float(r[0]))
                self.smax = max(self.smax, float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_evt.wait(0.1)

    def run(self):
        (self._nvml if self.nv else self._smi)()

    def summary(self):
        self.stop_evt.set()
        self.join(timeout=3)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax or None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nv else "nvidia-smi"}


def pin_to_gpu_numa(local_rank):
    """Bind this rank's host threads to the CPUs of its GPU's NUMA node (pinned-buffer copies of several ranks then
    do not share one socket's memory controller).  Returns what was done, for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev)
        node = int(open(path + "/numa_node").read())
        cpus = open(path + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        ids &= allowed
        if node >= 0 and ids and ids != allowed:
            os.sched_setaffinity(0, ids)
            return {"numa_node": node, "cpus": cpus, "pinned": True}
        return {"numa_node": node, "cpus": cpus, "pinned": False}
    except Exception as e:  # no sysfs entry (containers): leave the affinity alone
        return {"pinned": False, "why": repr(e)[:80]}


def parity_small(world, rank, local_rank, init_comm):
    """SolGMRe of a small channel (4*world x 5 x 4 hexes, one x-slab per rank, split ilwork segments) through the
    very transport the timed legs use, against the oracle run over all parts.  Every rank computes its own errors;
    the caller takes the max over ranks."""
    from phasta_b200 import SolverParams, make_box, make_state, make_tables, global_node_count
    from phasta_b200.solver import PhastaGPU
    from oracle import oracle_py
    nx, ny, nz = 4 * world, 5, 4
    params = SolverParams(etol=1e-7, Kspace=30)
    tables = make_tables(2, 2)
    parts = make_box(nx, ny, nz, nparts=world, bc="channel", max_seg=9)
    ng = global_node_count(nx, ny, nz)
    states = [make_state(p, ng) for p in parts]
    g = PhastaGPU(parts[rank], params, tables, device=local_rank)
    init_comm(g)
    res, Dy = g.SolGMRe(*states[rank])
    o = oracle_py.Oracle(parts, params, tables, states)
    iKs, _ = o.SolGMRe()
    op = o.parts[rank]

    def rel(a, b):
        return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
    out = [rel(res, op.res), rel(g.rmes, op.rmes), rel(Dy, op.Dy), float(abs(g.iKs - iKs)), float(iKs)]
    g.close()
    return out


def parity_at_size(g, part, params, tables, y, ac, workload, st):
    """rank 0: the part just assembled on the device against the oracle on its first / middle / last x-slab
    (oracle/spot_check.py).  g holds ElmGMRe(lhs=1) of (y, ac)."""
    from oracle.spot_check import slab_check
    ny, nz = WORKLOADS[workload][1:3]
    t0 = time.perf_counter()
    s = slab_check(g, part, params, tables, y, ac, (ny + 1) * (nz + 1), flavour="ebe", max_chunks=3)
    return {"res": s["res"], "BDiag": s["BDiag"], "EGmass": s["EGmass"], "nodes_checked": s["nodes"],
            "elements_checked": s["elements"], "last_element_checked": s["last_element_checked"], "slabs": s["slabs"],
            "seconds": time.perf_counter() - t0, "tol": 1e-10,
            "ok": bool(s["res"] < 1e-10 and s["BDiag"] < 1e-10 and s["EGmass"] < 1e-10)}


def build_part(workload, rank, world):
    from phasta_b200 import make_box, make_state, global_node_count
    w = WORKLOADS[workload]
    nxg, ny, nz = w[:3]
    topo, wl = (w[3], w[4]) if len(w) > 3 else ("tet", 1)
    nx = nxg * world
    L = (1.0 * world, 0.5, 0.64)
    parts = make_box(nx, ny, nz, L=L, nparts=world, bc="channel", periodic_z=True, ibksiz=1024,
                     only_rank=rank, topo=topo, wedge_layers=max(wl, 1))
    part = parts[0]
    y, ac = make_state(part, global_node_count(nx, ny, nz))
    return part, y, ac


def cpu_baseline(workload_name, cores, seconds_target=12.0):
    """Oracle port (-O3 -march=native), one independent replica per host core
    (the reference is flat MPI, one rank per core)."""
    from concurrent.futures import ThreadPoolExecutor
    from phasta_b200 import SolverParams, make_box, make_state, make_tables, global_node_count
    from oracle import oracle_py
    oracle_py.use_fast_build(True)
    params = SolverParams(ibksiz=1024)
    tables = make_tables(2, 2)
    nx, ny, nz = 16, 16, 16     # 24 576 tets per replica per pass
    objs = []
    for r in range(cores):
        parts = make_box(nx, ny, nz, bc="channel", periodic_z=True, ibksiz=1024, seed=1234 + r)
        st = make_state(parts[0], global_node_count(nx, ny, nz), seed=1234 + r)
        objs.append(oracle_py.Oracle(parts, params, tables, [st]))
    numel = objs[0].parts[0].mp.numel

    def run(o, reps):
        for _ in range(reps):
            o.ElmGMRe()

    run(objs[0], 1)
    t0 = time.perf_counter()
    run(objs[0], 1)
    t1 = time.perf_counter() - t0
    reps = max(1, int(seconds_target / max(t1, 1e-3)))
    with ThreadPoolExecutor(cores) as ex:
        t0 = time.perf_counter()
        list(ex.map(lambda o: run(o, reps), objs))
        dt = time.perf_counter() - t0
    asm = cores * reps * numel / dt
    # Ap on the same replicas
    for o in objs:
        o.i3LU(0)
        o.i3LU(1)
    objs[0].i3pre()
    vec = [np.asfortranarray(np.random.default_rng(0).standard_normal((objs[0].parts[0].mp.nshg, 5)))]
    t0 = time.perf_counter()
    nap = 20
    for _ in range(nap):
        objs[0].Au1GMR(vec)
    ap_elem_s_1core = nap * numel / (time.perf_counter() - t0)
    return {"value": asm, "unit": "elements/s", "cores": cores, "kind": "port",
            "sample": "%d replicas x %d passes of ElmGMRe(lhs=1) on a %dx%dx%dx6=%d-tet channel box "
                      "(oracle/ C port of the reference, gcc -O3 -march=native; no Fortran toolchain on the box)"
                      % (cores, reps, nx, ny, nz, numel),
            "ap_elements_per_s_1core": ap_elem_s_1core}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        tgt = args.cpu_seconds or max(2.0, 40.0 / max(1, args.warmup + args.steps))
        cb = cpu_baseline(args.workload, cores, seconds_target=tgt)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    numel = workload_elements(args.workload) * args.gpus
    line = {"impl": "reference", "metric": "fp64_elements_assembled_per_s", "value": v, "unit": "elements/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * numel / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "elements": numel,
                       "note": "CPU arm: bounded sample per step, ms_per_step extrapolated to the full mesh"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def bench_mfg(args, init_comm, part, params, tables, y, ac, local_rank, world, barrier, maxrank, numel_total):
    """SolMFG on the same mesh in acoustic units (the reference sizes its finite-difference interval for O(1)
    variables, see phasta_b200.mesh.nondimensional) with the state run through itrBC, as itrdrv does.  Timed:
    ElmMFG (residual + modified residual + e3bdg block diagonal), Au1MFG (one matrix-free Ap) and the whole
    solve, with SolGMRs on the same state beside it."""
    from phasta_b200 import nondimensional
    from phasta_b200.solver import PhastaGPU
    # second context on the same GPU (scaled BC values and parameters); it never allocates EGmass
    P2, _, parts2, states2 = nondimensional((params, tables, [part], [(y, ac)]))
    g2 = PhastaGPU(parts2[0], P2, tables, device=local_rank)
    init_comm(g2)
    y2, ac2 = states2[0]
    g2.set_state(y2, ac2)
    g2.itrBC()
    stm = g2.step(lhs=0, iprec=1, iter=1, istep=0)
    for _ in range(2):
        g2.dev_elmmfg(stm)
    barrier()
    g2.event(0)
    for _ in range(args.steps):
        g2.dev_elmmfg(stm)
    g2.event(1)
    barrier()
    elm_ms = maxrank(g2.elapsed_ms(0, 1)) / args.steps
    g2.dev_solve_mfg(stm)       # sets eGMRES through itrFDI
    barrier()
    s_ms, its = [], []
    for _ in range(3):
        g2.dev_elmmfg(stm)
        barrier()
        g2.event(2)
        g2.dev_solve_mfg(stm)
        g2.event(3)
        barrier()
        s_ms.append(maxrank(g2.elapsed_ms(2, 3)))
        its.append(g2.iKs)
    nap = max(20, args.steps)
    for _ in range(3):
        g2.dev_au1mfg(0)
    barrier()
    g2.event(4)
    for i in range(nap):
        g2.dev_au1mfg(i % 8)
    g2.event(5)
    barrier()
    ap_ms = maxrank(g2.elapsed_ms(4, 5)) / nap
    g2.profile(True)
    g2.profile_reset()
    for i in range(4):
        g2.dev_au1mfg(i)
    pk = g2.profile_get()
    g2.profile(False)
    # the sparse solver on the same state, same tolerance
    g2.genadj()
    sts = g2.step(lhs=1, iprec=1)
    g2.dev_elmgmrs(sts)
    g2.dev_solve_sparse(sts)
    barrier()
    g2.dev_elmgmrs(sts)
    barrier()
    g2.event(6)
    sp_its = g2.dev_solve_sparse(sts)
    g2.event(7)
    barrier()
    sp_ms = maxrank(g2.elapsed_ms(6, 7))
    out = {"elmmfg_ms": elm_ms, "elements_per_s": numel_total / (elm_ms * 1e-3),
           "au1mfg_per_s": 1e3 / ap_ms, "au1mfg_ms": ap_ms, "au1mfg_element_kernel_ms": pk["assembly"][0] / 4,
           "au1mfg_node_kernels_ms": (pk["node"][0] + pk["blas1"][0] + pk["halo"][0]) / 4,
           "solve_ms": float(np.mean(s_ms)), "gmres_iterations": int(its[-1]), "eGMRES": g2.eGMRES,
           "solgmrs_same_state": {"solve_ms": sp_ms, "gmres_iterations": int(sp_its)},
           "lhs_bytes_per_element": 0, "units": "acoustic (rho0=c0=T0=L=1), state through itrBC"}
    g2.close()
    return out


def bench_incomp(args, g, part, barrier, maxrank, numel_total, hbm):
    """Incompressible flavour (BASELINE.json configs[3]) on the same mesh and CSR structure: ElmGMR into
    lhsK(9,nnz)/lhsP(4,nnz) (AsIq + qpbc + AsIGMR/e3 + bc3LHS + fillsparseI + bc3Res) and fLesSparseApFull."""
    from phasta_b200 import IncompParams
    ip = IncompParams()
    if not getattr(g, "nnz_tot", 0):
        g.genadj()
    for _ in range(2):
        g.dev_inc_elmgmr(ip)
    barrier()
    g.event(10)
    for _ in range(args.steps):
        g.dev_inc_elmgmr(ip)
    g.event(11)
    barrier()
    asm_ms = maxrank(g.elapsed_ms(10, 11)) / args.steps
    g.profile(True)
    g.profile_reset()
    for _ in range(2):
        g.dev_inc_elmgmr(ip)
    pk = g.profile_get()
    g.profile(False)
    g.event(12)
    for _ in range(args.steps):
        g.dev_inc_elmgmr(ip, lhs=0)
    g.event(13)
    barrier()
    res_ms = maxrank(g.elapsed_ms(12, 13)) / args.steps
    g.dev_inc_elmgmr(ip)
    nap = max(20, args.steps)
    for _ in range(3):
        g.dev_inc_apfull()
    barrier()
    g.event(14)
    for _ in range(nap):
        g.dev_inc_apfull()
    g.event(15)
    barrier()
    ap_ms = maxrank(g.elapsed_ms(14, 15)) / nap
    nnz = g.nnz_tot
    # algorithmic traffic: per CSR entry kLhs 72 B + pLhs 32 B + column id 4 B; per node p, q (4 doubles each) + row ptr
    ap_bytes = nnz * 108.0 + part.nshg * 68.0
    # assembly: 16 blocks x 13 doubles written per tet + 19 doubles gathered per node (x, Y, Y,t, q) + eloc + ien
    asm_bytes = part.numel * (16 * 13 * 8 + 16 * 4 + 16) + part.nshg * (19 * 8)
    gbs = ap_bytes / (ap_ms * 1e-3) / 1e9
    return {"workload": "incompressible ElmGMR (lhs=1, idiff=1, convective form, itau=0) + fLesSparseApFull",
            "elements_assembled_per_s": numel_total / (asm_ms * 1e-3), "assembly_ms": asm_ms,
            "assembly_kernel_ms": pk["assembly"][0] / 2, "asiq_kernel_ms": pk["asiq"][0] / 2,
            "node_halo_ms": (pk["node"][0] + pk["halo"][0]) / 2,
            "residual_only_ms": res_ms, "nnz_tot_per_gpu": int(nnz),
            "assembly_GBps_algorithmic": asm_bytes / (pk["assembly"][0] / 2 * 1e-3) / 1e9,
            "apfull_per_s": 1e3 / ap_ms, "apfull_ms": ap_ms,
            "roofline_apfull": {"bound": "hbm", "kernel": "k_les_ap<15>", "unit": "GB/s", "achieved": gbs,
                                "algorithmic_bytes": ap_bytes, "peak": hbm, "frac": gbs / hbm}}


_STDOUT_FD = None


def _quiet_stdout():
    """The driver wants ONE JSON line on stdout: send everything else that writes to fd 1 (NCCL's version
    banner, library chatter) to stderr and keep the real stdout for emit()."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


class Env:
    """rank bookkeeping + the contract's barrier / max-over-ranks"""

    def __init__(self, world, rank, local_rank):
        self.world, self.rank, self.local_rank = world, rank, local_rank
        self.g = None

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.g is not None:
            self.g.sync()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(self, v):
        if self.world == 1:
            return v
        import torch
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def time_workload(env, g, part, steps, warmup, *, solve=True, sparse=True, residual_only=True):
    """The timed legs on a resident state: K ElmGMRe passes (the headline), per-kernel-class shares, residual-only
    assembly, SolGMRe + Ap/s, and the block-CSR flavour (genadj, ElmGMRs, SolGMRs, SparseAp/s)."""
    barrier, maxrank, world = env.barrier, env.maxrank, env.world
    env.g = g
    numel_total = part.numel * world
    st = g.step()
    for _ in range(warmup):
        g.dev_elmgmre(st)
    barrier()
    l0 = g.launches()
    g.profile_reset()
    g.event(0)
    for _ in range(steps):
        g.dev_elmgmre(st)
    g.event(1)
    barrier()
    asm_ms = maxrank(g.elapsed_ms(0, 1)) / steps
    out = {"asm_ms": asm_ms, "value": numel_total / (asm_ms * 1e-3), "launches": g.launches() - l0}
    # per-kernel-class share (events around every launch; separate pass)
    g.profile(True)
    g.profile_reset()
    for _ in range(2):
        g.dev_elmgmre(st)
    prof = g.profile_get()
    g.profile(False)
    out["prof"] = prof
    out["kern_ms"] = prof["assembly"][0] / max(1, prof["assembly"][1])
    out["asiq_ms"] = prof["asiq"][0] / max(1, prof["asiq"][1])
    if residual_only:
        st0 = g.step(lhs=0, iprec=0)
        for _ in range(2):
            g.dev_elmgmre(st0)
        barrier()
        g.event(2)
        for _ in range(steps):
            g.dev_elmgmre(st0)
        g.event(3)
        barrier()
        out["res_ms"] = maxrank(g.elapsed_ms(2, 3)) / steps
        g.dev_elmgmre(st)   # restore the LHS
    nap = max(20, steps)
    if solve:
        # ------------------------------------------------ full SolGMRe, then Ap/s on its basis
        g.dev_elmgmre(st)
        g.dev_solve(st)
        barrier()
        solve_ms, its = [], []
        for _ in range(3):
            g.dev_elmgmre(st)
            barrier()
            g.event(4)
            its.append(g.dev_solve(st))
            g.event(5)
            barrier()
            solve_ms.append(maxrank(g.elapsed_ms(4, 5)))
        for _ in range(3):
            g.dev_ap(0)
        barrier()
        g.event(6)
        for i in range(nap):
            g.dev_ap(i % 8)
        g.event(7)
        barrier()
        ap_ms = maxrank(g.elapsed_ms(6, 7)) / nap
        g.profile(True)
        g.profile_reset()
        for i in range(4):
            g.dev_ap(i)
        ap_k_ms = g.profile_get()["ap"][0] / 4
        g.profile(False)
        sm = float(np.mean(solve_ms))
        out["ap"] = {"value": 1e3 / ap_ms, "unit": "Ap/s", "ms_per_ap": ap_ms, "kernel_ms": ap_k_ms,
                     "elements_per_s": numel_total / (ap_ms * 1e-3)}
        out["solgmre"] = {"solve_ms": sm, "gmres_iterations": int(its[-1]), "ms_per_iteration": sm / max(1, its[-1]),
                          "implicit_solve_ms": sm + asm_ms, "krylov_its_per_s": its[-1] / (sm * 1e-3),
                          "etol": g.params.etol}
    if solve and sparse:
        # ------------------------------------------------ block-CSR flavour (SolGMRs)
        t0 = time.perf_counter()
        _, _, nnz_tot = g.genadj()
        genadj_s = time.perf_counter() - t0
        for _ in range(2):
            g.dev_elmgmrs(st)
        barrier()
        g.event(10)
        for _ in range(steps):
            g.dev_elmgmrs(st)
        g.event(11)
        barrier()
        asm_s_ms = maxrank(g.elapsed_ms(10, 11)) / steps
        g.profile(True)
        g.profile_reset()
        g.dev_elmgmrs(st)
        asm_s_k_ms = g.profile_get()["assembly"][0]
        g.profile(False)
        g.dev_solve_sparse(st)
        barrier()
        s_ms, s_its = [], []
        for _ in range(3):
            g.dev_elmgmrs(st)
            barrier()
            g.event(12)
            s_its.append(g.dev_solve_sparse(st))
            g.event(13)
            barrier()
            s_ms.append(maxrank(g.elapsed_ms(12, 13)))
        for _ in range(3):
            g.dev_sparseap(0)
        barrier()
        g.event(14)
        for i in range(nap):
            g.dev_sparseap(i % 8)
        g.event(15)
        barrier()
        sap_ms = maxrank(g.elapsed_ms(14, 15)) / nap
        g.profile(True)
        g.profile_reset()
        for i in range(4):
            g.dev_sparseap(i)
        sap_k_ms = g.profile_get()["ap"][0] / 4
        g.profile(False)
        csr_bytes = nnz_tot * 204.0 + part.nshg * 84.0
        ssm = float(np.mean(s_ms))
        out["sparse"] = {
            "nnz_tot_per_gpu": int(nnz_tot), "genadj_s": genadj_s,
            "elements_assembled_per_s": numel_total / (asm_s_ms * 1e-3), "assembly_ms": asm_s_ms,
            "assembly_kernel_ms": asm_s_k_ms,
            "sparseap_per_s": 1e3 / sap_ms, "sparseap_ms": sap_ms, "sparseap_kernel_ms": sap_k_ms,
            "solve_ms": ssm, "gmres_iterations": int(s_its[-1]), "ms_per_iteration": ssm / max(1, s_its[-1]),
            "implicit_solve_ms": ssm + asm_s_ms,
            "roofline_sparseap": {"bound": "hbm", "kernel": "k_sparseap_tma<false>", "unit": "GB/s",
                                  "achieved": csr_bytes / (sap_k_ms * 1e-3) / 1e9,
                                  "algorithmic_bytes": csr_bytes}}
    return out


def e2e_legs(env, g, part, y, ac, steps, sparse):
    """The same metric through the reference-facing C-ABI calls with HOST buffers: pinned y / ac in, res (and Dy) out,
    copies inside the timed region.  (1) phb200_elmgmre -- the headline's e2e; (2) the whole call the reference's
    itrdrv makes, phb200_solgmrs (assembly into CSR + SolGMRs), when the CSR structure is resident."""
    import ctypes as C
    import torch
    from phasta_b200.solver import _p, _chk
    numel_total = part.numel * env.world
    st = g.step()
    yp = torch.from_numpy(np.ascontiguousarray(y.T)).pin_memory()      # (5,nshg) C == (nshg,5) F
    acp = torch.from_numpy(np.ascontiguousarray(ac.T)).pin_memory()
    resp = torch.empty_like(yp).pin_memory()
    dyp = torch.empty_like(yp).pin_memory()
    yv, acv, resv, dyv = (t.numpy().T for t in (yp, acp, resp, dyp))
    for _ in range(2):
        _chk(g.L.phb200_elmgmre(g.ctx, _p(yv), _p(acv), C.byref(st), _p(resv), None, None, None), "elmgmre")
    env.barrier()
    g.event(8)
    ne2e = max(3, steps // 2)
    for _ in range(ne2e):
        _chk(g.L.phb200_elmgmre(g.ctx, _p(yv), _p(acv), C.byref(st), _p(resv), None, None, None), "elmgmre")
    g.event(9)
    env.barrier()
    e2e_ms = env.maxrank(g.elapsed_ms(8, 9)) / ne2e
    e2e = {"value": numel_total / (e2e_ms * 1e-3), "unit": "elements/s",
           "h2d_bytes_per_step": int(2 * y.nbytes), "d2h_bytes_per_step": int(y.nbytes), "ms_per_step": e2e_ms,
           "api": "phb200_elmgmre (host y, ac in; host res out; EGmass/BDiag stay in HBM)"}
    if sparse:
        iKs, lG, ntot = C.c_int(0), C.c_int(0), C.c_int(0)

        def call():
            _chk(g.L.phb200_solgmrs(g.ctx, _p(yv), _p(acv), C.byref(st), _p(resv), None, None, _p(dyv), None, None,
                                    None, None, None, C.byref(iKs), C.byref(lG), C.byref(ntot)), "solgmrs")
        call()
        env.barrier()
        g.event(8)
        for _ in range(3):
            call()
        g.event(9)
        env.barrier()
        ms = env.maxrank(g.elapsed_ms(8, 9)) / 3
        e2e["solgmrs"] = {"ms_per_call": ms, "elements_per_s": numel_total / (ms * 1e-3),
                          "gmres_iterations": int(iKs.value),
                          "h2d_bytes_per_step": int(2 * y.nbytes), "d2h_bytes_per_step": int(2 * y.nbytes),
                          "api": "phb200_solgmrs (host y, ac in; ElmGMRs + SolGMRs; host res, Dy out)"}
    return e2e


def side_workload(env, args, name, params, tables, init_comm, steps):
    """A second BASELINE.json configuration under the same driver call (its own context; the main one is closed
    first so that 103 GB of EGmass fit): assembly, Ap, both solves, and the slab parity check on rank 0."""
    from phasta_b200.solver import PhastaGPU
    t0 = time.perf_counter()
    part, y, ac = build_part(name, env.rank, env.world)
    g = PhastaGPU(part, params, tables, device=env.local_rank)
    init_comm(g)
    g.set_state(y, ac)
    setup_s = time.perf_counter() - t0
    st = g.step()
    g.dev_elmgmre(st)
    par = None
    if env.rank == 0 and not args.no_check:
        try:
            par = parity_at_size(g, part, params, tables, y, ac, name, st)
        except Exception as e:
            par = {"ok": False, "error": repr(e)[:200]}
    r = time_workload(env, g, part, steps, 3, residual_only=False)
    if "sparse" in r and not args.no_check:
        # the block-CSR flavour of the same part: lhsK blocks and CSR rows against the oracle's.  EVERY rank assembles
        # (the assembly exchanges halos); only rank 0 reads its part back and runs the oracle on slabs of it.
        g.dev_elmgmrs(st)
        if isinstance(par, dict) and "error" not in par:
            try:
                from oracle.spot_check import slab_check
                ny, nz = WORKLOADS[name][1:3]
                s2 = slab_check(g, part, params, tables, y, ac, (ny + 1) * (nz + 1), flavour="csr")
                par["csr"] = {"res": s2["res"], "BDiag": s2["BDiag"], "lhsK": s2["lhsK"],
                              "csr_rows_bit_exact": s2["csr_rows_bit_exact"], "blocks_checked": s2["blocks"]}
                par["ok"] = bool(par["ok"] and s2["csr_rows_bit_exact"]
                                 and max(s2["res"], s2["BDiag"], s2["lhsK"]) < 1e-10)
            except Exception as e:
                par["csr"] = {"error": repr(e)[:200]}
                par["ok"] = False
    if isinstance(par, dict) and "sparse" in r and "solgmre" in r:
        par["iterations_ebe_csr"] = [r["solgmre"]["gmres_iterations"], r["sparse"]["gmres_iterations"]]
        par["ok"] = bool(par["ok"] and abs(par["iterations_ebe_csr"][0] - par["iterations_ebe_csr"][1]) <= 1)
    all_tets = len(WORKLOADS[name]) == 3
    tf = part.numel * FLOP_PER_ELEM_KERNEL / (r["kern_ms"] * 1e-3) / 1e12 if all_tets else None
    out = {"workload": name, "elements": part.numel * env.world, "elements_per_gpu": part.numel,
           "nodes_per_gpu": part.nshg, "value": r["value"], "unit": "elements/s", "ms_per_step": r["asm_ms"],
           "assembly_kernel_ms": r["kern_ms"], "assembly_kernel_TFLOPs_reference_equivalent": tf,
           "ap": r.get("ap"), "solgmre": r.get("solgmre"), "sparse": r.get("sparse"), "parity": par,
           "egmass_GB_per_gpu": part.numel * (5 * max(int(b.shape[1]) for b in part.mien)) ** 2 * 8 / 1e9,
           "setup_s": setup_s}
    env.barrier()        # rank 0 has been checking its part: nobody tears its communicator down before that is over
    env.g = None
    g.close()
    return out


def main():
    _quiet_stdout()
    # a rank that leaves the SPMD sequence early (an exception on one rank only) would leave the others waiting in a
    # collective for ever: after 10 minutes every rank dumps its Python stack and exits instead
    import faulthandler
    faulthandler.dump_traceback_later(600, exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_channel_4M", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-solve", action="store_true", help="skip the Ap / SolGMRe legs (profiling runs)")
    ap.add_argument("--no-sparse", action="store_true", help="skip the block-CSR (SolGMRs) leg")
    ap.add_argument("--no-mfg", action="store_true", help="skip the matrix-free (SolMFG) leg")
    ap.add_argument("--no-incomp", action="store_true", help="skip the incompressible (ElmGMR + ApFull) leg")
    ap.add_argument("--no-check", action="store_true", help="skip the parity legs (profiling runs)")
    ap.add_argument("--no-side", action="store_true",
                    help="skip the second configuration (c5_tet_32M at N=1, c3_plate_mixed_4M at N>1)")
    ap.add_argument("--cpu-seconds", type=float, default=0.0, help="CPU sample length per step of --impl reference")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        print("bench.py: WORLD_SIZE=%d but --gpus %d" % (world, args.gpus), file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; phasta_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = pin_to_gpu_numa(local_rank) if world > 1 else {"pinned": False, "why": "single rank"}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from phasta_b200 import SolverParams, make_tables
    from phasta_b200.solver import PhastaGPU, nccl_unique_id

    env = Env(world, rank, local_rank)
    params = SolverParams(ibksiz=1024, etol=1e-3, Kspace=50)
    tables = make_tables(2, 2)

    def init_comm(gx):
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(idt, 0)
            gx.comm_init(bytes(idt.cpu().tolist()))

    # ------------------------------------------------ parity before anything is timed (exit 1 on failure)
    parity = None
    if not args.no_check:
        e = parity_small(world, rank, local_rank, init_comm)
        if world > 1:
            t = torch.tensor(e[:4], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e[:4] = t.cpu().tolist()
        parity = {"case": "SolGMRe, %dx5x4-hex channel, %d part(s), transport %s" % (
                      4 * world, world, "NCCL/NVLink (+peer-memory dots)" if world > 1 else "single part"),
                  "res": e[0], "rmes": e[1], "Dy": e[2], "gmres_iterations": int(e[4]), "iterations_equal": e[3] == 0,
                  "tol": {"res": 1e-10, "Dy": 1e-8},
                  "ok": bool(e[0] < 1e-10 and e[1] < 1e-10 and e[2] < 1e-8 and e[3] == 0)}

    part, y, ac = build_part(args.workload, rank, world)
    g = PhastaGPU(part, params, tables, device=local_rank)
    init_comm(g)
    env.g = g
    barrier, maxrank = env.barrier, env.maxrank
    numel_total = part.numel * world
    st = g.step()
    g.set_state(y, ac)
    fp64_peak = g.fp64_peak() if rank == 0 else 0.0
    if parity is not None:
        g.dev_elmgmre(st)
        if rank == 0:
            try:
                parity["at_size"] = parity_at_size(g, part, params, tables, y, ac, args.workload, st)
            except Exception as e:
                parity["at_size"] = {"ok": False, "error": repr(e)[:200]}
            parity["ok"] = bool(parity["ok"] and parity["at_size"]["ok"])
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    core = time_workload(env, g, part, args.steps, args.warmup, solve=not args.no_solve, sparse=not args.no_sparse)
    asm_ms, value, launches, prof = core["asm_ms"], core["value"], core["launches"], core["prof"]
    kern_ms, asiq_ms = core["kern_ms"], core["asiq_ms"]
    extra = {k: core[k] for k in ("ap", "solgmre", "sparse") if k in core}

    # ------------------------------------------------ incompressible flavour on the same mesh / CSR structure
    if not args.no_solve and not args.no_incomp:
        try:
            extra["incomp"] = bench_incomp(args, g, part, barrier, maxrank, numel_total, peaks()[0])
        except Exception as e:
            extra["incomp"] = {"error": repr(e)}

    # ------------------------------------------------ matrix-free flavour (SolMFG): no stored LHS at all
    if not args.no_solve and not args.no_mfg:
        try:
            extra["mfg"] = bench_mfg(args, init_comm, part, params, tables, y, ac, local_rank, world, barrier,
                                     maxrank, numel_total)
        except Exception as e:  # keep the headline line alive; say what failed
            extra["mfg"] = {"error": repr(e)}

    # ------------------------------------------------ e2e through the C-ABI with host buffers
    e2e = e2e_legs(env, g, part, y, ac, args.steps, "sparse" in extra)

    # the sampler ran through every timed leg above (assembly, Ap, solves, e2e): median SM clock under load
    clocks = sampler.summary() if rank == 0 else {}
    env.g = None
    g.close()

    # ------------------------------------------------ the other BASELINE.json configuration this driver call covers
    side = None
    if not args.no_side and not args.no_solve and args.workload == "c2_channel_4M":
        name = "c5_tet_32M" if world == 1 else "c3_plate_mixed_4M"
        fits = name != "c5_tet_32M" or torch.cuda.mem_get_info(local_rank)[1] > 150e9
        if fits:
            try:
                side = side_workload(env, args, name, params, tables, init_comm, max(3, args.steps // 2))
            except Exception as e:
                side = {"workload": name, "error": repr(e)[:300]}
        else:
            side = {"workload": name, "skipped": "needs a 180 GB GPU"}

    if rank == 0:
        hbm, src = peaks()
        elem_per_launch = part.numel
        c2 = args.workload == "c2_channel_4M"
        tr_asm, tr_ap = (ncu_traffic("asm"), ncu_traffic("ap")) if c2 else (None, None)
        all_tets = len(WORKLOADS[args.workload]) == 3
        # reference-equivalent flops of the dominant kernel alone (the AsIq pre-pass runs in k_asiq_tet)
        tf = elem_per_launch * FLOP_PER_ELEM_KERNEL / (kern_ms * 1e-3) / 1e12 if all_tets else None
        roof = {"bound": "fp64", "kernel": "k_asigmr_tet_ws2<1,4,8> (FP64 FMA pipe, not tensor cores; DESIGN 4.1f)",
                "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": (tf / fp64_peak) if (fp64_peak and all_tets) else None,
                "traffic": tr_asm and tr_asm["bytes_per_launch"], "traffic_source": tr_asm and tr_asm["source"],
                "peak_source": "FP64 DFMA-chain microbenchmark run in this process at the clocks of this run "
                               "(MEASURED_PEAKS.json has no FP64 entry); nominal %.0f TF" % FP64_NOMINAL_TF,
                "frac_of_nominal": (tf / FP64_NOMINAL_TF) if all_tets else None,
                "flop_per_element": FLOP_PER_ELEM_KERNEL,
                "flop_note": "reference-equivalent: SURVEY 8(d) 52 kflop/element minus the 2.8 k of the AsIq pre-pass; the "
                             "kernel itself executes fewer (rank-2 algebra), see fp64_pipe_active_ncu",
                "fp64_pipe_active_ncu": ncu_metric("asm", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "kernel_ms": kern_ms, "asiq_kernel_ms": asiq_ms,
                "hbm_GBps_algorithmic": elem_per_launch * BYTES_PER_ELEM_LHS / (kern_ms * 1e-3) / 1e9,
                "hbm_peak_GBps": hbm, "hbm_peak_source": src}
        if "sparse" in extra:
            rs = extra["sparse"]["roofline_sparseap"]
            rs["peak"], rs["frac"], rs["peak_source"] = hbm, rs["achieved"] / hbm, src
        if "ap" in extra:
            gbs = elem_per_launch * BYTES_PER_ELEM_AP / (extra["ap"]["kernel_ms"] * 1e-3) / 1e9
            extra["roofline_ap"] = {"bound": "hbm", "kernel": "k_ap_ebe_tet", "achieved": gbs, "peak": hbm,
                                    "unit": "GB/s", "frac": gbs / hbm,
                                    "traffic": tr_ap and tr_ap["bytes_per_launch"],
                                    "traffic_source": tr_ap and tr_ap["source"], "peak_source": src}
        cb = None
        if not args.no_cpu:
            cb = cpu_baseline(args.workload, os.cpu_count() or 1)
        # short summary first: the driver keeps the head and the tail of long lines
        krylov = {}
        if "ap" in extra:
            krylov = {"ap_per_s": extra["ap"]["value"], "solgmre_solve_ms": extra["solgmre"]["solve_ms"],
                      "solgmre_ms_per_iteration": extra["solgmre"]["ms_per_iteration"]}
        if "sparse" in extra:
            krylov.update({"solgmrs_solve_ms": extra["sparse"]["solve_ms"],
                           "solgmrs_ms_per_iteration": extra["sparse"]["ms_per_iteration"],
                           "sparseap_per_s": extra["sparse"]["sparseap_per_s"],
                           "elmgmrs_elements_per_s": extra["sparse"]["elements_assembled_per_s"]})
        line = {"metric": "fp64_elements_assembled_per_s", "value": value, "unit": "elements/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": asm_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "krylov": krylov,
                "parity": parity,
                "config": {"workload": args.workload, "elements": numel_total, "nodes_per_gpu": part.nshg,
                           "elements_per_gpu": part.numel, "quadrature": "rule 2 (4-pt tets, 6-pt wedges, 8-pt hexes)", "lhs": 1, "idiff": 1,
                           "l2": "inputs larger than L2 (EGmass %.1f GB per GPU)" % (part.numel * 3200 / 1e9),
                           "partition": "x-slabs, ilwork halo over NCCL" if world > 1 else "single part",
                           "numa": numa},
                "residual_only": {"value": numel_total / (core["res_ms"] * 1e-3), "unit": "elements/s", "ms": core["res_ms"]},
                "roofline": roof, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "kernel_class_ms": {k: v[0] / 2 for k, v in prof.items()}}
        line.update(extra)
        if side is not None:
            line["side_workload"] = side
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None:
        ok = parity["ok"] if rank == 0 else True
        if side and isinstance(side.get("parity"), dict):
            ok = ok and side["parity"].get("ok", False)
        if not ok:
            print("bench.py: PARITY FAILED: %s" % json.dumps(parity)[:600], file=sys.stderr)
            return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
