"""The peer-store halo transport (comm.cu commu_p2p: k_halo_send / k_halo_recv over arenas the GPUs map into each other
through CUDA IPC; opt-in with PHB200_P2P_HALO=1) run on the host: one PROCESS per rank executes the product's task
parsing, arena layout, task pairing and message addressing (phasta_b200/csrc/halo_task.h) and its two kernels
(halo_p2p.cuh, SIMT shim) over arenas in one shared mapping.  The ranks synchronise through the protocol's own flags
and acknowledgements only.  8 ranks is the configuration whose first GPU attempt failed (a rank addressed its peers'
arenas with its own task count / halo_cap): end ranks have one task and one plane, inner ranks two."""
import ctypes as C
import multiprocessing as mp
import os
import subprocess

import numpy as np
import pytest

from common import make_case, make_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "halo_host.cpp")
OUT = os.path.join(HERE, "host_emul", "_build", "libhalo_host.so")
SEQ = [(5, "in"), (5, "out"), (1, "in"), (12, "out"), (25, "in"), (25, "out")] * 3


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-pthread",
                           "-o", OUT, SRC])
    return OUT


def _rank(so, me, world, part, data, arenas, stride, tables, bar, q):
    try:
        L = C.CDLL(so)
        L.halo_host_setup.restype = C.c_long
        L.halo_host_setup.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.halo_host_pair.argtypes = [C.c_void_p]
        L.halo_host_commu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        il = np.ascontiguousarray(part.ilwork, dtype=np.int32)
        L.halo_host_setup(me, world, il.ctypes.data, C.addressof(arenas), stride, C.addressof(tables))
        bar.wait(60)                                   # every table is published (NCCL all-gather on the GPUs)
        ok = L.halo_host_pair(C.addressof(tables))
        bar.wait(60)
        outs = []
        if ok:
            for (n, code), v in zip(SEQ, data):
                g = np.asfortranarray(v)
                err = L.halo_host_commu(g.ctypes.data_as(C.c_void_p), part.nshg, n, 0 if code == "in" else 1)
                if err:
                    raise RuntimeError("peer wait timed out: %d" % err)
                outs.append(g)
        q.put((me, ok, outs))
    except Exception as e:  # pragma: no cover
        q.put((me, -1, repr(e)))


@pytest.mark.parametrize("world,max_seg", [(2, 0), (3, 7), (8, 0)])
def test_peer_store_halo_protocol_between_processes(lib, world, max_seg):
    case = make_case(world, 3, 2, nparts=world, bc="channel", max_seg=max_seg)
    parts = case[2]
    rng = np.random.default_rng(77)
    data = [[rng.standard_normal((p.nshg, n)) for n, _ in SEQ] for p in parts]
    # the same sequence through the oracle's in-process commu (commu.f), each exchange on fresh data
    o = make_oracle(case)
    ref = []
    for k, (n, code) in enumerate(SEQ):
        w = [np.asfortranarray(data[r][k].copy()) for r in range(world)]
        o.commu(w, n, code)
        ref.append(w)
    L = C.CDLL(lib)
    L.halo_host_arena_total.restype = C.c_long
    L.halo_host_arena_total.argtypes = [C.c_long]
    caps = []
    for p in parts:
        il, pos, nn = p.ilwork, 1, 0
        for _ in range(int(il[0])):
            nseg = int(il[pos + 3])
            nn += int(sum(il[pos + 5 + 2 * s] for s in range(nseg)))
            pos += 4 + 2 * nseg
        caps.append(25 * nn)
    assert len(set(caps)) > 1 or world == 2          # ranks differ in halo_cap (what the first 8-GPU run tripped on)
    stride = max(int(L.halo_host_arena_total(c)) for c in caps)
    ctx = mp.get_context("fork")
    arenas = ctx.RawArray("d", world * stride)       # zero-initialised shared mapping
    tables = ctx.RawArray("i", world * int(L.halo_host_table_words()))
    bar, q = ctx.Barrier(world), ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(lib, r, world, parts[r], data[r], arenas, stride, tables, bar, q))
             for r in range(world)]
    [p.start() for p in procs]
    got = {}
    try:
        for _ in range(world):
            me, ok, outs = q.get(timeout=240)
            assert ok == 1, (me, ok, outs)
            got[me] = outs
    except Exception:
        for p in procs:
            p.kill()
        raise
    [p.join(timeout=60) for p in procs]
    for p in procs:
        if p.is_alive():        # pragma: no cover
            p.kill()
    for r in range(world):
        for k in range(len(SEQ)):
            assert np.array_equal(got[r][k], ref[k][r]), (r, k, SEQ[k])
