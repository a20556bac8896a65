"""Partitioned path on ONE GPU: every part gets its own phb200 context and a
host thread; the in-process 'local group' transport (csrc/comm.cu) stands in
for NCCL so ilwork halo exchange + distributed dot products are parity-tested
against the oracle's in-process multi-part run without needing several GPUs.
(The NCCL transport itself is exercised by tests/test_gpu_nccl.py / bench.py
--gpus N.)"""
import threading

import numpy as np
import pytest

from common import make_case, make_oracle, rel_l2

pytestmark = pytest.mark.gpu


def run_parts(case, fn):
    from phasta_b200.solver import PhastaGPU
    params, tables, parts, states = case
    n = len(parts)
    gs = [PhastaGPU(mp, params, tables, device=0) for mp in parts]
    for g in gs:
        g.local_group_join(n)
    out, errs = [None] * n, []

    def work(i):
        try:
            out[i] = fn(gs[i], *states[i])
        except Exception as e:  # pragma: no cover
            errs.append(e)

    # daemon threads: a worker stuck in the transport can fail the test but never keep the interpreter from exiting
    th = [threading.Thread(target=work, args=(i,), daemon=True) for i in range(n)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not errs, errs
    assert all(not t.is_alive() for t in th), "deadlock in local-group exchange"
    return gs, out


@pytest.mark.parametrize("nparts,max_seg", [(2, 0), (3, 11)])
def test_elmgmre_partitioned(nparts, max_seg):
    case = make_case(6, 4, 3, nparts=nparts, bc="channel", max_seg=max_seg)
    o = make_oracle(case)
    o.ElmGMRe()
    gs, out = run_parts(case, lambda g, y, ac: g.ElmGMRe(y, ac, want_qres=True))
    for op, r in zip(o.parts, out):
        assert rel_l2(r["qres"], op.qres) < 1e-10
        assert rel_l2(r["res"], op.res) < 1e-10
        assert rel_l2(r["BDiag"], op.BDiag) < 1e-10
    [g.close() for g in gs]


@pytest.mark.parametrize("nparts", [2, 4])
def test_solgmre_partitioned(nparts):
    case = make_case(8, 4, 3, nparts=nparts, bc="channel", etol=1e-7, Kspace=30)
    o = make_oracle(case)
    iKs, lG = o.SolGMRe()
    gs, out = run_parts(case, lambda g, y, ac: g.SolGMRe(y, ac))
    for g, op, (res, Dy) in zip(gs, o.parts, out):
        assert (g.iKs, g.lGMRES) == (iKs, lG)
        assert rel_l2(res, op.res) < 1e-10
        assert rel_l2(Dy, op.Dy) < 1e-8
    [g.close() for g in gs]


def test_commu_and_sumgat_partitioned():
    case = make_case(6, 3, 3, nparts=3, bc="none", periodic_z=False, max_seg=5)
    o = make_oracle(case)
    rng = np.random.default_rng(5)
    vecs = [np.asfortranarray(rng.standard_normal((mp.nshg, 5))) for mp in case[2]]
    ref = [v.copy(order="F") for v in vecs]
    o.commu(ref, 5, "in")
    tot = o.sumgat(ref, 5)
    o.commu(ref, 5, "out")
    res = {}

    def fn(g, y, ac):
        v = vecs[g.part.rank].copy(order="F")
        g.commu(v, 5, "in")
        s = g.sumgat(v, 5)
        g.commu(v, 5, "out")
        return v, s

    gs, out = run_parts(case, fn)
    for (v, s), r in zip(out, ref):
        assert rel_l2(v, r) < 1e-14
        assert abs(s - tot) < 1e-10 * abs(tot)
    [g.close() for g in gs]
