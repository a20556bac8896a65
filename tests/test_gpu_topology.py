"""GPU parity on hexes, wedges and mixed tet/wedge meshes (BASELINE.json
configs[2]): the generic-topology kernels (k_asiq_gen, k_asigmr_gen,
k_i3pre<NSHL>, k_ap_ebe_gen) against the CPU oracle, EBE and block-CSR
flavours, single part and partitioned.  On mixed meshes the reference sizes
EGmass with nedof = 5*max(nshl) and tet blocks use the [1:20,1:20] corner
(genint.f:249-253, asaugmr.f:32-33)."""
import numpy as np
import pytest

from common import make_case, make_oracle, rel_l2
from test_gpu_multipart import run_parts

pytestmark = pytest.mark.gpu

TOL_ASM = 1e-10
TOL_SOL = 1e-8


def gpu(case, dev=0):
    from phasta_b200.solver import PhastaGPU
    params, tables, parts, states = case
    return PhastaGPU(parts[0], params, tables, device=dev)


@pytest.mark.parametrize("topo,bc,idiff", [("hex", "channel", 1), ("wedge", "mixed", 1), ("mixed", "channel", 1),
                                           ("mixed", "mixed", 0), ("hex", "none", 0)])
def test_elmgmre_parity_topologies(topo, bc, idiff):
    case = make_case(6, 5, 4, bc=bc, idiff=idiff, periodic_z=(bc != "none"), topo=topo, ibksiz=29)
    o = make_oracle(case)
    o.ElmGMRe()
    g = gpu(case)
    y, ac = case[3][0]
    out = g.ElmGMRe(y, ac, want_egmass=True, want_qres=True)
    op = o.parts[0]
    if idiff:
        assert rel_l2(out["qres"], op.qres) < TOL_ASM
    assert rel_l2(out["res"], op.res) < TOL_ASM
    assert rel_l2(out["BDiag"], op.BDiag) < TOL_ASM
    assert out["EGmass"].shape == op.EGmass.shape
    assert rel_l2(out["EGmass"], op.EGmass) < TOL_ASM
    d = np.abs(out["EGmass"] - op.EGmass).reshape(op.EGmass.shape[0], -1).max(axis=1)
    s = np.abs(op.EGmass).reshape(op.EGmass.shape[0], -1).max(axis=1)
    assert (d / s).max() < 1e-9
    # residual-only call gives the same residual
    r0 = g.ElmGMRe(y, ac, step=g.step(lhs=0, iprec=0))["res"]
    assert rel_l2(r0, out["res"]) < 1e-13
    g.close()


@pytest.mark.parametrize("topo", ["hex", "wedge", "mixed"])
def test_i3pre_au1gmr_parity_topologies(topo):
    case = make_case(5, 5, 4, bc="mixed", topo=topo)
    o = make_oracle(case)
    o.ElmGMRe()
    g = gpu(case)
    y, ac = case[3][0]
    out = g.ElmGMRe(y, ac)
    op = o.parts[0]
    BDg = out["BDiag"].copy(order="F")
    g.i3LU(BDg, None, "LU_Fact")
    o.i3LU(0)
    assert rel_l2(BDg, op.BDiag) < TOL_ASM
    EGg = g.i3pre(want_egmass=True)
    o.i3pre()
    assert rel_l2(EGg, op.EGmass) < TOL_ASM
    rng = np.random.default_rng(3)
    u = np.asfortranarray(rng.standard_normal((op.res.shape[0], 5)))
    ug, uo = u.copy(order="F"), u.copy(order="F")
    g.Au1GMR(ug)
    g.bc3per(ug)
    o.Au1GMR([uo])
    o.bc3per([uo])
    assert rel_l2(ug, uo) < 1e-12
    g.close()


@pytest.mark.parametrize("topo,bc", [("hex", "channel"), ("wedge", "channel"), ("mixed", "mixed")])
def test_solgmre_parity_topologies(topo, bc):
    case = make_case(7, 6, 4, bc=bc, etol=1e-6, Kspace=40, topo=topo)
    o = make_oracle(case)
    iKs_o, lG_o = o.SolGMRe()
    g = gpu(case)
    y, ac = case[3][0]
    res, Dy = g.SolGMRe(y, ac)
    op = o.parts[0]
    assert g.iKs == iKs_o and g.lGMRES == lG_o
    assert rel_l2(res, op.res) < TOL_ASM
    assert rel_l2(g.rmes, op.rmes) < TOL_ASM
    assert rel_l2(Dy, op.Dy) < TOL_SOL
    g.close()


@pytest.mark.parametrize("topo", ["hex", "mixed"])
def test_sparse_flavour_topologies(topo):
    case = make_case(6, 5, 4, bc="mixed", topo=topo, etol=1e-6, Kspace=40, minIters=0)
    o = make_oracle(case)
    (ntot,) = o.genadj()
    o.ElmGMRs()
    g = gpu(case)
    colm, rowp, n = g.genadj()
    op = o.parts[0]
    assert n == ntot and np.array_equal(colm, op.colm) and np.array_equal(rowp, op.rowp)
    y, ac = case[3][0]
    out = g.ElmGMRs(y, ac, want_lhsk=True)
    assert rel_l2(out["res"], op.res) < TOL_ASM
    assert rel_l2(out["BDiag"], op.BDiag) < TOL_ASM
    assert rel_l2(out["lhsK"], op.lhsK) < TOL_ASM
    o2 = make_oracle(case)
    o2.genadj()
    iKs_o, _ = o2.SolGMRs()
    res, Dy = g.SolGMRs(y, ac)
    assert g.iKs == iKs_o
    assert rel_l2(Dy, o2.parts[0].Dy) < TOL_SOL
    g.close()


def test_solgmre_mixed_partitioned():
    case = make_case(8, 5, 3, nparts=2, bc="channel", etol=1e-7, Kspace=30, topo="mixed", max_seg=7)
    o = make_oracle(case)
    iKs, lG = o.SolGMRe()
    gs, out = run_parts(case, lambda g, y, ac: g.SolGMRe(y, ac))
    for g, op, (res, Dy) in zip(gs, o.parts, out):
        assert (g.iKs, g.lGMRES) == (iKs, lG)
        assert rel_l2(res, op.res) < TOL_ASM
        assert rel_l2(Dy, op.Dy) < TOL_SOL
    [g.close() for g in gs]
