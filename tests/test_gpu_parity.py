"""GPU parity: libphb200.so (through the C-ABI) against the CPU oracle on the
same seeded inputs.  Tolerances are BASELINE.json's: <=1e-10 relative L2 for
assembled quantities, <=1e-8 for the converged step solution."""
import numpy as np
import pytest

from common import make_case, make_oracle, rel_l2

pytestmark = pytest.mark.gpu

TOL_ASM = 1e-10
TOL_SOL = 1e-8


def gpu(case, dev=0):
    from phasta_b200.solver import PhastaGPU
    params, tables, parts, states = case
    return PhastaGPU(parts[0], params, tables, device=dev)


@pytest.mark.parametrize("bc,rule,idiff", [("none", 2, 0), ("channel", 2, 1), ("mixed", 2, 1),
                                           ("channel", 1, 1), ("mixed", 1, 0)])
def test_elmgmre_parity(bc, rule, idiff):
    case = make_case(7, 5, 4, bc=bc, rule=rule, idiff=idiff, periodic_z=(bc != "none"))
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
    assert rel_l2(out["EGmass"], op.EGmass) < TOL_ASM
    # per-element check so one bad block cannot hide in the global norm
    d = np.abs(out["EGmass"] - op.EGmass).reshape(op.EGmass.shape[0], -1).max(axis=1)
    s = np.abs(op.EGmass).reshape(op.EGmass.shape[0], -1).max(axis=1)
    assert (d / s).max() < 1e-9
    g.close()


def test_residual_only_matches_lhs_call():
    case = make_case(6, 4, 4, bc="channel")
    g = gpu(case)
    y, ac = case[3][0]
    r1 = g.ElmGMRe(y, ac)["res"]
    r0 = g.ElmGMRe(y, ac, step=g.step(lhs=0, iprec=0))["res"]
    assert rel_l2(r0, r1) < 1e-13
    o = make_oracle(case, lhs=0, iprec=0)
    o.ElmGMRe()
    assert rel_l2(r0, o.parts[0].res) < TOL_ASM
    g.close()


def test_i3lu_i3pre_au1gmr_parity():
    case = make_case(6, 5, 4, bc="mixed")
    o = make_oracle(case)
    o.ElmGMRe()
    g = gpu(case)
    y, ac = case[3][0]
    out = g.ElmGMRe(y, ac)
    op = o.parts[0]
    # LU_Fact / forward / backward / product on the same block diagonal
    rng = np.random.default_rng(3)
    r = np.asfortranarray(rng.standard_normal((op.res.shape[0], 5)))
    BDg = out["BDiag"].copy(order="F")
    g.i3LU(BDg, None, "LU_Fact")
    o.i3LU(0)
    assert rel_l2(BDg, op.BDiag) < TOL_ASM
    for code, ic in (("forward", 1), ("backward", 2), ("product", 3)):
        rg, ro = r.copy(order="F"), r.copy(order="F")
        g.i3LU(None, rg, code)
        o.i3LU(ic, [ro])
        assert rel_l2(rg, ro) < 1e-12, code
    # i3pre
    EGg = g.i3pre(want_egmass=True)
    o.i3pre()
    assert rel_l2(EGg, op.EGmass) < TOL_ASM
    # Au1GMR + bc3per on a random vector that respects periodicity
    u = np.asfortranarray(rng.standard_normal((op.res.shape[0], 5)))
    ug, uo = u.copy(order="F"), u.copy(order="F")
    g.Au1GMR(ug)
    g.bc3per(ug)
    o.Au1GMR([uo])
    o.bc3per([uo])
    assert rel_l2(ug, uo) < 1e-12
    g.close()


@pytest.mark.parametrize("bc,nx", [("channel", 8), ("mixed", 6)])
def test_solgmre_parity(bc, nx):
    case = make_case(nx, 6, 5, bc=bc, etol=1e-6, Kspace=40)
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
    k = g.iKs
    assert rel_l2(g.HBrg[:k + 1, :k], o.HBrg[:k + 1, :k]) < 1e-8
    g.close()


def test_solgmre_restart_cycles():
    case = make_case(6, 5, 4, bc="channel", etol=1e-9, Kspace=8, nGMRES=4)
    o = make_oracle(case)
    iKs_o, lG_o = o.SolGMRe()
    g = gpu(case)
    y, ac = case[3][0]
    res, Dy = g.SolGMRe(y, ac)
    assert (g.iKs, g.lGMRES) == (iKs_o, lG_o)
    assert g.ntotGM == o.ntotGM.value
    assert rel_l2(Dy, o.parts[0].Dy) < 1e-7
    g.close()


def test_empty_and_ragged_blocks():
    # ibksiz that does not divide numel; element count not a multiple of the 32-element tile
    case = make_case(3, 3, 1, bc="none", periodic_z=False, ibksiz=7)
    assert case[2][0].numel % 32 != 0
    o = make_oracle(case)
    o.ElmGMRe()
    g = gpu(case)
    y, ac = case[3][0]
    out = g.ElmGMRe(y, ac, want_egmass=True)
    assert rel_l2(out["res"], o.parts[0].res) < TOL_ASM
    assert rel_l2(out["EGmass"], o.parts[0].EGmass) < TOL_ASM
    g.close()


def test_sumgat_and_fp64_peak():
    case = make_case(4, 4, 4, bc="none", periodic_z=False)
    g = gpu(case)
    u = np.asfortranarray(np.random.default_rng(0).standard_normal((case[2][0].nshg, 5)))
    assert abs(g.sumgat(u, 5) - u.sum()) < 1e-9 * np.abs(u).sum()
    assert g.fp64_peak() > 1.0
    assert g.launches() > 0
    g.close()


@pytest.mark.parametrize("natural,bc", [("none", "channel"), ("mixed", "channel"), ("mixed", "none")])
def test_boundary_elements_parity(natural, bc):
    """AsBMFG/e3b boundary flux (asbmfg.f, e3b.f, e3bvar.f) incl. /aerfrc/."""
    case = make_case(6, 5, 4, bc=bc, periodic_z=(bc != "none"), boundary=True, natural=natural)
    assert case[2][0].nelblb > 0
    o = make_oracle(case)
    o.ElmGMRe()
    g = gpu(case)
    y, ac = case[3][0]
    out = g.ElmGMRe(y, ac, want_egmass=True)
    op = o.parts[0]
    assert rel_l2(out["res"], op.res) < TOL_ASM
    assert rel_l2(out["EGmass"], op.EGmass) < TOL_ASM
    F, H, fl = g.aerfrc()
    assert rel_l2(F, op.aerfrc[:3]) < 1e-10 and abs(H - op.aerfrc[3]) <= 1e-10 * abs(op.aerfrc[3])
    assert rel_l2(fl.ravel(order="F"), op.aerfrc[4:]) < 1e-10
    # the boundary flux really contributes
    case0 = make_case(6, 5, 4, bc=bc, periodic_z=(bc != "none"))
    g0 = gpu(case0)
    r0 = g0.ElmGMRe(y, ac)["res"]
    assert rel_l2(r0, out["res"]) > 1e-3 or bc == "channel"
    g.close()
    g0.close()


def test_solgmre_with_boundary_elements():
    case = make_case(8, 6, 5, bc="channel", boundary=True, natural="mixed", etol=1e-6, Kspace=40)
    o = make_oracle(case)
    iKs_o, _ = o.SolGMRe()
    g = gpu(case)
    y, ac = case[3][0]
    res, Dy = g.SolGMRe(y, ac)
    assert g.iKs == iKs_o
    assert rel_l2(g.rmes, o.parts[0].rmes) < TOL_ASM
    assert rel_l2(Dy, o.parts[0].Dy) < TOL_SOL
    g.close()
