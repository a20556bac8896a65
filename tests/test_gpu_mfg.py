"""GPU parity of the matrix-free flavour (SolMFG) against the CPU oracle at
sizes beyond the interpreter-generated fixtures, on tets, hexes, wedges and a
mixed mesh, single part and partitioned (local-group transport).  Cases are in
acoustic units and the state has been through itrBC (tests/common.py
`nondimensional`, tests/golden_cases.py)."""
import numpy as np
import pytest

from common import make_case, make_oracle, nondimensional, rel_l2

pytestmark = pytest.mark.gpu


def prepared(nx, ny, nz, **kw):
    case = nondimensional(make_case(nx, ny, nz, bc="channel", etol=1e-4, **kw))
    o = make_oracle(case)
    o.itrBC()
    params, tables, parts, _ = case
    states = [(p.keep["y"].copy(order="F"), p.keep["ac"].copy(order="F")) for p in o.parts]
    return (params, tables, parts, states), o


@pytest.mark.parametrize("topo", ["tet", "hex", "wedge", "mixed"])
def test_elmmfg_itrres_au1mfg_parity(topo):
    from phasta_b200.solver import PhastaGPU
    case, o = prepared(6, 5, 4, topo=topo, ibksiz=29)
    params, tables, parts, states = case
    o.set_flags(lhs=0, iprec=1)
    o.ElmMFG()
    op = o.parts[0]
    g = PhastaGPU(parts[0], params, tables, device=0)
    y, ac = states[0]
    out = g.ElmMFG(y, ac)
    assert rel_l2(out["res"], op.res) < 1e-10
    assert rel_l2(out["rmes"], op.rmes) < 1e-10
    assert rel_l2(out["BDiag"], op.BDiag) < 1e-10
    rng = np.random.default_rng(5)
    yp = np.asfortranarray(y * (1.0 + 1e-3 * rng.standard_normal(y.shape)))
    for iab in (0, 1):
        assert rel_l2(g.ItrRes(yp, iab), o.ItrRes(yp, iab)) < 1e-10
    u = np.asfortranarray(rng.standard_normal(y.shape))
    u /= np.linalg.norm(u)
    assert rel_l2(g.Au1MFG(u, 1.0e-6), o.Au1MFG_once(u, 1.0e-6)) < 1e-6
    g.close()


@pytest.mark.parametrize("topo", ["tet", "mixed"])
def test_solmfg_parity(topo):
    from phasta_b200.solver import PhastaGPU
    case, o = prepared(6, 5, 4, topo=topo)
    params, tables, parts, states = case
    o.set_flags(lhs=0, iprec=1)
    iKs, lG, eG = o.SolMFG(eGMRES=0.0, iter=1, istep=0)
    g = PhastaGPU(parts[0], params, tables, device=0)
    y, ac = states[0]
    res, Dy = g.SolMFG(y, ac, step=g.step(lhs=0, iprec=1, iter=1, istep=0), eGMRES=0.0)
    assert abs(g.eGMRES - eG) < 1e-3 * eG
    assert abs(g.iKs - iKs) <= 1
    assert rel_l2(res, o.parts[0].res) < 1e-10
    assert rel_l2(Dy, o.parts[0].Dy) < 1e-4
    # and the matrix-free solution agrees with the EBE one (same Newton system up to the frozen coefficients)
    res2, Dy2 = g.SolGMRe(y, ac, step=g.step(lhs=1, iprec=1, etol=1e-6))
    assert rel_l2(Dy, Dy2) < 5e-3
    g.close()


def test_solmfg_partitioned():
    """two parts over the in-process transport: the halo exchanges inside Au1MFG / ItrRes and the (lag-one) device-side
    Krylov loop of the matrix-free flavour against the oracle's in-process two-part run"""
    from test_gpu_multipart import run_parts
    case, o = prepared(8, 4, 3, nparts=2)
    params, tables, parts, states = case
    o.set_flags(lhs=0, iprec=1)
    iKs, lG, eG = o.SolMFG(eGMRES=0.0, iter=1, istep=0)

    def fn(g, y, ac):
        return g.SolMFG(y, ac, step=g.step(lhs=0, iprec=1, iter=1, istep=0), eGMRES=0.0)

    gs, out = run_parts(case, fn)
    for g, op, (res, Dy) in zip(gs, o.parts, out):
        assert abs(g.eGMRES - eG) < 1e-3 * eG
        assert abs(g.iKs - iKs) <= 1
        assert rel_l2(res, op.res) < 1e-10
        assert rel_l2(Dy, op.Dy) < 1e-4
    [g.close() for g in gs]
