"""Newton / time-step shell (itrPC.f, itrbc.f, rstat.f, itrdrv.f step loop).

CPU: the oracle's step converges quadratically-ish on a smooth channel state
(an end-to-end consistency check of residual, tangent, BCs and update signs:
a wrong sign or a tangent that is not d(res)/dY diverges immediately).
GPU: phb200_timestep against the oracle, single part and partitioned, EBE and
block-CSR flavours, LHSupd > 1."""
import numpy as np
import pytest

from common import rel_l2
from phasta_b200 import SolverParams, make_box, make_smooth_state, make_tables


def smooth_case(nx=8, ny=6, nz=4, nparts=1, **pkw):
    parts = make_box(nx, ny, nz, bc="channel", boundary=True, nparts=nparts)
    states = [make_smooth_state(p) for p in parts]
    params = SolverParams(**pkw)
    return params, make_tables(2, 2), parts, states


def test_oracle_newton_converges_and_steps_are_stable():
    from oracle.oracle_py import Oracle
    params, tables, parts, states = smooth_case(etol=1e-4)
    o = Oracle(parts, params, tables, states)
    for _ in range(3):
        st = o.TimeStep(nitr=3)
        r = st[:, 1]                                   # un-preconditioned residual norm per Newton iteration
        assert r[1] < 5e-3 * r[0] and r[2] < 5e-2 * r[1], r
        assert (st[:, 2] < params.Kspace).all()        # every linear solve converged inside one cycle
    y = o.parts[0].keep["y"]
    assert np.isfinite(y).all() and 290 < y[:, 4].min() and y[:, 4].max() < 310


def test_oracle_itrbc_enforces_the_essential_bcs():
    from oracle.oracle_py import Oracle
    params, tables, parts, states = smooth_case()
    mp = parts[0]
    y0 = states[0][0]
    o = Oracle(parts, params, tables, [(y0 + 1.0, states[0][1] + 2.0)])
    o.itrBC()
    y, ac = o.parts[0].keep["y"], o.parts[0].keep["ac"]
    wall = ((mp.iBC >> 3) & 7) == 7
    master = (mp.iper - 1) == np.arange(mp.nshg)
    assert np.array_equal(y[wall & master, 0:3], mp.BC[wall & master, 2:5])
    tset = (mp.iBC & 2) != 0
    assert np.array_equal(y[tset & master, 4], mp.BC[tset & master, 1])
    pset = (mp.iBC & 4) != 0
    assert np.array_equal(y[pset & master, 3], mp.BC[pset & master, 0])
    assert np.array_equal(y, y[mp.iper - 1]) and np.array_equal(ac, ac[mp.iper - 1])


@pytest.mark.gpu
@pytest.mark.parametrize("sparse,LHSupd,ipred", [(False, 1, 1), (True, 1, 1), (True, 2, 1), (False, 1, 3)])
def test_timestep_parity(sparse, LHSupd, ipred):
    from oracle.oracle_py import Oracle
    from phasta_b200.solver import PhastaGPU
    params, tables, parts, states = smooth_case(etol=1e-5, minIters=0)
    o = Oracle(parts, params, tables, [(y.copy(order="F"), ac.copy(order="F")) for y, ac in states])
    g = PhastaGPU(parts[0], params, tables, device=0)
    if sparse:
        o.genadj()
        g.genadj()
    y, ac = states[0]
    g.set_state(y, ac)
    g.set_old_state(y, ac)
    for step in range(2):
        so = o.TimeStep(nitr=2, ipred=ipred, sparse=sparse, LHSupd=LHSupd)
        sg = g.TimeStep(nitr=2, ipred=ipred, sparse=sparse, LHSupd=LHSupd)
        assert np.array_equal(sg[:, 2:5], so[:, 2:5]), (sg, so)          # iKs, lGMRES, lhs
        assert np.allclose(sg[:, 1], so[:, 1], rtol=1e-8, atol=0)        # |b|
        assert np.allclose(sg[:, 0], so[:, 0], rtol=1e-6, atol=0)
        yg, acg, yog, acog = g.get_state(old=True)
        op = o.parts[0]
        assert rel_l2(yg, op.keep["y"]) < 1e-10
        assert rel_l2(yog, o.yold[0]) < 1e-10
        # ac = (y - yold) * Dtgl amplifies round-off of y by Dtgl * |y| / |ac|
        assert rel_l2(acg, op.keep["ac"]) < 1e-6
        assert rel_l2(acog, o.acold[0]) < 1e-6
    g.close()


@pytest.mark.gpu
def test_newton_shell_seams():
    """itrPredict / itrBC / itrCorrect / itrUpdate / rstat one by one."""
    from oracle.oracle_py import Oracle
    from phasta_b200.solver import PhastaGPU
    params, tables, parts, states = smooth_case(etol=1e-6)
    y, ac = states[0]
    g = PhastaGPU(parts[0], params, tables, device=0)
    rng = np.random.default_rng(7)
    yp = np.asfortranarray(y * (1 + 1e-3 * rng.standard_normal(y.shape)))
    acp = np.asfortranarray(1e2 * rng.standard_normal(y.shape))
    g.set_state(yp, acp)
    g.itrBC()
    o = Oracle(parts, params, tables, [(yp.copy(order="F"), acp.copy(order="F"))])
    o.itrBC()
    yg, acg = g.get_state()
    assert np.array_equal(yg, o.parts[0].keep["y"]) and np.array_equal(acg, o.parts[0].keep["ac"])
    # one solve, rstat, correct
    g.set_old_state(y, ac)
    res, Dy = g.SolGMRe(yg, acg)
    o.SolGMRe()
    nshgt = parts[0].nshg
    assert np.allclose(g.rstat(nshgt), o.rstat(nshgt), rtol=1e-9)
    g.itrCorrect()
    y2, ac2 = g.get_state()
    yo, aco = o.parts[0].keep["y"].copy(order="F"), o.parts[0].keep["ac"].copy(order="F")
    import ctypes as C
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    yold_f, acold_f = np.asfortranarray(y), np.asfortranarray(ac)
    o.L.orc_itrcorrect(C.byref(o.arr[0]), vp(yo), vp(aco), vp(yold_f), vp(acold_f), vp(o.parts[0].Dy))
    assert rel_l2(y2, yo) < 1e-10
    g.itrUpdate()
    _, _, yold, acold = g.get_state(old=True)
    assert rel_l2(yold, y2) < 1e-14      # backward Euler: yold <- y, acold <- ac
    g.close()


@pytest.mark.gpu
def test_timestep_partitioned():
    from oracle.oracle_py import Oracle
    from test_gpu_multipart import run_parts
    params, tables, parts, states = smooth_case(nparts=2, etol=1e-5)
    nshgt = int(max(p.gnode.max() for p in parts)) + 1
    o = Oracle(parts, params, tables, [(y.copy(order="F"), ac.copy(order="F")) for y, ac in states])
    so = [o.TimeStep(nitr=2, nshgt=nshgt) for _ in range(2)]

    def fn(g, y, ac):
        g.set_state(y, ac)
        g.set_old_state(y, ac)
        st = [g.TimeStep(nitr=2, nshgt=nshgt) for _ in range(2)]
        return st, g.get_state(old=True)

    gs, out = run_parts((params, tables, parts, states), fn)
    for i, (st, (yg, acg, yog, acog)) in enumerate(out):
        for a, b in zip(st, so):
            assert np.array_equal(a[:, 2:5], b[:, 2:5])
            assert np.allclose(a[:, 1], b[:, 1], rtol=1e-8, atol=0)
        assert rel_l2(yg, o.parts[i].keep["y"]) < 1e-10
        assert rel_l2(yog, o.yold[i]) < 1e-10
    [g.close() for g in gs]


# ------------------------------------------------------------------------------------------------
# one whole step of itrdrv.f's flow sequence executed by the reference's own Fortran (f77np):
# tests/golden/make_golden_step.py -> tests/golden/f77_step_*.npz
def _step_fixture(name):
    import os
    import sys
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    from make_golden_step import build_case
    from golden_cases import input_digest
    z = np.load(os.path.join(gold, "f77_step_%s.npz" % name))
    case, opt = build_case(name)
    assert np.array_equal(z["digest"], input_digest(case)), "seeded generators drifted from the fixture"
    params = case[0]
    for k in ("almi", "alfi", "gami", "Dtgl"):
        setattr(params, k, float(z[k]))        # what the reference's itrSetup derived from rhoinf / Delt
    return z, case, opt


STEP_CASES = ["be_channel", "genalpha_lhsupd2"]


def test_itrsetup_scalars_in_the_fixtures():
    z, _, opt = _step_fixture("genalpha_lhsupd2")
    rho = opt["rhoinf"]
    assert np.isclose(z["almi"], (3 - rho) / (1 + rho) / 2) and np.isclose(z["alfi"], 1 / (1 + rho))
    assert np.isclose(z["gami"], 0.5 + z["almi"] - z["alfi"])
    z, _, _ = _step_fixture("be_channel")
    assert (float(z["almi"]), float(z["alfi"]), float(z["gami"])) == (1.0, 1.0, 1.0)


@pytest.mark.parametrize("name", STEP_CASES)
def test_oracle_step_matches_reference_fortran(name):
    """predictor, nitr x (SolGMRe, itrCorrect, itrBC) with LHSupd reuse, itrUpdate -- against itrPC.f / itrbc.f /
    solgmr.f driven in itrdrv.f's order"""
    from oracle.oracle_py import Oracle
    z, case, opt = _step_fixture(name)
    params, tables, parts, states = case
    o = Oracle(parts, params, tables, [(y.copy(order="F"), ac.copy(order="F")) for y, ac in states])
    st = o.TimeStep(nitr=opt["nitr"], ipred=opt["ipred"], LHSupd=opt["LHSupd"])
    assert np.array_equal(st[:, 2].astype(int), z["iKs"]) and np.array_equal(st[:, 4].astype(int), z["lhs"])
    p = o.parts[0]
    assert rel_l2(p.keep["y"], z["y"]) < 1e-10 and rel_l2(o.yold[0], z["yold"]) < 1e-10
    assert rel_l2(p.keep["ac"], z["ac"]) < 1e-9 and rel_l2(o.acold[0], z["acold"]) < 1e-9
    assert np.allclose(st[:, 0], z["totres1"], rtol=1e-8)      # rstat's totres(1)


def _itrbc_fixture():
    import os
    from common import make_case
    from golden_cases import input_digest
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(gold, "f77_itrbc_allcodes.npz"))
    case = make_case(4, 4, 3, bc="allcodes", ibksiz=50)
    assert np.array_equal(z["digest"], input_digest(case))
    return z, case


def test_oracle_itrbc_all_codes_matches_reference_fortran():
    """itrbc.f on every essential-BC code (velocity 1..7, density -> pressure into y(:,1), pressure, temperature,
    periodic slaves): bit for bit"""
    from common import make_oracle
    _step_fixture("be_channel")          # puts tests/golden on sys.path
    z, case = _itrbc_fixture()
    iBC = case[2][0].iBC
    assert set(np.unique((iBC >> 3) & 7)) == set(range(8)) and (iBC & 1).any() and (iBC & 4).any()
    o = make_oracle(case)
    o.itrBC()
    assert np.array_equal(o.parts[0].keep["y"], z["y"]) and np.array_equal(o.parts[0].keep["ac"], z["ac"])
