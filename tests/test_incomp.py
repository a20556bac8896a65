"""Incompressible element assembly into block-CSR + the lesSparse matrix-vector products
(BASELINE.json configs[3]): the oracle against the reference's own Fortran executed by f77np
(tests/golden/f77_incomp_*.npz, made by tests/golden/make_golden_incomp.py), oracle self-checks, and the CUDA
path against both (GPU)."""
import os
import sys

import numpy as np
import pytest

from common import make_case, make_oracle, rel_l2
from phasta_b200 import IncompParams

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
from make_golden_incomp import CASES, build_case  # noqa: E402
from golden_cases import input_digest  # noqa: E402

TOL_ASM = 1e-10


def load(name):
    z = np.load(os.path.join(GOLD, "f77_incomp_%s.npz" % name))
    case, ip = build_case(name)
    assert np.array_equal(z["digest"], input_digest(case)), "seeded generators drifted from the fixture"
    return z, case, ip


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_fortran_bit_for_bit(name):
    z, case, ip = load(name)
    o = make_oracle(case)
    assert o.genadj()[0] == int(z["nnz_tot"])
    p = o.parts[0]
    assert np.array_equal(p.colm, z["colm"]) and np.array_equal(p.rowp, z["rowp"])
    o.IncElmGMR(ip)
    assert np.array_equal(p.res4, z["res"])
    if "flxID" in z:        # the boundary integral (incompressible/asbmfg.f, e3b.f, e3bvar.f): /aerfrc/
        assert np.array_equal(p.aerfrc[4:4 + 70].reshape((10, 7), order="F")[:5], z["flxID"])
        assert np.allclose(p.aerfrc[:3], z["Force"], rtol=1e-13, atol=0)      # sum() order in e3b.f:236-238
    if ip.lhs:
        assert np.array_equal(p.lhsK9, z["lhsK"])
        assert np.array_equal(p.lhsP4, z["lhsP"])
        pin = z["ap_in"]
        assert np.array_equal(o.LesAp("G", pin[:, 3].copy()), z["apG"])
        assert np.array_equal(o.LesAp("KG", pin), z["apKG"])
        assert np.array_equal(o.LesAp("NGt", pin[:, :3]), z["apNGt"])
        assert np.array_equal(o.LesAp("NGtC", pin), z["apNGtC"])
        assert np.array_equal(o.LesAp("Full", pin), z["apFull"])


def test_every_velocity_code_is_in_the_fixtures():
    _, case, _ = load("tet_allbc")
    iBC = case[2][0].iBC
    assert set(np.unique((iBC >> 3) & 7)) == set(range(8))
    assert (iBC & 4).any() and (iBC & (1 << 10)).any()


def _dense(o, part=0):
    """the coupled 4x4-block operator of fLesSparseApFull as a dense matrix (node-major unknowns)"""
    p = o.parts[part]
    n = p.mp.nshg
    A = np.zeros((n, 4, n, 4))
    for i in range(n):
        for k in range(p.colm[i] - 1, p.colm[i + 1] - 1):
            j = p.rowp[k] - 1
            K = p.lhsK9[:, k].reshape(3, 3)          # kLhs(3(c-1)+r): column-major 3x3 (lesSparse.f:283-294)
            A[i, :3, j, :3] += K.T
            A[i, 3, j, :3] += p.lhsP4[:3, k]
            A[i, 3, j, 3] += p.lhsP4[3, k]
            A[j, :3, i, 3] -= p.lhsP4[:3, k]
    return A.reshape(4 * n, 4 * n)


def test_apfull_is_the_dense_operator():
    case = make_case(4, 3, 3, bc="allcodes")
    o = make_oracle(case)
    o.genadj()
    o.IncElmGMR(IncompParams())
    n = o.parts[0].mp.nshg
    v = np.random.default_rng(3).standard_normal((n, 4))
    q = o.LesAp("Full", v)
    assert rel_l2(q.ravel(), _dense(o) @ v.ravel()) < 1e-13
    # the sub-products are the blocks of the same operator
    assert rel_l2(o.LesAp("KG", v), q[:, :3]) < 1e-13
    assert rel_l2(o.LesAp("NGtC", v), q[:, 3]) < 1e-13
    v0 = v.copy()
    v0[:, 3] = 0.0
    assert rel_l2(o.LesAp("NGt", v[:, :3]), o.LesAp("Full", v0)[:, 3]) < 1e-13


def test_uniform_flow_has_zero_interior_residual():
    """patch test: uniform velocity and pressure, no acceleration, no BCs -> res = 0 on interior nodes"""
    params, tables, parts, states = make_case(5, 4, 4, bc="none", periodic_z=False)
    y, ac = states[0]
    y[:, 0], y[:, 1], y[:, 2], y[:, 3], y[:, 4] = 1.0, 0.3, -0.2, 2.0, 300.0
    ac[:] = 0.0
    o = make_oracle((params, tables, parts, [(y, ac)]))
    o.genadj()
    o.IncElmGMR(IncompParams())
    g = parts[0].gnode
    nyp, nzp = 5, 5
    i, j, k = g // (nyp * nzp), (g // nzp) % nyp, g % nzp
    interior = (i > 0) & (i < 5) & (j > 0) & (j < 4) & (k > 0) & (k < 4)
    r = o.parts[0].res4
    assert np.abs(r[interior]).max() < 1e-12 * max(1.0, np.abs(r).max())


def test_tangent_is_the_derivative_of_the_residual_in_pressure_and_acceleration():
    """The velocity tangent freezes tau and the advective velocity (e3lhs.f:35-43 "lazy tangent"), but the
    pressure columns G and the mass term are exact: -d res/d p . dp = [-G^T; C] dp up to the frozen tau_M."""
    case = make_case(4, 3, 3, bc="none", periodic_z=False)
    params, tables, parts, states = case
    # small time step: tau_M -> Delt/2, so the SUPG pressure term tau_M (u.grad N_a) grad p, which the
    # tangent does not carry, is O(Delt |u| / h) of the Galerkin one
    ip = IncompParams(idiff=0, Delt=1.0e-5)
    o = make_oracle(case)
    o.genadj()
    o.IncElmGMR(ip)
    p = o.parts[0]
    r0 = p.res4.copy()
    n = p.mp.nshg
    dp = np.random.default_rng(5).standard_normal(n)
    eps = 1e-3
    y, ac = states[0]
    y2 = y.copy(order="F")
    y2[:, 3] += eps * dp
    o2 = make_oracle((params, tables, parts, [(y2, ac)]))
    o2.genadj()
    o2.IncElmGMR(ip)
    fd = -(o2.parts[0].res4 - r0) / eps               # res is -G(Y): the solve is K dY = res
    v = np.zeros((n, 4))
    v[:, 3] = dp
    lhsFct = ip.alfi * ip.gami * ip.Delt
    q = o.LesAp("Full", v) / lhsFct
    assert rel_l2(q[:, :3], fd[:, :3]) < 2e-2          # momentum rows: -G^T dp (+ tau_M terms the tangent drops)
    assert rel_l2(q[:, 3], fd[:, 3]) < 1e-6            # continuity row: C dp = tau_M grad N . grad dp, exact


def test_partitioned_equals_serial():
    ip = IncompParams()
    c1 = make_case(8, 3, 3, bc="allcodes", nparts=1)
    c2 = make_case(8, 3, 3, bc="allcodes", nparts=2)
    o1, o2 = make_oracle(c1), make_oracle(c2)
    o1.genadj()
    o2.genadj()
    o1.IncElmGMR(ip)
    o2.IncElmGMR(ip)
    g1 = c1[2][0].gnode
    ref = np.zeros((g1.max() + 1, 4))
    ref[g1] = o1.parts[0].res4
    for p in o2.parts:
        own = np.ones(p.mp.nshg, dtype=bool)
        il, pos = p.mp.ilwork, 1
        for _ in range(int(il[0])):
            iacc, nseg = il[pos + 1], il[pos + 3]
            if iacc == 0:
                for s in range(nseg):
                    a, ln = il[pos + 4 + 2 * s], il[pos + 5 + 2 * s]
                    own[a - 1:a - 1 + ln] = False
            pos += 4 + 2 * nseg
        assert rel_l2(p.res4[own], ref[p.mp.gnode[own]]) < 1e-12
        assert not p.res4[~own].any()                   # bc3per.f:28-43 zeroes the rows another part owns


# ------------------------------------------------------------------------------------------------ GPU
def _gpu(case):
    from phasta_b200.solver import PhastaGPU
    params, tables, parts, states = case
    g = PhastaGPU(parts[0], params, tables, device=0)
    g.genadj()
    return g


# fixtures added after the round's GPU budget was spent: their device runs are in tests/test_zz_gpu_late.py
LATE = ("tet_bnd", "hex_bnd", "mixed_bnd")


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in CASES if n not in LATE])
def test_gpu_matches_reference_fortran(name):
    """CUDA ElmGMR / fLesSparseAp* through the C-ABI against the reference's Fortran (f77np fixtures)."""
    check_gpu_case(name)


def check_gpu_case(name):
    z, case, ip = load(name)
    g = _gpu(case)
    y, ac = case[3][0]
    out = g.IncElmGMR(y, ac, ip)
    assert rel_l2(out["res"], z["res"]) < TOL_ASM
    if ip.lhs:
        assert rel_l2(out["lhsK"], z["lhsK"]) < TOL_ASM
        assert rel_l2(out["lhsP"], z["lhsP"]) < TOL_ASM
        # per-entry check so that one wrong block cannot hide in the norm
        d = np.abs(out["lhsK"] - z["lhsK"]).max(axis=0)
        s = np.abs(z["lhsK"]).max(axis=0) + 1e-300
        assert (d / np.maximum(s, 1e-8 * s.max())).max() < 1e-8
        pin = z["ap_in"]
        assert rel_l2(g.LesAp("G", pin[:, 3].copy()), z["apG"]) < TOL_ASM
        assert rel_l2(g.LesAp("KG", pin), z["apKG"]) < TOL_ASM
        assert rel_l2(g.LesAp("NGt", pin[:, :3]), z["apNGt"]) < TOL_ASM
        assert rel_l2(g.LesAp("NGtC", pin), z["apNGtC"]) < TOL_ASM
        assert rel_l2(g.LesAp("Full", pin), z["apFull"]) < TOL_ASM
    if "flxID" in z:
        Fo, _, fl = g.aerfrc()
        assert rel_l2(fl[:5, :7], z["flxID"]) < TOL_ASM
        assert rel_l2(Fo, z["Force"]) < TOL_ASM
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("topo,n", [("tet", (12, 10, 9)), ("mixed", (8, 10, 6)), ("hex", (9, 7, 6))])
def test_gpu_matches_oracle_on_a_larger_mesh(topo, n):
    case = make_case(*n, bc="allcodes", topo=topo, ibksiz=128)
    ip = IncompParams(rho=1.1, rmu=2.0e-3).with_rhoinf(0.2)
    o = make_oracle(case)
    o.genadj()
    o.IncElmGMR(ip)
    p = o.parts[0]
    g = _gpu(case)
    y, ac = case[3][0]
    out = g.IncElmGMR(y, ac, ip)
    assert rel_l2(out["res"], p.res4) < TOL_ASM
    assert rel_l2(out["lhsK"], p.lhsK9) < TOL_ASM
    assert rel_l2(out["lhsP"], p.lhsP4) < TOL_ASM
    v = np.random.default_rng(9).standard_normal((p.mp.nshg, 4))
    assert rel_l2(g.LesAp("Full", v), o.LesAp("Full", v)) < TOL_ASM
    # residual-only call leaves the resident matrices alone
    out0 = g.IncElmGMR(y, ac, ip, lhs=0)
    assert rel_l2(out0["res"], p.res4) < TOL_ASM
    assert rel_l2(g.LesAp("Full", v), o.LesAp("Full", v)) < TOL_ASM
    g.close()


@pytest.mark.gpu
def test_gpu_refuses_what_is_not_built():
    from phasta_b200.solver import PhastaError
    case = make_case(4, 3, 3, bc="channel")
    g = _gpu(case)
    y, ac = case[3][0]
    with pytest.raises(PhastaError):
        g.IncElmGMR(y, ac, IncompParams(itau=1))
    g.close()
