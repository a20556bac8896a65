"""The product's boundary-flux kernel for hexes and wedges (phasta_b200/csrc/boundary.cuh: k_asbmfg_gen, with the
group packing and face tables of bnd_pack.h) compiled for the HOST behind tests/host_emul/cuda_shim.h and run one
thread after the other, against what the reference's asbmfg.f / e3b.f / e3bvar.f added to the residual and to
/aerfrc/ in the f77np fixtures.  This checks the kernel's arithmetic and data layout where there is no GPU; the
same cases run on the device in the GPU suite."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from common import rel_l2
from test_golden_f77 import load

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "bnd_host.cpp")
OUT = os.path.join(HERE, "host_emul", "_build", "libbnd_host.so")


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas",
                           "-o", OUT, SRC])
    return C.CDLL(OUT)


@pytest.fixture(scope="module")
def emul_inc():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    out = os.path.join(os.path.dirname(OUT), "libinc_bnd_host.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas",
                           "-o", out, os.path.join(HERE, "host_emul", "inc_bnd_host.cpp")])
    return C.CDLL(out)


def _ptrs(arrs, ctype):
    keep = [np.ascontiguousarray(np.asfortranarray(a).ravel(order="F")) for a in arrs]
    return keep, (C.POINTER(ctype) * len(keep))(*[a.ctypes.data_as(C.POINTER(ctype)) for a in keep])


@pytest.mark.parametrize("name", ["hex_bnd", "wedge_bnd", "mixed_bnd"])
def test_boundary_kernel_on_the_host_matches_reference_fortran(emul, name):
    z, case, _ = load(name)
    params, tables, parts, states = case
    mp = parts[0]
    y = np.asfortranarray(states[0][0])
    res = np.zeros((mp.nshg, 5), order="F")
    aer = np.zeros(4 + 10 * 1001)
    k1, pien = _ptrs([b.astype(np.int32) for b in mp.mienb], C.c_int)
    k2, pibc = _ptrs([b.astype(np.int32) for b in mp.miBCB], C.c_int)
    k3, pbcb = _ptrs([b.astype(np.float64) for b in mp.mBCB], C.c_double)
    lcb = np.ascontiguousarray(mp.lcblkb.T.astype(np.int32).ravel())           # column b at lcblkb + 10 b
    P = params
    phys = np.array([P.Rgas, P.gamma, P.gamma1, P.pr, P.datmat121, P.datmat221, P.datmat321, P.datmat131])
    iphys = np.array([P.matflg2, P.matflg3], dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)       # noqa: E731
    T = {k: np.asfortranarray(tables[k]) for k in ("nintb", "Qwtb", "shpb", "shglb")}
    nintb = T["nintb"].astype(np.int32)
    n = emul.bnd_host_asbmfg(mp.nelblb, vp(lcb), pien, pibc, pbcb, mp.nshg, mp.numnp, vp(np.asfortranarray(mp.x)),
                             vp(y), vp(nintb), vp(T["Qwtb"]), vp(T["shpb"]), vp(T["shglb"]), vp(phys), vp(iphys),
                             vp(res), vp(aer), 1)
    nontet = sum(b.shape[0] for b, lc in zip(mp.mienb, mp.lcblkb.T) if lc[2] != 1)
    assert n == nontet > 0
    ref = z["elmgmre.res_nobc"] - z["elmgmre.res_interior"]        # what the boundary blocks added
    if name == "mixed_bnd":
        # the tets' share comes from k_asbmfg_tet (GPU suite); compare on the nodes no boundary tet touches
        tetnodes = np.unique(np.concatenate([b[:, :3].ravel() for b, lc in zip(mp.mienb, mp.lcblkb.T) if lc[2] == 1])) - 1
        keep = np.ones(mp.nshg, dtype=bool)
        keep[tetnodes] = False
        assert keep.sum() > 0 and np.abs(ref[keep]).max() > 0
        assert rel_l2(res[keep], ref[keep]) < 1e-12
        return
    assert rel_l2(res, ref) < 1e-12
    ref_f = np.r_[z["elmgmre.Force"], z["elmgmre.HFlux"]]
    assert rel_l2(aer[:4], ref_f) < 1e-12
    fl = aer[4:4 + 20].reshape((10, 2), order="F")
    assert rel_l2(fl, z["elmgmre.flxID"]) < 1e-12


@pytest.mark.parametrize("name", ["tet_bnd", "hex_bnd", "mixed_bnd"])
def test_incompressible_boundary_kernel_on_the_host_matches_reference_fortran(emul_inc, name):
    """k_inc_asbmfg (inc_boundary.cuh) against incompressible/asbmfg.f + e3b.f + e3bvar.f: flxID and Force come from
    the boundary blocks alone; their share of the residual is isolated through the linearity of bc3Res:
    res(with boundary blocks) - res(without) = bc3Res(what the kernel scattered)."""
    import copy
    from common import make_oracle
    from test_incomp import load as load_inc
    z, case, ip = load_inc(name)
    params, tables, parts, states = case
    mp = parts[0]
    y = np.asfortranarray(states[0][0])
    contrib = np.zeros((mp.nshg, 4), order="F")
    aer = np.zeros(4 + 10 * 1001)
    k1, pien = _ptrs([b.astype(np.int32) for b in mp.mienb], C.c_int)
    k2, pibc = _ptrs([b.astype(np.int32) for b in mp.miBCB], C.c_int)
    k3, pbcb = _ptrs([b.astype(np.float64) for b in mp.mBCB], C.c_double)
    lcb = np.ascontiguousarray(mp.lcblkb.T.astype(np.int32).ravel())
    vp = lambda a: a.ctypes.data_as(C.c_void_p)       # noqa: E731
    T = {k: np.asfortranarray(tables[k]) for k in ("nintb", "Qwtb", "shpb", "shglb")}
    nintb = T["nintb"].astype(np.int32)
    nsrf = np.zeros(1001, dtype=np.int32)
    nsrf[list(ip.surfaces)] = 1
    emul_inc.inc_bnd_host_asbmfg.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
    n = emul_inc.inc_bnd_host_asbmfg(mp.nelblb, vp(lcb), pien, pibc, pbcb, mp.nshg, mp.numnp,
                                     vp(np.asfortranarray(mp.x)), vp(y), vp(nintb), vp(T["Qwtb"]), vp(T["shpb"]),
                                     vp(T["shglb"]), float(ip.rho), float(ip.rmu), int(ip.iviscflux),
                                     int(ip.iconvflow), int(ip.itwmod), vp(nsrf), vp(contrib), vp(aer))
    assert n == sum(b.shape[0] for b in mp.mienb) > 0
    fl = aer[4:4 + 70].reshape((10, 7), order="F")[:5]
    assert rel_l2(fl, z["flxID"]) < 1e-12 and np.abs(z["flxID"]).max() > 0
    assert rel_l2(aer[:3], z["Force"]) < 1e-12 and np.abs(z["Force"]).max() > 0
    # the residual: oracle == reference bit for bit with the boundary blocks (tests/test_incomp.py); without them:
    o = make_oracle(case)
    o.genadj()
    o.IncElmGMR(ip)
    assert np.array_equal(o.parts[0].res4, z["res"])
    bare = copy.copy(mp)
    bare.lcblkb, bare.mienb, bare.miBCB, bare.mBCB = None, [], [], []
    o2 = make_oracle((params, tables, [bare], states))
    o2.genadj()
    o2.IncElmGMR(ip)
    ref = z["res"] - o2.parts[0].res4
    o.IncBc3Res(contrib)
    assert np.abs(ref).max() > 0
    assert np.abs(contrib - ref).max() < 1e-11 * np.abs(z["res"]).max()


# ------------------------------------------------------------------------------------------------
# the hex / wedge assembly kernel (k_asigmr_gen in assembly.cu: AsIGMR + e3 [+ e3dc] + BDiag + bc3LHS) on the host:
# assembly.cu's device code compiled with g++ behind tests/host_emul/cuda_shim_simt.h (one pthread per CUDA thread)
@pytest.fixture(scope="module")
def emul_asm():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    out = os.path.join(os.path.dirname(OUT), "libasm_host.so")
    subprocess.check_call(["g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas",
                           "-pthread", "-x", "c++", "-o", out, os.path.join(HERE, "host_emul", "asm_host.cpp")])
    return C.CDLL(out)


@pytest.mark.parametrize("name,lhs", [("hex_channel", 1), ("hex_dc1", 1), ("wedge_dc3", 1), ("wedge_dc3", 0),
                                      ("wedge_allbc", 1)])
def test_hex_wedge_assembly_kernel_on_the_host_matches_reference_fortran(emul_asm, name, lhs):
    """hex_channel / wedge_allbc have run on a B200 (they calibrate the emulation); the discontinuity-capturing
    instantiations (DCON) of the same kernel have not"""
    z, case, _ = load(name)
    params, tables, parts, states = case
    mp = parts[0]
    P = params
    run = "elmgmre" if lhs else "elmgmre0"
    nshl = mp.mien[0].shape[1]
    ien = np.concatenate([np.asarray(b) for b in mp.mien], axis=0).astype(np.int32) - 1     # (numel,nshl)
    numel = ien.shape[0]
    pad = (numel + 31) // 32 * 32
    ienp = np.zeros((nshl, pad), dtype=np.int32)
    ienp[:, :numel] = ien.T
    y, ac = (np.asfortranarray(a) for a in states[0])
    q = np.asfortranarray(z[run + ".qres"])
    res = np.zeros((mp.nshg, 5), order="F")
    BDiag = np.zeros((mp.nshg, 5, 5), order="F")
    nedof = 5 * nshl
    EG = np.zeros(pad * nedof * nedof)
    fct1 = P.almi / P.gami / P.alfi * P.Dtgl
    phys = np.array([P.Rgas, P.gamma, P.gamma1, P.pr, P.datmat121, P.datmat221, P.datmat321, P.datmat131, P.dtsfct,
                     P.taucfct, P.temper, P.Dtgl, fct1, P.epsM])
    iphys = np.array([P.matflg2, P.matflg3, P.idiff, P.iremoveStabTimeTerm, P.ipord, lhs, lhs, P.iDC], dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)       # noqa: E731
    T = {k: np.asfortranarray(tables[k]) for k in ("Qwt", "shp", "shgl")}
    nint = np.asarray(tables["nint"]).astype(np.int32)
    iBC = np.ascontiguousarray(mp.iBC, dtype=np.int32)
    BC = np.asfortranarray(mp.BC)
    rc = emul_asm.asm_host_gen(nshl, lhs, numel, mp.nshg, mp.numnp, vp(ienp), vp(np.asfortranarray(mp.x)), vp(y), vp(ac),
                               vp(q), vp(iBC), vp(BC), vp(nint), vp(T["Qwt"]), vp(T["shp"]), vp(T["shgl"]), vp(phys),
                               vp(iphys), vp(res), vp(BDiag), vp(EG))
    assert rc == 0
    assert rel_l2(res, z[run + ".res_interior"]) < 1e-12
    if lhs:
        assert rel_l2(BDiag, z[run + ".BDiag_nobc"]) < 1e-12
        # EBE tiles EG[tile][c][r][lane] -> EGmass(e, r, c) (post bc3LHS, asigmr.f:36 + bc3lhs.f)
        eg = EG.reshape(pad // 32, nedof, nedof, 32).transpose(0, 3, 2, 1).reshape(pad, nedof, nedof)[:numel]
        assert rel_l2(eg, z[run + ".EGmass"]) < 1e-12
    if P.iDC:
        # the fixture is sensitive to the operator: the same kernel without it is far off
        iphys[7] = 0
        res0 = np.zeros_like(res)
        EG0, BD0 = np.zeros_like(EG), np.zeros_like(BDiag)
        emul_asm.asm_host_gen(nshl, lhs, numel, mp.nshg, mp.numnp, vp(ienp), vp(np.asfortranarray(mp.x)), vp(y), vp(ac),
                              vp(q), vp(iBC), vp(BC), vp(nint), vp(T["Qwt"]), vp(T["shp"]), vp(T["shgl"]), vp(phys),
                              vp(iphys), vp(res0), vp(BD0), vp(EG0))
        assert rel_l2(res0, z[run + ".res_interior"]) > 1e-3
        if lhs:
            assert rel_l2(EG0, EG) > 1e-3


@pytest.mark.parametrize("name", ["hex_nd_mfg_dc3_raw", "tet_nd_mfg_dc1_raw"])
def test_matrix_free_kernels_with_dc_on_the_host_match_reference_fortran(emul_asm, name):
    """The matrix-free flavour with discontinuity capturing, kernel source on the host against the reference's
    solmfg.f chain on a case without essential BCs (residuals and block diagonal = raw element sums):
    ElmMFG's element pass (k_asigmr_tet / k_asigmr_gen in e3bdg mode with DCON: DC flux in res, none in BDiag),
    its modified residual (k_asires DCM 3: ires=3, incl. the rmi(:,11) statement of e3dc.f:262) and ItrRes
    (k_asires DCM 2) with iabres 0 and 1."""
    from common import make_oracle
    z, case, _ = load(name)
    params, tables, parts, states = case
    mp, P = parts[0], params
    assert not mp.iBC.any() and P.iDC != 0
    nshl = mp.mien[0].shape[1]
    ien = np.concatenate([np.asarray(b) for b in mp.mien], axis=0).astype(np.int32) - 1
    numel = ien.shape[0]
    pad = (numel + 31) // 32 * 32
    ienp = np.zeros((nshl, pad), dtype=np.int32)
    ienp[:, :numel] = ien.T
    y, ac = np.asfortranarray(z["solmfg.y_bc"]), np.asfortranarray(z["solmfg.ac_bc"])
    o = make_oracle(case)
    o.itrBC()
    o.set_flags(lhs=0, iprec=1)
    o.ElmMFG()
    q = np.asfortranarray(o.parts[0].qres)              # q = qres / rmass after qpbc (bit-equal to the reference's)
    fct1 = P.almi / P.gami / P.alfi * P.Dtgl
    phys = np.array([P.Rgas, P.gamma, P.gamma1, P.pr, P.datmat121, P.datmat221, P.datmat321, P.datmat131, P.dtsfct,
                     P.taucfct, P.temper, P.Dtgl, fct1, P.epsM])
    iphys = np.array([P.matflg2, P.matflg3, P.idiff, P.iremoveStabTimeTerm, P.ipord, 0, 1, P.iDC], dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)       # noqa: E731
    T = {k: np.asfortranarray(tables[k]) for k in ("Qwt", "shp", "shgl")}
    nint = np.asarray(tables["nint"]).astype(np.int32)
    iBC = np.ascontiguousarray(mp.iBC, dtype=np.int32)
    BC = np.asfortranarray(mp.BC)
    x = np.asfortranarray(mp.x)
    res = np.zeros((mp.nshg, 5), order="F")
    BDiag = np.zeros((mp.nshg, 5, 5), order="F")
    assert emul_asm.asm_host_bdg(nshl, numel, mp.nshg, mp.numnp, vp(ienp), vp(x), vp(y), vp(ac), vp(q), vp(iBC), vp(BC),
                                 vp(nint), vp(T["Qwt"]), vp(T["shp"]), vp(T["shgl"]), vp(phys), vp(iphys), vp(res),
                                 vp(BDiag)) == 0
    assert rel_l2(res, z["solmfg.elm_res"]) < 1e-11          # the residual nearly cancels: |res| << |terms|
    assert rel_l2(BDiag, z["solmfg.elm_BDiag"]) < 1e-12

    def asires(ires, iabres, yp):
        out = np.zeros((mp.nshg, 5), order="F")
        yp = np.asfortranarray(yp)
        assert emul_asm.asm_host_asires(nshl, ires, iabres, numel, mp.nshg, mp.numnp, vp(ienp), vp(x), vp(y), vp(ac),
                                        vp(q), vp(yp), vp(nint), vp(T["Qwt"]), vp(T["shp"]), vp(T["shgl"]), vp(phys),
                                        vp(iphys), vp(out)) == 0
        return out

    assert rel_l2(asires(3, 0, y), z["solmfg.elm_rmes"]) < 1e-12
    r2 = asires(2, 0, y)
    assert rel_l2(r2, z["solmfg.elm_rmes"]) > 1e-6           # ires=2 on the same state is a different vector
    for iab in (0, 1):
        assert rel_l2(asires(2, iab, z["solmfg.itrres_in"]), z["solmfg.itrres_out%d" % iab]) < 1e-12
