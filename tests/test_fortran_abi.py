"""The reference's own entry points -- solgmre_, solgmrs_, solmfg_ with the argument lists of solgmr.f / solmfg.f and
the COMMON blocks as hidden inputs (phasta_b200/csrc/fortran_abi.c -> libphb200_f.so).  CPU: the library exports
them, and the C argument lists are the Fortran subroutine statements name for name.  GPU: a C stand-in for the
Fortran executable (tests/fortran_abi/commons.c) owns the COMMON blocks, registers the block pointers as the
five-line Fortran hook would, calls the three routines and must get what phb200_solgmre / solgmrs / solmfg return."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from common import make_case, rel_l2
from phasta_b200 import lib as _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBF = os.path.join(ROOT, "phasta_b200", "libphb200_f.so")
SRC = os.path.join(ROOT, "phasta_b200", "csrc", "fortran_abi.c")
REF = "/root/reference/phSolver/compressible"

# the dummy arguments of the three subroutine statements (solgmr.f:1-5, solgmr.f:368-373, solmfg.f:1-5)
ARGS = {
    "solgmre": "y ac yold acold x iBC BC EGmass res BDiag HBrg eBrg yBrg Rcos Rsin iper ilwork shp shgl shpb shglb "
               "Dy rerr".split(),
    "solgmrs": "y ac yold acold x iBC BC col row lhsk res BDiag HBrg eBrg yBrg Rcos Rsin iper ilwork shp shgl shpb "
               "shglb Dy rerr".split(),
    "solmfg": "y ac yold acold x iBC BC res BDiag HBrg eBrg yBrg Rcos Rsin iper ilwork shp shgl shpb shglb Dy "
              "rerr".split(),
}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c_args(name):
    src = open(SRC).read()
    m = re.search(r"void %s_\((.*?)\)\s*\{" % name, src, re.S)
    return [a.split()[-1].lstrip("*") for a in m.group(1).replace("\n", " ").split(",")]


def _fortran_args(path, name):
    txt = open(path, errors="replace").read()
    m = re.search(r"subroutine\s+%s\s*\((.*?)\)" % name, txt, re.S | re.I)
    body = re.sub(r"\n\s{5}\S", " ", m.group(1))          # continuation lines
    return [a.strip() for a in body.replace("\n", " ").replace("\t", " ").split(",")]


@pytest.mark.parametrize("name", list(ARGS))
def test_exported_with_the_reference_argument_list(name):
    out = subprocess.run(["nm", "-D", "--defined-only", LIBF], capture_output=True, text=True, check=True).stdout
    assert re.search(r" T %s_$" % name, out, re.M), "libphb200_f.so does not export %s_" % name
    assert [a.lower() for a in _c_args(name)] == [a.lower() for a in ARGS[name]]
    f = {"solgmre": ("solgmr.f", "SolGMRe"), "solgmrs": ("solgmr.f", "SolGMRs"), "solmfg": ("solmfg.f", "SolMFG")}[name]
    if os.path.exists(os.path.join(REF, f[0])):            # pin the list above to the reference source when it is here
        assert [a.lower() for a in _fortran_args(os.path.join(REF, f[0]), f[1])] == [a.lower() for a in ARGS[name]]


def test_library_exports_every_symbol_the_header_declares():
    txt = open(os.path.join(ROOT, "include", "phb200_fortran.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    syms = sorted(set(re.findall(r"\bvoid\s+([a-z0-9_]+_)\s*\(", txt)))
    assert syms == sorted(["solgmre_", "solgmrs_", "solmfg_", "phb200_register_block_", "phb200_register_blockb_",
                           "phb200_fortran_unique_id_", "phb200_fortran_comm_id_", "phb200_fortran_finalize_"])
    out = subprocess.run(["nm", "-D", "--defined-only", LIBF], capture_output=True, text=True, check=True).stdout
    for s_ in syms:
        assert re.search(r" T %s$" % s_, out, re.M), "libphb200_f.so does not export %s" % s_
    # and the definitions compile against the declarations (same argument types)
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-include",
                           os.path.join(ROOT, "include", "phb200_fortran.h"), SRC])


def test_common_blocks_stay_undefined_in_the_drop_in():
    """the Fortran executable owns the COMMON storage: the library must only reference it"""
    out = subprocess.run(["nm", "-D", LIBF], capture_output=True, text=True, check=True).stdout
    for blk in ("conpar", "genpar", "timdat", "solpar", "itrpar", "blkdat", "intpt", "workfc", "fronts", "elmpar",
                "matdat", "mmatpar", "precis", "outpar", "incomp", "shpdat"):
        assert re.search(r"^\s+U %s_$" % blk, out, re.M), blk


# ------------------------------------------------- the host side of the boundary, against a recording stub of the C-ABI
def test_common_blocks_reach_the_c_abi_and_itrpar_comes_back(tmp_path):
    """fortran_abi.c compiled together with a stub of libphb200.so that records its arguments: every COMMON scalar the
    path reads must arrive in phb200_common / phb200_step, the registered block pointers and the call's own array
    arguments must be the ones handed to phb200_init / phb200_set_sparse, and iKs / lGMRES / ntotGM (iKss ... for
    SolGMRs, eGMRES for SolMFG) must come back in COMMON /itrpar/."""
    from phasta_b200.lib import PhbCommon, PhbStep
    so = str(tmp_path / "libfabi_stub.so")
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-o", so, SRC, os.path.join(HERE, "fortran_abi", "commons.c"),
                           os.path.join(HERE, "fortran_abi", "stub_phb200.c")])
    L = C.CDLL(so)
    case = make_case(6, 4, 3, bc="channel", ibksiz=50, etol=3e-5, Kspace=37)
    params, tables, parts, states = case
    mp = parts[0]
    # the two structs exactly as the ctypes binding of the product fills them (without creating a device context)
    c = PhbCommon()
    c.nshg, c.numnp, c.numel, c.numelb = mp.nshg, mp.numnp, mp.numel, 0
    c.nflow, c.ndof, c.ndofBC, c.nshape, c.nedof = 5, 5, 6, 4, 20
    c.nelblk, c.nelblb, c.nlwork, c.numpe, c.myrank = mp.nelblk, 0, mp.nlwork, 1, 0
    for nm in ("ipord", "idiff", "itau", "iremoveStabTimeTerm", "EntropyPressure", "iDC", "Navier", "Kspace", "nGMRES",
               "minIters", "matflg2", "matflg3"):
        setattr(c, nm, int(getattr(params, nm)))
    for nm in ("Rgas", "gamma", "gamma1", "pr", "datmat121", "datmat221", "datmat321", "datmat131", "epsM", "dtsfct",
               "taucfct", "temper"):
        setattr(c, nm, float(getattr(params, nm)))
    for i in range(6):
        c.nint[i], c.nintb[i] = int(tables["nint"][i]), int(tables["nintb"][i])
    q = np.asfortranarray(tables["Qwt"]).ravel(order="F")
    C.memmove(c.Qwt, q.ctypes.data, q.nbytes)
    st = PhbStep()
    st.lhs, st.iprec, st.iter, st.nitr, st.lstep, st.istep = 1, 1, 2, 3, 120, 7
    st.Dtgl, st.almi, st.alfi, st.gami, st.etol = 1.0e3, 0.9, 0.8, 0.7, 3e-5
    lcblk = np.asfortranarray(mp.lcblk, dtype=np.int32)
    L.drv_fill_commons(C.byref(c), C.byref(st), _p(lcblk), None, 4321)
    mien = [np.asfortranarray(b, dtype=np.int32) for b in mp.mien]
    for b, ien in enumerate(mien):
        L.phb200_register_block_(C.byref(C.c_int(b + 1)), _p(ien))
    nshg = mp.nshg
    arr = {k: np.zeros(8) for k in ("yold", "acold", "EG", "BD", "H", "e", "yb", "rc", "rs", "rerr", "lhsk")}
    y, ac, res, Dy = (np.zeros((nshg, 5), order="F") for _ in range(4))
    x, iBC, BC = np.asfortranarray(mp.x), np.ascontiguousarray(mp.iBC, dtype=np.int32), np.asfortranarray(mp.BC)
    iper, il = np.ascontiguousarray(mp.iper, dtype=np.int32), np.ascontiguousarray(mp.ilwork, dtype=np.int32)
    tabs = [np.asfortranarray(tables[k], dtype=np.float64) for k in ("shp", "shgl", "shpb", "shglb")]
    L.solgmre_(_p(y), _p(ac), _p(arr["yold"]), _p(arr["acold"]), _p(x), _p(iBC), _p(BC), _p(arr["EG"]), _p(res),
               _p(arr["BD"]), _p(arr["H"]), _p(arr["e"]), _p(arr["yb"]), _p(arr["rc"]), _p(arr["rs"]), _p(iper), _p(il),
               *[_p(t) for t in tabs], _p(Dy), _p(arr["rerr"]))
    got = PhbCommon.in_dll(L, "stub_common")
    for name, _ in PhbCommon._fields_:
        a, b = getattr(got, name), getattr(c, name)
        if hasattr(a, "__len__"):
            if name in ("Qwtb",):
                continue                                    # not set above
            assert list(a) == list(b), name
        else:
            assert a == b, name
    gst = PhbStep.in_dll(L, "stub_step")
    for name, _ in PhbStep._fields_:
        assert getattr(gst, name) == getattr(st, name), name
    ptrs = (C.c_void_p * 16).in_dll(L, "stub_ptrs")
    want = [lcblk, mien[0], mien[-1], x, iBC, BC, iper, il] + tabs
    # lcblk is the COMMON block's own storage (filled from `lcblk` by the stand-in), the rest are the caller's arrays
    assert [ptrs[i] for i in range(1, 12)] == [w.ctypes.data for w in want[1:]]
    itr, eg = (C.c_int * 6)(), C.c_double(0)
    L.drv_get_itrpar(itr, C.byref(eg))
    assert (res[0, 0], Dy[0, 0]) == (11.0, 12.0) and list(itr)[:3] == [17, 0, 17]
    # SolGMRs: colm / rowp go to phb200_set_sparse once, with COMMON nnz_tot; counters land in iKss / lGMRESs / ntotGMs
    colm, rowp = np.arange(5, dtype=np.int32), np.arange(7, dtype=np.int32)
    for _ in range(2):
        L.solgmrs_(_p(y), _p(ac), _p(arr["yold"]), _p(arr["acold"]), _p(x), _p(iBC), _p(BC), _p(colm), _p(rowp),
                   _p(arr["lhsk"]), _p(res), _p(arr["BD"]), _p(arr["H"]), _p(arr["e"]), _p(arr["yb"]), _p(arr["rc"]),
                   _p(arr["rs"]), _p(iper), _p(il), *[_p(t) for t in tabs], _p(Dy), _p(arr["rerr"]))
    calls = (C.c_int * 8).in_dll(L, "stub_calls")
    assert list(calls)[:6] == [1, 1, 2, 0, 1, 0]            # one context, one set_sparse, two sparse solves
    assert (ptrs[12], ptrs[13], C.c_int.in_dll(L, "stub_nnz_tot").value) == (colm.ctypes.data, rowp.ctypes.data, 4321)
    L.drv_get_itrpar(itr, C.byref(eg))
    assert list(itr) == [17, 0, 17, 9, 1, 18]
    L.solmfg_(_p(y), _p(ac), _p(arr["yold"]), _p(arr["acold"]), _p(x), _p(iBC), _p(BC), _p(res), _p(arr["BD"]),
              _p(arr["H"]), _p(arr["e"]), _p(arr["yb"]), _p(arr["rc"]), _p(arr["rs"]), _p(iper), _p(il),
              *[_p(t) for t in tabs], _p(Dy), _p(arr["rerr"]))
    L.drv_get_itrpar(itr, C.byref(eg))
    assert list(itr)[:3] == [5, 0, 22] and (res[0, 0], Dy[0, 0]) == (31.0, 32.0)
    L.phb200_fortran_finalize_()
    assert list(calls)[5] == 1


# ---------------------------------------------------------------------------------------------------- GPU
def _stand_in():
    bdir = os.path.join(HERE, "fortran_abi", "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libcommons.so")
    src = os.path.join(HERE, "fortran_abi", "commons.c")
    hdr = os.path.join(ROOT, "phasta_b200", "csrc", "fortran_commons.h")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-o", so, src])
    com = C.CDLL(so, mode=C.RTLD_GLOBAL)        # the COMMON blocks must be visible when the drop-in is loaded
    return com, C.CDLL(LIBF)


@pytest.mark.gpu
def test_gpu_reference_entry_points_match_the_c_abi():
    from phasta_b200.solver import PhastaGPU
    case = make_case(8, 5, 4, bc="channel", etol=1e-7, Kspace=30)
    params, tables, parts, states = case
    mp = parts[0]
    y, ac = states[0]
    g = PhastaGPU(mp, params, tables, device=0)           # the C-ABI path: what the drop-in must reproduce
    res0, Dy0 = g.SolGMRe(y, ac)
    iKs0 = g.iKs
    H0 = g.HBrg.copy()
    colm, rowp, nnz_tot = g.genadj()
    colm, rowp = colm.copy(), g.rowp.copy()
    res1, Dy1 = g.SolGMRs(y, ac)
    iKs1 = g.iKs

    com, f = _stand_in()
    st = g.step()
    k = g._keep
    com.drv_fill_commons(C.byref(g.common), C.byref(st), _p(k["lcblk"]), None, int(nnz_tot))
    for b, ien in enumerate(k["mien"]):
        f.phb200_register_block_(C.byref(C.c_int(b + 1)), _p(ien))
    nshg, K = mp.nshg, params.Kspace
    vec = lambda n=5: np.zeros((nshg, n), order="F")  # noqa: E731
    yold, acold, res, Dy, rerr = y.copy(order="F"), ac.copy(order="F"), vec(), vec(), vec(10)
    BD = np.zeros((nshg, 5, 5), order="F")
    H, e, yb, rc, rs = np.zeros((K + 1, K), order="F"), np.zeros(K + 1), np.zeros(K + 1), np.zeros(K + 1), np.zeros(K + 1)
    yf, acf = np.asfortranarray(y), np.asfortranarray(ac)
    EG = np.zeros(1)                                        # never touched: EGmass stays in HBM
    f.solgmre_(_p(yf), _p(acf), _p(yold), _p(acold), _p(k["x"]), _p(k["iBC"]), _p(k["BC"]), _p(EG), _p(res), _p(BD),
               _p(H), _p(e), _p(yb), _p(rc), _p(rs), _p(k["iper"]), _p(k["ilwork"]), _p(k["shp"]), _p(k["shgl"]),
               _p(k["shpb"]), _p(k["shglb"]), _p(Dy), _p(rerr))
    itr, eg = (C.c_int * 6)(), C.c_double(0)
    com.drv_get_itrpar(itr, C.byref(eg))
    assert itr[0] == iKs0 and itr[2] == iKs0
    assert rel_l2(res, res0) < 1e-12 and rel_l2(Dy, Dy0) < 1e-10 and rel_l2(H, H0) < 1e-8
    lhsk = np.zeros(1)
    res[:], Dy[:] = 0, 0
    f.solgmrs_(_p(yf), _p(acf), _p(yold), _p(acold), _p(k["x"]), _p(k["iBC"]), _p(k["BC"]), _p(colm), _p(rowp),
               _p(lhsk), _p(res), _p(BD), _p(H), _p(e), _p(yb), _p(rc), _p(rs), _p(k["iper"]), _p(k["ilwork"]),
               _p(k["shp"]), _p(k["shgl"]), _p(k["shpb"]), _p(k["shglb"]), _p(Dy), _p(rerr))
    com.drv_get_itrpar(itr, C.byref(eg))
    assert itr[3] == iKs1 and itr[5] == iKs1
    assert rel_l2(res, res1) < 1e-12 and rel_l2(Dy, Dy1) < 1e-10
    f.phb200_fortran_finalize_()
    g.close()
