#!/usr/bin/env python
"""Quadrature / shape-function COMMON tables as the reference builds them: common/genint.f, genshp.f and genshpb.f
executed by f77np, with the reference's own C generators (oracle/_ref/libref_tables.so, compiled from
phSolver/common/sym*.c, shp*.c, newshape.cc and shapeFunction/src by oracle/Makefile) behind the calls those
routines make.  Writes tests/golden/tables_f77.npz: Qwt, Qwtb, nint, nintb, shp, shgl, shpb, shglb for
quadrature rule 2 (interior and boundary), ipord = 1, interior blocks of lcsyst 1..3, boundary blocks of
lcsyst 1..4.  Only runs where /root/reference exists; the tests read the committed .npz."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from f77np import Program, scan_functions  # noqa: E402

REF = "/root/reference/phSolver/common"
subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_tables.so"))
vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731


def sym(fn):
    def stub(prog, n, pt, wt, nerr):
        n = int(n)
        assert pt.flags.f_contiguous and pt.shape == (4, n) and wt.shape == (n,)
        err = C.c_int(0)
        fn(C.byref(C.c_int(n)), vp(pt), vp(wt), C.byref(err))   # pt[i][4] in C == pt(4,i) in Fortran
        return {3: err.value}
    return stub


def zero(prog, n, pt, wt, nerr):       # pyramids (lcsyst 5, 6) are not on the path
    pt[...] = 0.0
    wt[...] = 0.0
    return {3: 0}


def shapefn(fn):
    def stub(prog, p, par, N, dN):
        # actual arguments are sections (Qpt(1,1:3,i), shp(1,:,i), shgl(1,:,:,i)): copy-in / copy-out
        par_c = np.ascontiguousarray(par, dtype=np.float64)
        Nc = np.zeros(N.shape[0])
        dNc = np.zeros((dN.shape[1], 3))                         # C dN[a][3] == Fortran dN(3,a)
        fn(C.byref(C.c_int(int(p))), vp(par_c), vp(Nc), vp(dNc))
        N[...] = Nc
        dN[...] = dNc.T
    return stub


def main():
    stubs = dict(symtet=sym(L.symtet_), symtri=sym(L.symtri_), symhex=sym(L.symhex_), symquad=sym(L.symquad_),
                 symwdg=sym(L.symwdg_), symquadw=sym(L.symquadw_), sympyr=zero, symtripyr=zero,
                 shptet=shapefn(L.shptet_), shphex=shapefn(L.shphex_), shp6w=shapefn(L.shp6w_))
    prog = Program([REF], modules={}, stubs=stubs)
    for f in ("genint.f", "genshp.f", "genshpb.f"):
        scan_functions(os.path.join(REF, f))
        prog.load(os.path.join(REF, f))
    G = prog.G
    G.update(ipord=1, nen=4, nenb=3, nsd=3)
    G["intg"][...] = 0
    G["intg"][0, 0] = 2
    G["intg"][1, 0] = 2
    for k in ("qpt", "qwt", "qptb", "qwtb"):
        G[k][...] = 0.0
    G["nint"][...] = 0
    G["nintb"][...] = 0
    prog.call("genint")
    MAXTOP, MAXSH, MAXQPT = 6, 32, 125
    shp = np.zeros((MAXTOP, MAXSH, MAXQPT), order="F")
    shgl = np.zeros((MAXTOP, 3, MAXSH, MAXQPT), order="F")
    shpb = np.zeros((MAXTOP, MAXSH, MAXQPT), order="F")
    shglb = np.zeros((MAXTOP, 3, MAXSH, MAXQPT), order="F")
    # one interior block per topology 1..3, one boundary block per lcsyst 1..4 (rows 3 and 9/10 are what
    # genshp.f / genshpb.f read)
    G["lcblk"][...] = 0
    G["lcblkb"][...] = 0
    for b, (lcsyst, nshl) in enumerate(((1, 4), (2, 8), (3, 6))):
        G["lcblk"][2, b], G["lcblk"][9, b] = lcsyst, nshl
    for b, (lcsyst, nshl) in enumerate(((1, 4), (2, 8), (3, 6), (4, 6))):
        G["lcblkb"][2, b], G["lcblkb"][8, b] = lcsyst, nshl
    prog.call("genshp", shp, shgl, 0, 3)
    prog.call("genshpb", shpb, shglb, 0, 4)
    out = dict(Qwt=np.array(G["qwt"]), Qwtb=np.array(G["qwtb"]), nint=np.array(G["nint"]), nintb=np.array(G["nintb"]),
               Qpt=np.array(G["qpt"]), Qptb=np.array(G["qptb"]), shp=shp, shgl=shgl, shpb=shpb, shglb=shglb)
    np.savez_compressed(os.path.join(HERE, "tables_f77.npz"), **out)
    for k in ("nint", "nintb"):
        print(k, out[k].tolist())
    for top in range(4):
        n = int(out["nintb"][top])
        print("lcsyst", top + 1, "Qwtb", out["Qwtb"][top, :n].tolist())
        print("   Qptb", out["Qptb"][top, :3, :n].T.tolist())
        print("   shpb(q=1)", out["shpb"][top, :8, 0].tolist())


if __name__ == "__main__":
    main()
