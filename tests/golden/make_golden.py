#!/usr/bin/env python
"""Generate tests/golden/tables_ref.npz from the REFERENCE's own C table
generators (phSolver/common/symtet.c, symtri.c, shptet.c, shptri.c and
shapeFunction/src/*.c), compiled into oracle/_ref/libref_tables.so by
oracle/Makefile.  Run in the build container (needs /root/reference); the
.npz travels with the repo so the checks run anywhere."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_tables.so"))
out = {}
for n in (1, 4):
    pt = np.zeros((n, 4))
    wt = np.zeros(n)
    err = C.c_int(0)
    L.symtet_(C.byref(C.c_int(n)), pt.ctypes.data_as(C.c_void_p), wt.ctypes.data_as(C.c_void_p), C.byref(err))
    out["tet%d_pt" % n], out["tet%d_wt" % n] = pt, wt
    N = np.zeros((n, 4))
    dN = np.zeros((n, 4, 3))
    for i in range(n):
        Ni = np.zeros(32)
        dNi = np.zeros((32, 3))
        par = pt[i, :3].copy()
        L.shptet_(C.byref(C.c_int(1)), par.ctypes.data_as(C.c_void_p), Ni.ctypes.data_as(C.c_void_p),
                  dNi.ctypes.data_as(C.c_void_p))
        N[i], dN[i] = Ni[:4], dNi[:4]
    out["tet%d_N" % n], out["tet%d_dN" % n] = N, dN
for n in (1, 3):
    pt = np.zeros((n, 4))
    wt = np.zeros(n)
    err = C.c_int(0)
    L.symtri_(C.byref(C.c_int(n)), pt.ctypes.data_as(C.c_void_p), wt.ctypes.data_as(C.c_void_p), C.byref(err))
    out["tri%d_pt" % n], out["tri%d_wt" % n] = pt, wt
# hexes (symhex 8-pt, shphex p=1) and wedges (symwdg 6-pt, shp6w p=1)
for name, n, sym, shpf, nsh in (("hex", 8, L.symhex_, L.shphex_, 8), ("wdg", 6, L.symwdg_, L.shp6w_, 6)):
    pt = np.zeros((n, 4))
    wt = np.zeros(n)
    err = C.c_int(0)
    sym(C.byref(C.c_int(n)), pt.ctypes.data_as(C.c_void_p), wt.ctypes.data_as(C.c_void_p), C.byref(err))
    out["%s%d_pt" % (name, n)], out["%s%d_wt" % (name, n)] = pt, wt
    N = np.zeros((n, nsh))
    dN = np.zeros((n, nsh, 3))
    for i in range(n):
        Ni = np.zeros(32)
        dNi = np.zeros((32, 3))
        par = pt[i, :3].copy()
        shpf(C.byref(C.c_int(1)), par.ctypes.data_as(C.c_void_p), Ni.ctypes.data_as(C.c_void_p),
             dNi.ctypes.data_as(C.c_void_p))
        N[i], dN[i] = Ni[:nsh], dNi[:nsh]
    out["%s%d_N" % (name, n)], out["%s%d_dN" % (name, n)] = N, dN
np.savez(os.path.join(ROOT, "tests", "golden", "tables_ref.npz"), **out)
for k, v in out.items():
    print(k, v.tolist())
