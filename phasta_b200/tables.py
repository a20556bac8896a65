"""Quadrature / shape-function tables in the reference's COMMON layout.

Qwt(MAXTOP,MAXQPT), nint(MAXTOP)        phSolver/common/common.h:92-96
shp(MAXTOP,MAXSH,MAXQPT), shgl(MAXTOP,3,MAXSH,MAXQPT)
                                         phSolver/compressible/elmgmr.f:31-34
Linear tets only for now: genint.f:30-75 (symtet 1-/4-pt rule, Qwt*4/3),
genshp.f:34-37 (TetShapeAndDrv p=1, shgl/2).  In production the Fortran host
passes its own tables; these exist so the Python host mirror and the tests
can drive the C-ABI without Fortran.  Pinned against the reference's own C
generators in tests/test_tables.py (golden fixture tests/golden/tables_tet.npz).
"""
import numpy as np

MAXTOP, MAXSH, MAXQPT = 6, 32, 125
_A4, _B4 = 0.5854101966249685, 0.1381966011250150


def tet_points(rule):
    if rule == 1:
        return np.full((1, 4), 0.25), np.array([1.0])
    if rule == 2:
        pts = np.full((4, 4), _B4)
        np.fill_diagonal(pts, _A4)
        return pts, np.full(4, 0.25)
    raise NotImplementedError("tet quadrature rule %d" % rule)


def tri_points(rule):
    """symtri 1/3-pt rules (phSolver/common/symtri.c)."""
    if rule == 1:
        return np.array([[0.333333333333333, 0.333333333333333, 0.333333333333333, 0.0]]), np.array([1.0])
    if rule == 2:
        a, b = 0.666666666666667, 0.166666666666667
        pts = np.array([[a, b, b, 0.0], [b, a, b, 0.0], [b, b, a, 0.0]])
        return pts, np.full(3, 0.333333333333333)
    raise NotImplementedError("tri quadrature rule %d" % rule)


def make_tables(rule=2, ruleb=2):
    nint = np.zeros(MAXTOP, dtype=np.int32)
    nintb = np.zeros(MAXTOP, dtype=np.int32)
    Qwt = np.zeros((MAXTOP, MAXQPT), order="F")
    Qwtb = np.zeros((MAXTOP, MAXQPT), order="F")
    shp = np.zeros((MAXTOP, MAXSH, MAXQPT), order="F")
    shgl = np.zeros((MAXTOP, 3, MAXSH, MAXQPT), order="F")
    shpb = np.zeros((MAXTOP, MAXSH, MAXQPT), order="F")
    shglb = np.zeros((MAXTOP, 3, MAXSH, MAXQPT), order="F")
    pts, w = tet_points(rule)
    n = len(w)
    nint[0] = n
    Qwt[0, :n] = (4.0 / 3.0) * w
    dN = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [-1.0, -1.0, -1.0]])
    for i in range(n):
        r, s, t = pts[i, 0], pts[i, 1], pts[i, 2]
        shp[0, :4, i] = [r, s, t, 1.0 - r - s - t]
        shgl[0, :, :4, i] = dN.T / 2.0
    # boundary faces of tets: genint.f:44-75 (symtri, Qwtb*2), genshpb.f
    ptsb, wb = tri_points(ruleb)
    nb = len(wb)
    nintb[0] = nb
    Qwtb[0, :nb] = 2.0 * wb
    for i in range(nb):
        r, s, t = ptsb[i, 0], ptsb[i, 1], ptsb[i, 2]   # Qptb(1,1:3,i) (genshpb.f:24)
        shpb[0, :4, i] = [r, s, t, 1.0 - r - s - t]
        shglb[0, :, :4, i] = dN.T / 2.0
    return dict(nint=nint, nintb=nintb, Qwt=Qwt, Qwtb=Qwtb, shp=shp, shgl=shgl,
                shpb=shpb, shglb=shglb)
