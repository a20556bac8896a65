"""Quadrature / shape-function tables in the reference's COMMON layout.

Qwt(MAXTOP,MAXQPT), nint(MAXTOP)        phSolver/common/common.h:92-96
shp(MAXTOP,MAXSH,MAXQPT), shgl(MAXTOP,3,MAXSH,MAXQPT)
                                         phSolver/compressible/elmgmr.f:31-34
Linear tets: genint.f:30-75 (symtet 1-/4-pt rule, Qwt*4/3), genshp.f:34-37
(TetShapeAndDrv p=1, shgl/2); linear hexes (topology 2): genint.f:105-147
(symhex 8-pt rule), genshp.f:39-45 (HexShapeAndDrv, newshape.cc:420-500);
linear wedges (topology 3): genint.f:294-318 (symwdg 6-pt rule), genshp.f:47-53
(WedgeShapeAndDrv, newshape.cc:706-768).  In production the Fortran host
passes its own tables; these exist so the Python host mirror and the tests
can drive the C-ABI without Fortran.  Boundary faces: triangles of tets
(genint.f:44-75), quadrilaterals of hexes, triangles / quadrilaterals of wedges (genint.f:116-141,303-343,
genshpb.f).  Pinned against the reference's own C generators and against genint.f / genshp.f / genshpb.f
executed by f77np in tests/test_tables.py (tests/golden/tables_ref.npz, tables_f77.npz).
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


_G2 = 0.577350269189626           # symhex.c / symwdg.c Qp21, Qp22
_P23, _P24 = 0.166666666666667, 0.666666666666667


def hex_points(rule):
    """symhex 8-pt rule (phSolver/common/symhex.c rstw8/twt8)."""
    if rule != 2:
        raise NotImplementedError("hex quadrature rule %d" % rule)
    pts = np.array([[sx * _G2, sy * _G2, sz * _G2, 0.0]
                    for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)])
    return pts, np.full(8, 1.0)


def wedge_points(rule):
    """symwdg 6-pt rule (phSolver/common/symwdg.c rstw6/twt6)."""
    if rule != 2:
        raise NotImplementedError("wedge quadrature rule %d" % rule)
    tri = [(_P23, _P23), (_P24, _P23), (_P23, _P24)]
    pts = np.array([[r, s, z * _G2, 0.0] for z in (-1, 1) for r, s in tri])
    return pts, np.full(6, 0.666666666666667)


_PW1, _PW2 = 0.211324865405187, 0.788675134594813     # symquadw.c Pp21, Pp22


def quad_points(rule):
    """boundary quadrilateral of a hex: symquad 4-pt rule on the zeta = -1 face (phSolver/common/symquad.c)."""
    if rule != 2:
        raise NotImplementedError("quad quadrature rule %d" % rule)
    pts = np.array([[sx * _G2, sy * _G2, -1.0, 0.0] for sy in (-1, 1) for sx in (-1, 1)])
    return pts, np.full(4, 1.0)


def wedge_tri_points(rule):
    """boundary triangle of a wedge (genint.f:320-331): the symtri points rotated by one (the last point comes
    first), third coordinate zeta = -1, weights as symtri gives them (not doubled)."""
    pts, w = tri_points(rule)
    n = len(w)
    out = np.zeros((n, 4))
    out[1:, :2] = pts[:n - 1, :2]
    out[0, :2] = pts[n - 1, :2]
    out[:, 2:] = -1.0
    return out, w


def wedge_quad_points(rule):
    """boundary quadrilateral of a wedge: symquadw 4-pt rule on the s = 0 face (phSolver/common/symquadw.c)."""
    if rule != 2:
        raise NotImplementedError("wedge quad quadrature rule %d" % rule)
    pts = np.array([[r, 0.0, z * _G2, 0.0] for z in (-1, 1) for r in (_PW1, _PW2)])
    return pts, np.full(4, 1.0)


def hex_shape(xi, eta, zeta):
    """HexShapeAndDrv p=1 (phSolver/common/newshape.cc:420-500), same operation order."""
    xim, etam, zetam = 1 - xi, 1 - eta, 1 - zeta
    xip, etap, zetap = 1 + xi, 1 + eta, 1 + zeta
    N = np.array([0.125 * xim * etam * zetam, 0.125 * xip * etam * zetam, 0.125 * xip * etap * zetam,
                  0.125 * xim * etap * zetam, 0.125 * xim * etam * zetap, 0.125 * xip * etam * zetap,
                  0.125 * xip * etap * zetap, 0.125 * xim * etap * zetap])
    dN = np.array([
        [-0.125 * etam * zetam, -0.125 * xim * zetam, -0.125 * xim * etam],
        [0.125 * etam * zetam, -0.125 * xip * zetam, -0.125 * xip * etam],
        [0.125 * etap * zetam, 0.125 * xip * zetam, -0.125 * xip * etap],
        [-0.125 * etap * zetam, 0.125 * xim * zetam, -0.125 * xim * etap],
        [-0.125 * etam * zetap, -0.125 * xim * zetap, 0.125 * xim * etam],
        [0.125 * etam * zetap, -0.125 * xip * zetap, 0.125 * xip * etam],
        [0.125 * etap * zetap, 0.125 * xip * zetap, 0.125 * xip * etap],
        [-0.125 * etap * zetap, 0.125 * xim * zetap, 0.125 * xim * etap]])
    return N, dN


def wedge_shape(r, s, zeta):
    """WedgeShapeAndDrv p=1 (phSolver/common/newshape.cc:706-768)."""
    p0, p1, p2, p3 = 1.0 - r - s, r, s, zeta
    N = np.array([0.5 * p0 * (1.0 - p3), 0.5 * p1 * (1.0 - p3), 0.5 * p2 * (1.0 - p3),
                  0.5 * p0 * (1.0 + p3), 0.5 * p1 * (1.0 + p3), 0.5 * p2 * (1.0 + p3)])
    dN = np.array([
        [-0.25 * (1.0 - p3), -0.25 * (1.0 - p3), -0.5 * p0],
        [0.25 * (1.0 - p3), 0.0, -0.5 * p1],
        [0.0, 0.25 * (1.0 - p3), -0.5 * p2],
        [-0.25 * (1.0 + p3), -0.25 * (1.0 + p3), 0.5 * p0],
        [0.25 * (1.0 + p3), 0.0, 0.5 * p1],
        [0.0, 0.25 * (1.0 + p3), 0.5 * p2]])
    return N, dN


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
    if rule == 2:
        # hexes (lcsyst 2) and wedges (lcsyst 3); the reference defines nint(3) only for rules 2..4
        for top, points, shape, nsh in ((1, hex_points, hex_shape, 8), (2, wedge_points, wedge_shape, 6)):
            pts, w = points(rule)
            n = len(w)
            nint[top] = n
            Qwt[top, :n] = w
            for i in range(n):
                N, d = shape(pts[i, 0], pts[i, 1], pts[i, 2])
                shp[top, :nsh, i] = N
                shgl[top, :, :nsh, i] = d.T
    # boundary faces of tets: genint.f:44-75 (symtri, Qwtb*2), genshpb.f
    ptsb, wb = tri_points(ruleb)
    nb = len(wb)
    nintb[0] = nb
    Qwtb[0, :nb] = 2.0 * wb
    for i in range(nb):
        r, s, t = ptsb[i, 0], ptsb[i, 1], ptsb[i, 2]   # Qptb(1,1:3,i) (genshpb.f:24)
        shpb[0, :4, i] = [r, s, t, 1.0 - r - s - t]
        shglb[0, :, :4, i] = dN.T / 2.0
    if ruleb == 2:
        # boundary faces of hexes (lcsyst 2), wedges with a triangular (3) or quadrilateral (4) boundary face:
        # genint.f:116-141,303-343 (no weight rescaling), genshpb.f:36-56 (volume shape functions at the face points)
        for top, points, shape, nsh in ((1, quad_points, hex_shape, 8), (2, wedge_tri_points, wedge_shape, 6),
                                        (3, wedge_quad_points, wedge_shape, 6)):
            pts, w = points(ruleb)
            n = len(w)
            nintb[top] = n
            Qwtb[top, :n] = w
            for i in range(n):
                N, d = shape(pts[i, 0], pts[i, 1], pts[i, 2])
                shpb[top, :nsh, i] = N
                shglb[top, :, :nsh, i] = d.T
    return dict(nint=nint, nintb=nintb, Qwt=Qwt, Qwtb=Qwtb, shp=shp, shgl=shgl,
                shpb=shpb, shglb=shglb)
