"""Quadrature / shape tables: bit-exact against the reference's own C
generators (golden fixture made by tests/golden/make_golden.py; and live
against oracle/_ref when it was built in this container)."""
import ctypes as C
import os

import numpy as np
import pytest

from phasta_b200.tables import make_tables, tet_points, tri_points, MAXTOP, MAXSH, MAXQPT

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "tables_ref.npz"))


@pytest.mark.parametrize("rule,n", [(1, 1), (2, 4)])
def test_tet_rule_bit_exact(rule, n):
    pts, w = tet_points(rule)
    assert np.array_equal(pts, GOLD["tet%d_pt" % n])
    assert np.array_equal(w, GOLD["tet%d_wt" % n])
    T = make_tables(rule, 2)
    assert T["nint"][0] == n
    # genint.f:74 Qwt*4/3 ; genshp.f:34-37 shgl/2
    assert np.array_equal(T["Qwt"][0, :n], (4.0 / 3.0) * GOLD["tet%d_wt" % n])
    assert np.array_equal(T["shp"][0, :4, :n], GOLD["tet%d_N" % n].T)
    for i in range(n):
        assert np.array_equal(T["shgl"][0, :, :4, i], GOLD["tet%d_dN" % n][i].T / 2.0)
    assert T["shp"].shape == (MAXTOP, MAXSH, MAXQPT) and T["shgl"].shape == (MAXTOP, 3, MAXSH, MAXQPT)


@pytest.mark.parametrize("rule,n", [(1, 1), (2, 3)])
def test_tri_rule_bit_exact(rule, n):
    pts, w = tri_points(rule)
    assert np.array_equal(pts, GOLD["tri%d_pt" % n])
    assert np.array_equal(w, GOLD["tri%d_wt" % n])


@pytest.mark.parametrize("name,top,n,nsh", [("hex", 1, 8, 8), ("wdg", 2, 6, 6)])
def test_hex_wedge_tables_bit_exact(name, top, n, nsh):
    """genint.f:105-147,294-318 + genshp.f:39-53: no weight / derivative rescaling for hexes and wedges."""
    T = make_tables(2, 2)
    assert T["nint"][top] == n
    assert np.array_equal(T["Qwt"][top, :n], GOLD["%s%d_wt" % (name, n)])
    assert np.array_equal(T["shp"][top, :nsh, :n], GOLD["%s%d_N" % (name, n)].T)
    for i in range(n):
        assert np.array_equal(T["shgl"][top, :, :nsh, i], GOLD["%s%d_dN" % (name, n)][i].T)


def test_all_tables_match_genint_genshp_genshpb_executed_by_f77np():
    """the whole COMMON tables (rule 2, interior lcsyst 1..3, boundary lcsyst 1..4) against the reference's
    genint.f / genshp.f / genshpb.f run on top of its own C generators: bit for bit"""
    Z = np.load(os.path.join(HERE, "golden", "tables_f77.npz"))
    T = make_tables(2, 2)
    assert np.array_equal(T["nint"][:3], Z["nint"][:3]) and np.array_equal(T["nintb"][:4], Z["nintb"][:4])
    for k in ("Qwt", "shp", "shgl"):
        assert np.array_equal(T[k][:3], Z[k][:3]), k
    for k in ("Qwtb", "shpb", "shglb"):
        assert np.array_equal(T[k][:4], Z[k][:4]), k
    # the wedge's triangular face: rotated points, weights NOT doubled (e3bvar.f:147 uses 1 - Qwtb)
    assert np.array_equal(Z["Qptb"][2, :3, 0], [0.166666666666667, 0.166666666666667, -1.0])
    assert np.array_equal(Z["Qwtb"][2, :3], np.full(3, 0.333333333333333))


def test_oracle_tables_match_python_tables():
    from oracle.oracle_py import lib
    L = lib()
    for rule in (1, 2):
        T = make_tables(rule, 2)
        nint = (C.c_int * MAXTOP)()
        Qwt = np.zeros((MAXTOP, MAXQPT), order="F")
        shp = np.zeros((MAXTOP, MAXSH, MAXQPT), order="F")
        shgl = np.zeros((MAXTOP, 3, MAXSH, MAXQPT), order="F")
        L.orc_tet_tables(rule, nint, Qwt.ctypes.data_as(C.c_void_p), shp.ctypes.data_as(C.c_void_p),
                         shgl.ctypes.data_as(C.c_void_p))
        assert nint[0] == T["nint"][0]
        assert np.array_equal(Qwt[0], T["Qwt"][0]) and np.array_equal(shp[0], T["shp"][0])
        assert np.array_equal(shgl[0], T["shgl"][0])


def test_live_reference_generators_if_present():
    so = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_tables.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built here (reference sources absent)")
    L = C.CDLL(so)
    pt = np.zeros((4, 4))
    wt = np.zeros(4)
    err = C.c_int(0)
    L.symtet_(C.byref(C.c_int(4)), pt.ctypes.data_as(C.c_void_p), wt.ctypes.data_as(C.c_void_p), C.byref(err))
    assert np.array_equal(pt, GOLD["tet4_pt"]) and np.array_equal(wt, GOLD["tet4_wt"])
