"""Golden vectors for the incompressible assembly + lesSparse products from the reference's own Fortran
(phSolver/incompressible/elmgmr.f ElmGMR and everything it calls, common/fillsparse.f fillsparseI,
incompressible/lesSparse.f fLesSparseAp*), executed UNMODIFIED by the f77np interpreter.

    python tests/golden/make_golden_incomp.py [--check]     -> tests/golden/f77_incomp_*.npz
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from f77np import Program, scan_functions  # noqa: E402
from make_golden_f77 import F, set_commons, set_pointer_data, full_tables  # noqa: E402

REF = "/root/reference/phSolver"
INC = ["elmgmr.f", "asigmr.f", "asiq.f", "e3.f", "e3ivar.f", "e3stab.f", "e3res.f", "e3lhs.f", "e3q.f", "e3qvar.f",
       "getdiff.f", "bc3lhs.f", "bc3res.f", "bc3per.f", "lesSparse.f", "asbmfg.f", "e3b.f", "e3bvar.f"]
COMMON = ["clear.f", "mpitools.f", "e3metric.f", "local.f", "localy.f", "hierarchic.f", "qpbc.f", "fillsparse.f", "genadj.f",
          "asadj.f"]

# name -> (make_case args, kwargs, IncompParams overrides)
CASES = {
    "tet_allbc": ((4, 4, 3), dict(bc="allcodes", ibksiz=50), dict()),
    "tet_conservative_nodiff": ((3, 2, 2), dict(bc="channel", ibksiz=64), dict(iconvflow=1, idiff=0, matflg5=1,
                                                                              bf=(0.3, -9.81, 0.1), flmpl=0.25,
                                                                              flmpr=0.5)),
    "tet_genalpha": ((3, 3, 3), dict(bc="allcodes", ibksiz=5), dict(rhoinf=0.5, Delt=2.0e-3, rho=1.2, rmu=1.8e-5)),
    "hex_allbc": ((4, 3, 3), dict(bc="allcodes", topo="hex", ibksiz=8), dict()),
    "mixed_topo": ((2, 4, 2), dict(bc="channel", topo="mixed", ibksiz=16), dict(iconvflow=1)),
    "resonly": ((3, 3, 3), dict(bc="allcodes", ibksiz=16), dict(lhs=0)),
    # the boundary integral (incompressible/asbmfg.f, e3b.f, e3bvar.f; rigid walls): tets, then hexes, then a mixed
    # mesh with triangular and quadrilateral wedge faces; every natural-BC code; flux through surfaces 1..3 recorded
    "tet_bnd": ((3, 2, 2), dict(bc="channel", ibksiz=16, boundary=True, natural="mixed"),
                dict(itwmod=1, surfaces=(1, 2, 3))),
    "hex_bnd": ((3, 2, 2), dict(bc="channel", topo="hex", ibksiz=8, boundary=True, natural="mixed", periodic_z=False),
                dict(iconvflow=1, itwmod=1, surfaces=(2, 3, 5), iviscflux=0)),
    "mixed_bnd": ((2, 4, 2), dict(bc="channel", topo="mixed", ibksiz=16, boundary=True, natural="mixed",
                                  periodic_z=False), dict(itwmod=-1, surfaces=(1, 4, 6))),
}


def build_case(name):
    from common import make_case
    from phasta_b200 import IncompParams
    a, kw, over = CASES[name]
    over = dict(over)
    rhoinf = over.pop("rhoinf", None)
    ip = IncompParams(**over)
    if rhoinf is not None:
        ip.with_rhoinf(rhoinf)
    return make_case(*a, **kw), ip


def _noop(prog, *a):
    return None


def make_program():
    stubs = {n: _noop for n in ("timer", "error", "mpi_barrier", "timeseries", "getsgn", "rotabc", "commu",
                                "mpi_allreduce", "flush", "mpi_abort", "elmpvsq", "lmassadd", "asiqgradv",
                                "solvegradv", "e3stsres")}
    modules = dict(exts=False, freq=1, rls=None, stsresflg=0, have_local_mass=0, gmass=None, stsvec=None,
                   nresdims=11, nabi=None, d2wall=None)
    prog = Program([os.path.join(REF, "common")], modules=modules, stubs=stubs)
    for f in INC:
        scan_functions(os.path.join(REF, "incompressible", f))
    for f in COMMON:
        scan_functions(os.path.join(REF, "common", f))
    for f in INC:
        prog.load(os.path.join(REF, "incompressible", f))
    for f in COMMON:
        prog.load(os.path.join(REF, "common", f))
    return prog


def run(prog, case, ip, nnz=35):
    params, tables, parts, states = case
    mp = parts[0]
    y, ac = (F(a) for a in states[0])
    nshape = max(int(b.shape[1]) for b in mp.mien)
    set_commons(prog, params, tables, mp, 5 * nshape)
    set_pointer_data(prog, mp, tables)
    G = prog.G
    # what input.f / input_fform.cc / itrSetup would have set for an incompressible run
    G.update(nflow=4, idflx=9 if ip.idiff >= 1 else 0, idiff=int(ip.idiff), itau=int(ip.itau), lhs=int(ip.lhs),
             iconvflow=int(ip.iconvflow), ipord=int(ip.ipord), flmpl=float(ip.flmpl), flmpr=float(ip.flmpr),
             almi=float(ip.almi), alfi=float(ip.alfi), gami=float(ip.gami), dtgl=float(ip.Dtgl),
             dtsfct=float(ip.dtsfct), taucfct=float(ip.taucfct), itseq=1, iles=0, irans=0, ilset=0, isurf=0,
             ierrcalc=0, icomputevort=0, ipvsq=0, iabc=0, intpres=0, nnz=nnz, iter=1, nitr=1, numpe=1,
             ideformwall=0, iviscflux=int(ip.iviscflux), itwmod=int(ip.itwmod), navier=1)
    G["nsrflist"][...] = 0
    for k in ip.surfaces:           # the force / flux list (input.config "Surface ID's for Integrated Mass")
        G["nsrflist"][int(k)] = 1
    G["flxid"][...] = 0.0
    G["force"][...] = 0.0
    G["delt"][0] = float(ip.Delt)
    G["impl"][0] = 10
    G["datmat"][...] = 0.0
    G["matflg"][...] = 0
    G["matflg"][0, 0] = -1
    G["datmat"][0, 0, 0] = ip.rho
    G["datmat"][0, 1, 0] = ip.rmu
    G["matflg"][4, 0] = int(ip.matflg5)
    G["datmat"][0:3, 4, 0] = ip.bf
    nshg = mp.nshg
    colm = np.zeros(nshg + 1, dtype=np.int64)
    rowp = np.zeros(nshg * nnz, dtype=np.int64)
    L = prog.call("genadj", colm, rowp, 0)
    nnz_tot = int(L["icnt"])
    G["nnz_tot"] = nnz_tot
    x, BC = F(mp.x), F(mp.BC)
    iBC = np.array(mp.iBC, dtype=np.int64)
    iper = np.array(mp.iper, dtype=np.int64)
    ilwork = np.zeros(1, dtype=np.int64)
    shp, shgl, shpb, shglb = full_tables(tables)
    u = np.zeros((nshg, 3), order="F")
    res = np.full((nshg, 4), np.nan, order="F")
    lhsK = np.full((9, nnz_tot), np.nan if ip.lhs else 0.0, order="F")
    lhsP = np.full((4, nnz_tot), np.nan if ip.lhs else 0.0, order="F")
    rerr = np.zeros((nshg, 10), order="F")
    GradV = np.zeros((nshg, 9), order="F")
    prog.call("elmgmr", u, y, ac, x, shp, shgl, iBC, BC, shpb, shglb, res, iper, ilwork, rowp, colm, lhsK, lhsP,
              rerr, GradV)
    out = dict(res=res, colm=colm, rowp=rowp[:nnz_tot].copy(), nnz_tot=nnz_tot)
    if mp.nelblb:
        out["flxID"] = np.array(G["flxid"][:5, :7], order="F")
        out["Force"] = np.array(G["force"])
    if ip.lhs:
        out.update(lhsK=lhsK, lhsP=lhsP)
        rng = np.random.default_rng(4242)
        p4 = F(rng.standard_normal((nshg, 4)))
        row = rowp[:nnz_tot].copy()
        q3 = np.full((nshg, 3), np.nan, order="F")
        prog.call("flessparseapg", colm, row, lhsP, p4[:, 3].copy(), q3, nshg, nnz_tot)
        out["ap_in"], out["apG"] = p4, q3.copy(order="F")
        prog.call("flessparseapkg", colm, row, lhsK, lhsP, p4, q3, nshg, nnz_tot)
        out["apKG"] = q3.copy(order="F")
        q1 = np.full(nshg, np.nan)
        prog.call("flessparseapngt", colm, row, lhsP, F(p4[:, :3]), q1, nshg, nnz_tot)
        out["apNGt"] = q1.copy()
        prog.call("flessparseapngtc", colm, row, lhsP, p4, q1, nshg, nnz_tot)
        out["apNGtC"] = q1.copy()
        q4 = np.full((nshg, 4), np.nan, order="F")
        prog.call("flessparseapfull", colm, row, lhsK, lhsP, p4, q4, nshg, nnz_tot)
        out["apFull"] = q4
    return out


def check(name, a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    d = np.abs(a - b).max()
    s = np.abs(b).max()
    ok = d <= tol * max(s, 1e-300)
    print("   %-10s max|diff| %.3e  max|ref| %.3e  bit-equal %s  %s" % (name, d, s, np.array_equal(a, b),
                                                                       "ok" if ok else "MISMATCH"))
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    from common import make_oracle
    from golden_cases import input_digest
    prog = make_program()
    allok = True
    for name in CASES:
        if args.only and name not in args.only.split(","):
            continue
        case, ip = build_case(name)
        t0 = time.time()
        r = run(prog, case, ip)
        print("incomp/%s via f77np: %.1f s" % (name, time.time() - t0))
        r["digest"] = input_digest(case)
        np.savez_compressed(os.path.join(HERE, "f77_incomp_%s.npz" % name), **r)
        if args.check:
            o = make_oracle(case)
            o.genadj()
            o.IncElmGMR(ip)
            p = o.parts[0]
            allok &= check("res", p.res4, r["res"], 1e-13)
            if "flxID" in r:
                allok &= check("flxID", p.aerfrc[4:4 + 70].reshape((10, 7), order="F")[:5], r["flxID"], 1e-13)
                allok &= check("Force", p.aerfrc[:3], r["Force"], 1e-13)
            if ip.lhs:
                allok &= check("lhsK", p.lhsK9, r["lhsK"], 1e-13)
                allok &= check("lhsP", p.lhsP4, r["lhsP"], 1e-13)
                pin = r["ap_in"]
                allok &= check("ApG", o.LesAp("G", pin[:, 3].copy()), r["apG"], 1e-14)
                allok &= check("ApKG", o.LesAp("KG", pin), r["apKG"], 1e-14)
                allok &= check("ApNGt", o.LesAp("NGt", pin[:, :3]), r["apNGt"], 1e-14)
                allok &= check("ApNGtC", o.LesAp("NGtC", pin), r["apNGtC"], 1e-14)
                allok &= check("ApFull", o.LesAp("Full", pin), r["apFull"], 1e-14)
    if args.check:
        print("ALL OK" if allok else "MISMATCH")


if __name__ == "__main__":
    main()
