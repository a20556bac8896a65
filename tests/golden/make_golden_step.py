"""Golden vectors of one whole time step from the reference's own Fortran (f77np): the flow sequence of
compressible/itrdrv.f:398-652 with `Step Construction 0 1 0 1 ...` --

    itrSetup (itrPC.f:1-30)                       -> almi, alfi, gami, Dtgl from rhoinf / Delt
    itrPredict, itrBC                              (itrdrv.f:398-399)
    nitr x [ lhs = 1 - min(1, mod(ifuncs-1, LHSupd)); SolGMRe (-> rstat); itrCorrect; itrBC ]   (:435-598)
    itrUpdate, itrBC(yold, acold)                  (:651-652)

on seeded cases with every essential-BC code.  Writes tests/golden/f77_step_*.npz (y, ac, yold, acold after the
step, per-iteration iKs and rstat's totres through resfrt/jtotrs inputs).

    python tests/golden/make_golden_step.py [--check]
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import make_golden_f77 as mg  # noqa: E402
from golden_cases import input_digest  # noqa: E402

# name -> (make_case args, kwargs, step options)
# smooth channel states (phasta_b200.mesh.make_smooth_state) so that the Newton iteration converges, as in
# tests/test_timestep.py; boundary elements with natural BCs on the second case
CASES = {
    "be_channel": ((5, 4, 3), dict(boundary=False, etol=1e-3), dict(rhoinf=-1.0, nitr=2, ipred=1, LHSupd=1)),
    "genalpha_lhsupd2": ((4, 4, 3), dict(boundary=True, etol=1e-3), dict(rhoinf=0.5, nitr=3, ipred=1, LHSupd=2)),
}


def build_case(name):
    from phasta_b200 import SolverParams, make_box, make_smooth_state, make_tables
    (nx, ny, nz), kw, opt = CASES[name]
    kw = dict(kw)
    boundary = kw.pop("boundary")
    parts = make_box(nx, ny, nz, bc="channel", boundary=boundary, natural="mixed" if boundary else "none", ibksiz=32)
    states = [make_smooth_state(p) for p in parts]
    return (SolverParams(ibksiz=32, **kw), make_tables(2, 2), parts, states), dict(opt)


def run(prog, case, opt):
    params, tables, parts, states = case
    mp = parts[0]
    nshape = max(int(b.shape[1]) for b in mp.mien)
    nedof = 5 * nshape
    mg.set_commons(prog, params, tables, mp, nedof)
    mg.set_pointer_data(prog, mp, tables)
    G = prog.G
    nshg, numel, K = mp.nshg, mp.numel, int(params.Kspace)
    y0, ac0 = (mg.F(a) for a in states[0])
    yold, acold = y0.copy(order="F"), ac0.copy(order="F")
    y, ac = y0.copy(order="F"), ac0.copy(order="F")
    x, BC = mg.F(mp.x), mg.F(mp.BC)
    iBC = np.array(mp.iBC, dtype=np.int64)
    iper = np.array(mp.iper, dtype=np.int64)
    ilwork = np.zeros(1, dtype=np.int64)
    shp, shgl, shpb, shglb = mg.full_tables(tables)
    # itrSetup: the reference derives the time-integration scalars itself
    G.update(itseq=1, ipred=int(opt["ipred"]), lctime=0, irscale=-1, ntotgm=0, iter=0, lstep=0, istep=0,
             etol=float(params.etol), ylimit=G["ylimit"])
    G["rhoinf"][0] = float(opt["rhoinf"])
    G["delt"][0] = 1.0 / float(params.Dtgl)
    G["cflfl"][0] = 1.0
    G["cflsl"][0] = 1.0
    G["resfrt"][...] = 0.0
    prog.call("itrsetup", y, acold)
    scal = dict(almi=float(G["almi"]), alfi=float(G["alfi"]), gami=float(G["gami"]), Dtgl=float(G["dtgl"]))
    prog.call("itrpredict", yold, acold, y, ac)
    prog.call("itrbc", y, ac, iBC, BC, iper, ilwork)
    res = np.zeros((nshg, 5), order="F")
    BDiag = np.zeros((nshg, 5, 5), order="F")
    EGmass = np.zeros((numel, nedof, nedof), order="F")
    HBrg = np.zeros((K + 1, K), order="F")
    eBrg, yBrg, Rcos, Rsin = (np.zeros(K + 1) for _ in range(4))
    solinc = np.zeros((nshg, 5), order="F")
    rerr = np.zeros((nshg, 10), order="F")
    ifuncs, iks, lhss, unpre = 0, [], [], []
    # rstat (called by SolGMRe, solgmr.f:352) needs nshgt; its totres is local, so keep the norms it is built from
    G["nshgt"] = nshg
    for it in range(1, int(opt["nitr"]) + 1):
        G["iter"] = it
        ifuncs += 1
        lhs = 1 - min(1, (ifuncs - 1) % int(opt["LHSupd"]))
        G["lhs"], G["iprec"] = lhs, lhs
        G["force"][...] = 0.0
        G["hflux"] = 0.0
        prog.call("solgmre", y, ac, yold, acold, x, iBC, BC, EGmass, res, BDiag, HBrg, eBrg, yBrg, Rcos, Rsin, iper,
                  ilwork, shp, shgl, shpb, shglb, solinc, rerr)
        iks.append(int(G["iks"]))
        lhss.append(lhs)
        unpre.append(float(np.sqrt(np.sum(res ** 2)) / nshg))      # totres(1) of rstat.f:94-103 (preconditioned res)
        prog.call("itrcorrect", y, ac, yold, acold, solinc)
        prog.call("itrbc", y, ac, iBC, BC, iper, ilwork)
    prog.call("itrupdate", yold, acold, y, ac)
    prog.call("itrbc", yold, acold, iBC, BC, iper, ilwork)
    return dict(y=y, ac=ac, yold=yold, acold=acold, iKs=np.array(iks), lhs=np.array(lhss), totres1=np.array(unpre),
                ntotGM=int(G["ntotgm"]), Dy_last=solinc, **{k: np.array(v) for k, v in scal.items()})


def oracle_step(case, opt, scal):
    """the oracle's orc_timestep on the same case with the scalars itrSetup produced"""
    from common import make_oracle
    params, tables, parts, states = case
    for k in ("almi", "alfi", "gami", "Dtgl"):
        setattr(params, k, float(scal[k]))
    o = make_oracle((params, tables, parts, states))
    st = o.TimeStep(nitr=int(opt["nitr"]), ipred=int(opt["ipred"]), LHSupd=int(opt["LHSupd"]))
    p = o.parts[0]
    return o, st, p.keep["y"], p.keep["ac"], o.yold[0], o.acold[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    stubs_extra = ("genscale", "asbwmod")
    prog = mg.make_program()
    for s in stubs_extra:
        prog.stubs[s] = mg._noop
    ok = True
    for name in CASES:
        case, opt = build_case(name)
        r = run(prog, case, opt)
        r["digest"] = input_digest(case)
        np.savez_compressed(os.path.join(HERE, "f77_step_%s.npz" % name), **r)
        print("%s: almi %.6f alfi %.6f gami %.6f  iKs %s lhs %s" % (name, r["almi"], r["alfi"], r["gami"], r["iKs"], r["lhs"]))
        if args.check:
            o, st, y, ac, yold, acold = oracle_step(case, opt, r)
            for nm, a, b in (("y", y, r["y"]), ("ac", ac, r["ac"]), ("yold", yold, r["yold"]), ("acold", acold, r["acold"])):
                d = np.linalg.norm(a - b) / np.linalg.norm(b)
                print("   %-6s rel-L2 %.3e" % (nm, d))
                ok &= d < 1e-9
            print("   iKs oracle", st[:, 2].astype(int), "reference", r["iKs"])
            ok &= np.array_equal(st[:, 2].astype(int), r["iKs"])
    case, yb, acb = run_itrbc_allcodes(prog)
    if args.check:
        from common import make_oracle
        o = make_oracle(case)
        o.itrBC()
        same = np.array_equal(o.parts[0].keep["y"], yb) and np.array_equal(o.parts[0].keep["ac"], acb)
        print("itrbc_allcodes: oracle bit-equal", same)
        ok &= same
    if args.check:
        print("ALL OK" if ok else "MISMATCH")


def run_itrbc_allcodes(prog):
    """itrBC (itrbc.f) on a random state with every essential-BC code (velocity codes 1..7, density,
    pressure, temperature, periodic slaves) -> tests/golden/f77_itrbc_allcodes.npz"""
    from common import make_case
    case = make_case(4, 4, 3, bc="allcodes", ibksiz=50)
    params, tables, parts, states = case
    mp = parts[0]
    mg.set_commons(prog, params, tables, mp, 20)
    y, ac = (mg.F(a) for a in states[0])
    iBC = np.array(mp.iBC, dtype=np.int64)
    prog.G["ires"] = 1
    prog.call("itrbc", y, ac, iBC, mg.F(mp.BC), np.array(mp.iper, dtype=np.int64), np.zeros(1, dtype=np.int64))
    np.savez_compressed(os.path.join(HERE, "f77_itrbc_allcodes.npz"), y=y, ac=ac, digest=input_digest(case))
    return case, y, ac


if __name__ == "__main__":
    main()
