"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see
oracle/phasta_oracle.h).  Importable only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MAXTOP, MAXQPT = 6, 125


class OrcCommon(C.Structure):
    _fields_ = [
        *[(n, C.c_int) for n in (
            "nshg", "numnp", "numel", "numelb", "nflow", "ndof", "ndofBC", "nshape", "nedof",
            "nelblk", "nelblb", "nlwork", "numpe", "myrank",
            "ipord", "idiff", "itau", "iprec", "lhs", "ires", "iremoveStabTimeTerm",
            "EntropyPressure",
            "iDC", "Navier", "Kspace", "nGMRES", "minIters",
            "matflg2", "matflg3", "pad0")],
        *[(n, C.c_double) for n in (
            "Rgas", "gamma", "gamma1", "pr", "datmat121", "datmat221", "datmat321", "datmat131",
            "epsM", "dtsfct", "taucfct", "temper",
            "Dtgl", "almi", "alfi", "gami", "etol")],
        ("nint", C.c_int * MAXTOP), ("nintb", C.c_int * MAXTOP),
        ("Qwt", C.c_double * (MAXTOP * MAXQPT)), ("Qwtb", C.c_double * (MAXTOP * MAXQPT)),
    ]


_PTRS = ["lcblk", "ien", "ien_off", "lcblkb", "ienb", "ienb_off", "iBCB", "iBCB_off", "BCB",
         "BCB_off", "x", "iBC", "BC", "iper", "ilwork", "shp", "shgl", "shpb", "shglb",
         "y", "ac", "res", "rmes", "BDiag", "EGmass", "qres", "rmass", "Dy", "uBrg", "temp",
         "lhsK", "colm", "rowp", "aerfrc"]


class OrcPart(C.Structure):
    _fields_ = [("c", OrcCommon)] + [(n, C.c_void_p) for n in _PTRS]


class OrcIncomp(C.Structure):
    """oracle_incomp.h orc_incomp"""
    _fields_ = [*[(n, C.c_int) for n in ("iconvflow", "itau", "idiff", "ipord", "lhs", "matflg5")],
                ("rho", C.c_double), ("rmu", C.c_double), ("bf", C.c_double * 3),
                *[(n, C.c_double) for n in ("flmpl", "flmpr", "Delt", "Dtgl", "almi", "alfi", "gami",
                                            "dtsfct", "taucfct")],
                ("iviscflux", C.c_int), ("itwmod", C.c_int), ("nsrflist", C.POINTER(C.c_int))]

    @classmethod
    def from_params(cls, ip):
        s = cls()
        for n in ("iconvflow", "itau", "idiff", "ipord", "lhs", "matflg5"):
            setattr(s, n, int(getattr(ip, n)))
        for n in ("rho", "rmu", "flmpl", "flmpr", "Delt", "Dtgl", "almi", "alfi", "gami", "dtsfct", "taucfct"):
            setattr(s, n, float(getattr(ip, n)))
        for i in range(3):
            s.bf[i] = float(ip.bf[i])
        s.iviscflux, s.itwmod = int(ip.iviscflux), int(ip.itwmod)
        s._nsrf = (C.c_int * 1001)()            # nsrflist(0:MAXSURF), kept alive with the struct
        for k in ip.surfaces:
            s._nsrf[int(k)] = 1
        s.nsrflist = C.cast(s._nsrf, C.POINTER(C.c_int))
        return s


def build(force=False):
    so = os.path.join(_HERE, "libphasta_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libphasta_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/phSolver/common") and \
            not os.path.exists(os.path.join(_HERE, "_ref", "libref_tables.so")):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


_FAST = False


def use_fast_build(on=True):
    """bench.py's CPU baseline: load the -O3 -march=native build instead."""
    global _FAST, _LIB
    _FAST, _LIB = bool(on), None


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if _FAST:
            fast = os.path.join(_HERE, "libphasta_oracle_fast.so")
            try:
                subprocess.check_call(["make", "-B", "-C", _HERE, "libphasta_oracle_fast.so"],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            except Exception:
                pass
            if os.path.exists(fast):
                so = fast
        _LIB = C.CDLL(so)
        assert _LIB.orc_sizeof_part() == C.sizeof(OrcPart), "orc_part layout mismatch"
        assert _LIB.orc_sizeof_common() == C.sizeof(OrcCommon)
        _LIB.orc_sumgat.restype = C.c_double
        assert _LIB.orc_sizeof_incomp() == C.sizeof(OrcIncomp)
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OraclePart:
    """Owns every array of one part in the reference layouts."""

    def __init__(self, mp, params, tables, y, ac, nedof=None):
        self.mp, self.params, self.tables = mp, params, tables
        nshg, numel = mp.nshg, mp.numel
        nshape = max(int(b.shape[1]) for b in mp.mien)
        self.nedof = nedof or 5 * nshape
        self.keep = {}
        k = self.keep
        k["lcblk"] = np.asfortranarray(mp.lcblk, dtype=np.int32)
        k["ien"] = np.concatenate([np.asfortranarray(b, dtype=np.int32).ravel(order="F") for b in mp.mien])
        off = np.zeros(len(mp.mien), dtype=np.int64)
        off[1:] = np.cumsum([b.size for b in mp.mien])[:-1]
        k["ien_off"] = off
        if mp.nelblb:
            k["lcblkb"] = np.asfortranarray(mp.lcblkb, dtype=np.int32)
            for nm, lst, dt in (("ienb", mp.mienb, np.int32), ("iBCB", mp.miBCB, np.int32),
                                ("BCB", mp.mBCB, np.float64)):
                k[nm] = np.concatenate([np.asfortranarray(b, dtype=dt).ravel(order="F") for b in lst])
                o = np.zeros(len(lst), dtype=np.int64)
                o[1:] = np.cumsum([b.size for b in lst])[:-1]
                k[nm + "_off"] = o
        k["x"] = np.asfortranarray(mp.x, dtype=np.float64)
        k["iBC"] = np.ascontiguousarray(mp.iBC, dtype=np.int32)
        k["BC"] = np.asfortranarray(mp.BC, dtype=np.float64)
        k["iper"] = np.ascontiguousarray(mp.iper, dtype=np.int32)
        k["ilwork"] = np.ascontiguousarray(mp.ilwork, dtype=np.int32)
        for nm in ("shp", "shgl", "shpb", "shglb"):
            k[nm] = np.asfortranarray(tables[nm], dtype=np.float64)
        k["y"] = np.asfortranarray(y, dtype=np.float64).copy(order="F")
        k["ac"] = np.asfortranarray(ac, dtype=np.float64).copy(order="F")
        K = params.Kspace
        self.res = k["res"] = np.zeros((nshg, 5), order="F")
        self.rmes = k["rmes"] = np.zeros((nshg, 5), order="F")
        self.BDiag = k["BDiag"] = np.zeros((nshg, 5, 5), order="F")
        self.EGmass = k["EGmass"] = np.zeros((numel, self.nedof, self.nedof), order="F")
        self.qres = k["qres"] = np.zeros((nshg, 12), order="F")
        self.rmass = k["rmass"] = np.zeros(nshg)
        self.Dy = k["Dy"] = np.zeros((nshg, 5), order="F")
        self.uBrg = k["uBrg"] = np.zeros((nshg, 5, K + 1), order="F")
        k["temp"] = np.zeros((nshg, 5), order="F")
        self.aerfrc = k["aerfrc"] = np.zeros(4 + 10 * 1001)

    def fill(self, s: OrcPart):
        mp, P, T = self.mp, self.params, self.tables
        c = s.c
        c.nshg, c.numnp, c.numel = mp.nshg, mp.numnp, mp.numel
        c.numelb = int(sum(b.shape[0] for b in mp.mienb)) if mp.nelblb else 0
        c.nflow, c.ndof, c.ndofBC = 5, 5, 6
        c.nshape, c.nedof = self.nedof // 5, self.nedof
        c.nelblk, c.nelblb, c.nlwork = mp.nelblk, mp.nelblb, mp.nlwork
        c.numpe, c.myrank = mp.numpe, mp.rank
        for nm in ("ipord", "idiff", "itau", "iprec", "lhs", "iremoveStabTimeTerm", "EntropyPressure",
                   "iDC", "Navier", "Kspace", "nGMRES", "minIters", "matflg2", "matflg3"):
            setattr(c, nm, int(getattr(P, nm)))
        c.ires = 1
        for nm in ("Rgas", "gamma", "gamma1", "pr", "datmat121", "datmat221", "datmat321", "datmat131",
                   "epsM", "dtsfct", "taucfct", "temper", "Dtgl", "almi", "alfi", "gami", "etol"):
            setattr(c, nm, float(getattr(P, nm)))
        for i in range(MAXTOP):
            c.nint[i] = int(T["nint"][i])
            c.nintb[i] = int(T["nintb"][i])
        q = np.asfortranarray(T["Qwt"]).ravel(order="F")
        qb = np.asfortranarray(T["Qwtb"]).ravel(order="F")
        for i in range(MAXTOP * MAXQPT):
            c.Qwt[i] = q[i]
            c.Qwtb[i] = qb[i]
        for nm in _PTRS:
            setattr(s, nm, _ptr(self.keep.get(nm)))


class Oracle:
    """All parts of one case; methods mirror the reference routine names."""

    def __init__(self, parts, params, tables, states):
        self.L = lib()
        self.parts = [OraclePart(mp, params, tables, y, ac) for mp, (y, ac) in zip(parts, states)]
        self.n = len(self.parts)
        self.arr = (OrcPart * self.n)()
        for p, s in zip(self.parts, self.arr):
            p.fill(s)
        K = params.Kspace
        self.HBrg = np.zeros((K + 1, K), order="F")
        self.eBrg = np.zeros(K + 1)
        self.yBrg = np.zeros(K + 1)
        self.Rcos = np.zeros(K + 1)
        self.Rsin = np.zeros(K + 1)
        self.ntotGM = C.c_int(0)

    def set_flags(self, **kw):
        for s in self.arr:
            for k, v in kw.items():
                setattr(s.c, k, v)

    def _vecs(self, arrs):
        pp = (C.POINTER(C.c_double) * self.n)()
        for i, a in enumerate(arrs):
            pp[i] = a.ctypes.data_as(C.POINTER(C.c_double))
        return pp

    def ElmGMRe(self):
        self.L.orc_elmgmre(self.n, self.arr)

    def i3LU(self, code, vecs=None):
        for i, p in enumerate(self.parts):
            r = p.res if vecs is None else vecs[i]
            self.L.orc_i3lu(C.byref(self.arr[i].c), _ptr(p.BDiag), _ptr(r), code)

    def i3pre(self):
        self.L.orc_i3pre(self.n, self.arr)

    def Au1GMR(self, vecs):
        self.L.orc_au1gmr(self.n, self.arr, self._vecs(vecs))

    def bc3per(self, vecs):
        for i, v in enumerate(vecs):
            self.L.orc_bc3per(C.byref(self.arr[i]), _ptr(v), 5)

    def commu(self, vecs, n, code):
        self.L.orc_commu(self.n, self.arr, self._vecs(vecs), n, {"in": 0, "out": 1}[code])

    def sumgat(self, vecs, n):
        return self.L.orc_sumgat(self.n, self.arr, self._vecs(vecs), n)

    # ---- block-CSR flavour ------------------------------------------------
    def genadj(self, nnz=35):
        """genadj.f: fills colm/rowp of every part; returns [nnz_tot]."""
        out = []
        for i, p in enumerate(self.parts):
            nshg = p.mp.nshg
            p.colm = p.keep["colm"] = np.zeros(nshg + 1, dtype=np.int32)
            rowp = np.zeros(nnz * nshg, dtype=np.int32)
            ntot = self.L.orc_genadj(C.byref(self.arr[i]), nnz, _ptr(p.colm), _ptr(rowp))
            p.rowp = p.keep["rowp"] = rowp[:ntot].copy()
            p.lhsK = p.keep["lhsK"] = np.zeros((25, ntot), order="F")
            for nm in ("colm", "rowp", "lhsK"):
                setattr(self.arr[i], nm, _ptr(p.keep[nm]))
            out.append(ntot)
        return out

    def ElmGMRs(self):
        self.L.orc_elmgmrs(self.n, self.arr)

    def Spsi3pre(self):
        self.L.orc_spsi3pre(self.n, self.arr)

    def SparseAp(self, vecs):
        self.L.orc_sparseap(self.n, self.arr, self._vecs(vecs))

    def SolGMRs(self):
        iKs, lG = C.c_int(0), C.c_int(0)
        self.L.orc_solgmrs(self.n, self.arr, _ptr(self.HBrg), _ptr(self.eBrg), _ptr(self.yBrg),
                           _ptr(self.Rcos), _ptr(self.Rsin), C.byref(iKs), C.byref(lG),
                           C.byref(self.ntotGM))
        return iKs.value, lG.value

    # ---- incompressible assembly + lesSparse products (oracle_incomp.c) -----
    def IncElmGMR(self, ip, want_ebe=False):
        """incompressible ElmGMR on every part (genadj first): sets p.res4 (nshg,4), p.lhsK9 (9,nnz_tot),
        p.lhsP4 (4,nnz_tot) and, if asked, p.xKebe (numel,9,nshape,nshape) / p.xGoC (numel,4,..)."""
        s = OrcIncomp.from_params(ip)
        for p in self.parts:
            ntot = p.rowp.size
            nsm = self.arr[0].c.nshape
            p.res4 = np.zeros((p.mp.nshg, 4), order="F")
            p.lhsK9 = np.zeros((9, ntot), order="F")
            p.lhsP4 = np.zeros((4, ntot), order="F")
            if want_ebe:
                p.xKebe = np.zeros((p.mp.numel, 9, nsm, nsm), order="F")
                p.xGoC = np.zeros((p.mp.numel, 4, nsm, nsm), order="F")
        ebe = (self._vecs([p.xKebe for p in self.parts]), self._vecs([p.xGoC for p in self.parts])) \
            if want_ebe else (None, None)
        self.L.orc_inc_elmgmr(self.n, self.arr, C.byref(s), self._vecs([p.res4 for p in self.parts]),
                              self._vecs([p.lhsK9 for p in self.parts]), self._vecs([p.lhsP4 for p in self.parts]),
                              ebe[0], ebe[1])

    def IncBc3Res(self, res, part=0):
        """bc3Res of the incompressible code alone, in place on res (nshg,4)"""
        assert res.flags.f_contiguous and res.shape == (self.parts[part].mp.nshg, 4)
        self.L.orc_inc_bc3res(C.byref(self.arr[part]), _ptr(res))
        return res

    def LesAp(self, kind, pvec, part=0):
        """fLesSparseAp{G,KG,NGt,NGtC,Full} (lesSparse.f:204-492) on one part's lhsK9/lhsP4."""
        p = self.parts[part]
        n = p.mp.nshg
        v = np.asfortranarray(pvec, dtype=np.float64)
        col, row = _ptr(p.colm), _ptr(p.rowp)
        if kind == "G":
            q = np.zeros((n, 3), order="F")
            self.L.orc_les_apg(n, col, row, _ptr(p.lhsP4), _ptr(v), _ptr(q))
        elif kind == "KG":
            q = np.zeros((n, 3), order="F")
            self.L.orc_les_apkg(n, col, row, _ptr(p.lhsK9), _ptr(p.lhsP4), _ptr(v), _ptr(q))
        elif kind in ("NGt", "NGtC"):
            q = np.zeros(n)
            self.L.orc_les_apngt(n, col, row, _ptr(p.lhsP4), _ptr(v), _ptr(q), int(kind == "NGtC"))
        elif kind == "Full":
            q = np.zeros((n, 4), order="F")
            self.L.orc_les_apfull(n, col, row, _ptr(p.lhsK9), _ptr(p.lhsP4), _ptr(v), _ptr(q))
        else:
            raise ValueError(kind)
        return q

    # ---- Newton / time-step shell (oracle_step.c) --------------------------
    def _state(self):
        if not hasattr(self, "yold"):
            self.yold = [p.keep["y"].copy(order="F") for p in self.parts]
            self.acold = [p.keep["ac"].copy(order="F") for p in self.parts]
            self.ifuncs = C.c_int(0)
        return [p.keep["y"] for p in self.parts], [p.keep["ac"] for p in self.parts]

    def itrBC(self, ires=1):
        y, ac = self._state()
        self.L.orc_itrbc(self.n, self.arr, self._vecs(y), self._vecs(ac), ires)

    def rstat(self, nshgt):
        out = np.zeros(2)
        self.L.orc_rstat(self.n, self.arr, int(nshgt), _ptr(out))
        return out

    def TimeStep(self, nitr=2, ipred=1, sparse=False, LHSupd=1, nshgt=None):
        """One step of itrdrv.f's flow sequence; y/ac/yold/acold updated in place
        (yold/acold start as copies of the initial y/ac).  Returns stats (nitr,6)."""
        y, ac = self._state()
        if nshgt is None:
            nshgt = sum(p.mp.nshg for p in self.parts)
        stats = np.zeros((nitr, 6))
        self.L.orc_timestep(self.n, self.arr, self._vecs(y), self._vecs(ac), self._vecs(self.yold),
                            self._vecs(self.acold), int(ipred), int(nitr), int(bool(sparse)), int(LHSupd),
                            int(nshgt), C.byref(self.ifuncs), C.byref(self.ntotGM), _ptr(stats))
        return stats

    # ---- matrix-free flavour (oracle_mfg.c) ----------------------------------
    def ElmMFG(self):
        self.L.orc_elmmfg(self.n, self.arr)

    def ItrRes(self, yp, iabres=0):
        """itrres.f on part 0 (single part): returns the modified residual of yp."""
        yp = [np.asfortranarray(yp, dtype=np.float64).copy(order="F")]
        out = [np.zeros((self.parts[0].mp.nshg, 5), order="F")]
        self.L.orc_itrres(self.n, self.arr, self._vecs(yp), self._vecs(out), int(iabres))
        return out[0]

    def Au1MFG_once(self, u, eGMRES):
        """solmfg.f:97-135 set-up on the current ElmMFG outputs, then one Au1MFG."""
        for i, p in enumerate(self.parts):
            c = C.byref(self.arr[i].c)
            self.L.orc_i3lu(c, _ptr(p.BDiag), _ptr(p.res), 0)
            self.L.orc_i3lu(c, _ptr(p.BDiag), _ptr(p.res), 1)
            self.L.orc_i3lu(c, _ptr(p.BDiag), _ptr(p.rmes), 1)
        self.L.orc_mfg_begin(self.n, self.arr, C.c_double(eGMRES))
        v = [np.asfortranarray(u, dtype=np.float64).copy(order="F")]
        self.L.orc_au1mfg(self.n, self.arr, self._vecs(v))
        self.L.orc_mfg_end(self.n)
        return v[0]

    def SolMFG(self, eGMRES=0.0, iter=1, istep=0):
        iKs, lG, eG = C.c_int(0), C.c_int(0), C.c_double(eGMRES)
        self.L.orc_solmfg(self.n, self.arr, _ptr(self.HBrg), _ptr(self.eBrg), _ptr(self.yBrg),
                          _ptr(self.Rcos), _ptr(self.Rsin), C.byref(iKs), C.byref(lG),
                          C.byref(self.ntotGM), C.byref(eG), int(iter), int(istep))
        return iKs.value, lG.value, eG.value

    def SolGMRe(self):
        iKs, lG = C.c_int(0), C.c_int(0)
        self.L.orc_solgmre(self.n, self.arr, _ptr(self.HBrg), _ptr(self.eBrg), _ptr(self.yBrg),
                           _ptr(self.Rcos), _ptr(self.Rsin), C.byref(iKs), C.byref(lG),
                           C.byref(self.ntotGM))
        return iKs.value, lG.value
