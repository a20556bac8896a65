"""Host-side mirror of the reference's solver interface for the hot path.

`PhastaGPU` plays the role of the Fortran caller (itrdrv.f:427-525): it owns
the mesh part, hands it once to `phb200_init`, and then calls `SolGMRe` /
`ElmGMRe` / `Au1GMR` / `i3LU` / `commu` / `sumgat` with the reference's
argument meaning (solgmr.f:1-8, elmgmr.f:1-6, au1gmr.f:1, i3lu.f:1,
commu.f:1, mpitools.f:107).  All numerics run in libphb200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib
from .lib import PhbCommon, PhbStep, PhbIncomp, MAXTOP, MAXQPT
from .mesh import MeshPart
from .params import SolverParams

KCLASS = {"assembly": 0, "asiq": 1, "ap": 2, "i3pre": 3, "blas1": 4, "node": 5, "halo": 6}


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class PhastaError(RuntimeError):
    pass


def _chk(rc, what):
    if rc != 0:
        raise PhastaError("phb200 %s failed (rc=%d); see stderr" % (what, rc))


class PhastaGPU:
    def __init__(self, part: MeshPart, params: SolverParams, tables: dict, device: int = 0):
        self.L = _lib.load()
        self.part, self.params, self.tables = part, params, tables
        nshape = max(int(b.shape[1]) for b in part.mien)
        self.nedof = 5 * nshape
        c = PhbCommon()
        c.nshg, c.numnp, c.numel = part.nshg, part.numnp, part.numel
        c.numelb = int(sum(b.shape[0] for b in part.mienb)) if part.nelblb else 0
        c.nflow, c.ndof, c.ndofBC, c.nshape, c.nedof = 5, 5, 6, nshape, self.nedof
        c.nelblk, c.nelblb, c.nlwork = part.nelblk, part.nelblb, part.nlwork
        c.numpe, c.myrank = part.numpe, part.rank
        for nm in ("ipord", "idiff", "itau", "iremoveStabTimeTerm", "EntropyPressure", "iDC", "Navier",
                   "Kspace", "nGMRES", "minIters", "matflg2", "matflg3"):
            setattr(c, nm, int(getattr(params, nm)))
        for nm in ("Rgas", "gamma", "gamma1", "pr", "datmat121", "datmat221", "datmat321", "datmat131",
                   "epsM", "dtsfct", "taucfct", "temper"):
            setattr(c, nm, float(getattr(params, nm)))
        for i in range(MAXTOP):
            c.nint[i] = int(tables["nint"][i])
            c.nintb[i] = int(tables["nintb"][i])
        q = np.asfortranarray(tables["Qwt"]).ravel(order="F")
        qb = np.asfortranarray(tables["Qwtb"]).ravel(order="F")
        C.memmove(c.Qwt, q.ctypes.data, q.nbytes)
        C.memmove(c.Qwtb, qb.ctypes.data, qb.nbytes)
        self.common = c
        k = self._keep = {}
        k["lcblk"] = np.asfortranarray(part.lcblk, dtype=np.int32)
        k["mien"] = [np.asfortranarray(b, dtype=np.int32) for b in part.mien]
        mien_ptrs = (C.POINTER(C.c_int) * len(k["mien"]))(*[_p(b, C.c_int) for b in k["mien"]])
        k["x"] = np.asfortranarray(part.x, dtype=np.float64)
        k["iBC"] = np.ascontiguousarray(part.iBC, dtype=np.int32)
        k["BC"] = np.asfortranarray(part.BC, dtype=np.float64)
        k["iper"] = np.ascontiguousarray(part.iper, dtype=np.int32)
        k["ilwork"] = np.ascontiguousarray(part.ilwork, dtype=np.int32)
        for nm in ("shp", "shgl", "shpb", "shglb"):
            k[nm] = np.asfortranarray(tables[nm], dtype=np.float64)
        lcblkb = ienb_ptrs = ibcb_ptrs = bcb_ptrs = None
        if part.nelblb:
            k["lcblkb"] = np.asfortranarray(part.lcblkb, dtype=np.int32)
            k["mienb"] = [np.asfortranarray(b, dtype=np.int32) for b in part.mienb]
            k["miBCB"] = [np.asfortranarray(b, dtype=np.int32) for b in part.miBCB]
            k["mBCB"] = [np.asfortranarray(b, dtype=np.float64) for b in part.mBCB]
            nb = len(k["mienb"])
            lcblkb = _p(k["lcblkb"], C.c_int)
            ienb_ptrs = (C.POINTER(C.c_int) * nb)(*[_p(b, C.c_int) for b in k["mienb"]])
            ibcb_ptrs = (C.POINTER(C.c_int) * nb)(*[_p(b, C.c_int) for b in k["miBCB"]])
            bcb_ptrs = (C.POINTER(C.c_double) * nb)(*[_p(b) for b in k["mBCB"]])
        self.ctx = C.c_void_p()
        _chk(self.L.phb200_init(C.byref(self.ctx), C.byref(c), _p(k["lcblk"], C.c_int), mien_ptrs,
                                lcblkb, ienb_ptrs, ibcb_ptrs, bcb_ptrs, _p(k["x"]), _p(k["iBC"], C.c_int), _p(k["BC"]),
                                _p(k["iper"], C.c_int), _p(k["ilwork"], C.c_int), _p(k["shp"]), _p(k["shgl"]),
                                _p(k["shpb"]), _p(k["shglb"]), int(device)), "init")
        K = params.Kspace
        self.HBrg = np.zeros((K + 1, K), order="F")
        self.eBrg = np.zeros(K + 1)
        self.yBrg = np.zeros(K + 1)
        self.Rcos = np.zeros(K + 1)
        self.Rsin = np.zeros(K + 1)
        self.ntotGM = 0
        self.iKs = 0
        self.lGMRES = 0

    # ------------------------------------------------------------------ util
    def close(self):
        if self.ctx:
            self.L.phb200_finalize(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, **over) -> PhbStep:
        P = self.params
        s = PhbStep()
        s.lhs, s.iprec, s.iter, s.nitr, s.lstep = P.lhs, P.iprec, 1, 1, 0
        s.Dtgl, s.almi, s.alfi, s.gami, s.etol = P.Dtgl, P.almi, P.alfi, P.gami, P.etol
        for k, v in over.items():
            setattr(s, k, v)
        return s

    def _vec(self, n=5):
        return np.zeros((self.part.nshg, n), order="F")

    # ------------------------------------------------- reference-named calls
    def SolGMRe(self, y, ac, yold=None, acold=None, *, step=None, want_bdiag=False):
        """solgmr.f:1-362.  Returns (res, Dy); res is the preconditioned
        residual, self.rmes the saved right-hand side (solgmr.f:83)."""
        st = step or self.step()
        y = np.asfortranarray(y, dtype=np.float64)
        ac = np.asfortranarray(ac, dtype=np.float64)
        res, rmes, Dy = self._vec(), self._vec(), self._vec()
        BD = np.zeros((self.part.nshg, 5, 5), order="F") if want_bdiag else None
        iKs, lG, ntot = C.c_int(0), C.c_int(0), C.c_int(self.ntotGM)
        _chk(self.L.phb200_solgmre(self.ctx, _p(y), _p(ac), C.byref(st), _p(res), _p(rmes), _p(BD), _p(Dy),
                                   _p(self.HBrg), _p(self.eBrg), _p(self.yBrg), _p(self.Rcos), _p(self.Rsin),
                                   C.byref(iKs), C.byref(lG), C.byref(ntot)), "solgmre")
        self.iKs, self.lGMRES, self.ntotGM = iKs.value, lG.value, ntot.value
        self.rmes, self.BDiag = rmes, BD
        return res, Dy

    def ElmGMRe(self, y, ac, *, step=None, want_egmass=False, want_qres=False):
        """elmgmr.f:1-274.  Returns dict(res, BDiag, EGmass?, qres?)."""
        st = step or self.step()
        y = np.asfortranarray(y, dtype=np.float64)
        ac = np.asfortranarray(ac, dtype=np.float64)
        res = self._vec()
        BD = np.zeros((self.part.nshg, 5, 5), order="F") if st.iprec else None
        EG = (np.zeros((self.part.numel, self.nedof, self.nedof), order="F")
              if (want_egmass and st.lhs == 1) else None)
        qres = self._vec(12) if want_qres else None
        _chk(self.L.phb200_elmgmre(self.ctx, _p(y), _p(ac), C.byref(st), _p(res), _p(BD), _p(EG), _p(qres)),
             "elmgmre")
        return dict(res=res, BDiag=BD, EGmass=EG, qres=qres)

    # ------------------------------------------------------ block-CSR flavour
    def genadj(self, nnz=35):
        """common/genadj.f: returns (colm, rowp, nnz_tot) with the reference's
        1-based conventions; the device keeps the structure it has just built."""
        nshg = self.part.nshg
        colm = np.zeros(nshg + 1, dtype=np.int32)
        rowp = np.zeros(nnz * nshg, dtype=np.int32)
        ntot = C.c_int(0)
        _chk(self.L.phb200_genadj(self.ctx, int(nnz), _p(colm, C.c_int), _p(rowp, C.c_int), C.byref(ntot)),
             "genadj")
        # (phb200_genadj builds the structure on the device and installs it: no phb200_set_sparse round trip)
        self.colm, self.rowp, self.nnz_tot = colm, rowp, ntot.value
        return colm, rowp[:ntot.value], ntot.value

    def set_sparse(self, colm, rowp, nnz_tot):
        colm = np.ascontiguousarray(colm, dtype=np.int32)
        rowp = np.ascontiguousarray(rowp, dtype=np.int32)
        self.nnz_tot = int(nnz_tot)
        _chk(self.L.phb200_set_sparse(self.ctx, _p(colm, C.c_int), _p(rowp, C.c_int), int(nnz_tot)), "set_sparse")

    def ElmGMRs(self, y, ac, *, step=None, want_lhsk=False):
        """elmgmr.f:280-612 (+ fillsparseC).  Returns dict(res, BDiag, lhsK?)."""
        st = step or self.step()
        y = np.asfortranarray(y, dtype=np.float64)
        ac = np.asfortranarray(ac, dtype=np.float64)
        res = self._vec()
        BD = np.zeros((self.part.nshg, 5, 5), order="F") if st.iprec else None
        K = np.zeros((25, self.nnz_tot), order="F") if (want_lhsk and st.lhs == 1) else None
        _chk(self.L.phb200_elmgmrs(self.ctx, _p(y), _p(ac), C.byref(st), _p(res), _p(BD), _p(K)), "elmgmrs")
        return dict(res=res, BDiag=BD, lhsK=K)

    def Spsi3pre(self, want_lhsk=False):
        K = np.zeros((25, self.nnz_tot), order="F") if want_lhsk else None
        _chk(self.L.phb200_spsi3pre(self.ctx, _p(K)), "spsi3pre")
        return K

    def SparseAp(self, p):
        """sparseap.f:1-138, in place on p(nshg,5)."""
        assert p.flags.f_contiguous and p.shape == (self.part.nshg, 5)
        _chk(self.L.phb200_sparseap(self.ctx, _p(p)), "sparseap")
        return p

    def SolGMRs(self, y, ac, yold=None, acold=None, *, step=None):
        """solgmr.f:368-744 with the CSR structure given to set_sparse/genadj."""
        st = step or self.step()
        y = np.asfortranarray(y, dtype=np.float64)
        ac = np.asfortranarray(ac, dtype=np.float64)
        res, rmes, Dy = self._vec(), self._vec(), self._vec()
        iKs, lG, ntot = C.c_int(0), C.c_int(0), C.c_int(self.ntotGM)
        _chk(self.L.phb200_solgmrs(self.ctx, _p(y), _p(ac), C.byref(st), _p(res), _p(rmes), None, _p(Dy),
                                   _p(self.HBrg), _p(self.eBrg), _p(self.yBrg), _p(self.Rcos), _p(self.Rsin),
                                   C.byref(iKs), C.byref(lG), C.byref(ntot)), "solgmrs")
        self.iKs, self.lGMRES, self.ntotGM = iKs.value, lG.value, ntot.value
        self.rmes = rmes
        return res, Dy

    # ------------------------------------------------------ matrix-free flavour
    def ElmMFG(self, y, ac, *, step=None):
        """elmmfg.f:1-256.  Returns dict(res, rmes, BDiag): residual, modified
        residual and the e3bdg block diagonal (lhs=0, iprec=1; itrdrv.f:496-498)."""
        st = step or self.step(lhs=0, iprec=1)
        y = np.asfortranarray(y, dtype=np.float64)
        ac = np.asfortranarray(ac, dtype=np.float64)
        res, rmes = self._vec(), self._vec()
        BD = np.zeros((self.part.nshg, 5, 5), order="F") if st.iprec else None
        _chk(self.L.phb200_elmmfg(self.ctx, _p(y), _p(ac), C.byref(st), _p(res), _p(rmes), _p(BD)), "elmmfg")
        return dict(res=res, rmes=rmes, BDiag=BD)

    def ItrRes(self, yp, iabres=0):
        """itrres.f:1-171: modified residual of yp(nshg,5) {u,v,w,p,T}."""
        yp = np.asfortranarray(yp, dtype=np.float64)
        out = self._vec()
        _chk(self.L.phb200_itrres(self.ctx, _p(yp), _p(out), int(iabres)), "itrres")
        return out

    def Au1MFG(self, u, eGMRES, setup=True):
        """au1mfg.f:1-98 on a copy of u; setup performs solmfg.f:97-135 first."""
        u = np.asfortranarray(u, dtype=np.float64).copy(order="F")
        _chk(self.L.phb200_au1mfg(self.ctx, _p(u), C.c_double(eGMRES), int(bool(setup))), "au1mfg")
        return u

    def SolMFG(self, y, ac, *, step=None, eGMRES=None):
        """solmfg.f:1-381.  Returns (res, Dy); self.eGMRES carries COMMON /itrpar/'s
        interval between calls."""
        st = step or self.step(lhs=0, iprec=1)
        y = np.asfortranarray(y, dtype=np.float64)
        ac = np.asfortranarray(ac, dtype=np.float64)
        res, Dy = self._vec(), self._vec()
        BD = np.zeros((self.part.nshg, 5, 5), order="F")
        iKs, lG, ntot = C.c_int(0), C.c_int(0), C.c_int(self.ntotGM)
        eG = C.c_double(self.eGMRES if eGMRES is None else eGMRES)
        _chk(self.L.phb200_solmfg(self.ctx, _p(y), _p(ac), C.byref(st), _p(res), _p(BD), _p(Dy), _p(self.HBrg),
                                  C.byref(iKs), C.byref(lG), C.byref(ntot), C.byref(eG)), "solmfg")
        self.iKs, self.lGMRES, self.ntotGM, self.BDiag = iKs.value, lG.value, ntot.value, BD
        return res, Dy

    def dev_elmmfg(self, step=None):
        st = step or self.step(lhs=0, iprec=1)
        _chk(self.L.phb200_dev_elmmfg(self.ctx, C.byref(st)), "dev_elmmfg")

    def dev_solve_mfg(self, step=None):
        st = step or self.step(lhs=0, iprec=1)
        iKs, lG, ntot = C.c_int(0), C.c_int(0), C.c_int(self.ntotGM)
        _chk(self.L.phb200_dev_solve_mfg(self.ctx, C.byref(st), C.byref(iKs), C.byref(lG), C.byref(ntot)),
             "dev_solve_mfg")
        self.iKs, self.lGMRES, self.ntotGM = iKs.value, lG.value, ntot.value

    @property
    def eGMRES(self):
        e = C.c_double(0)
        _chk(self.L.phb200_egmres(self.ctx, C.byref(e), 0), "egmres")
        return e.value

    @eGMRES.setter
    def eGMRES(self, v):
        e = C.c_double(v)
        _chk(self.L.phb200_egmres(self.ctx, C.byref(e), 1), "egmres")

    def dev_au1mfg(self, slot=0):
        _chk(self.L.phb200_dev_au1mfg(self.ctx, int(slot)), "dev_au1mfg")

    def dev_elmgmrs(self, step=None):
        st = step or self.step()
        _chk(self.L.phb200_dev_elmgmrs(self.ctx, C.byref(st)), "dev_elmgmrs")

    def dev_solve_sparse(self, step=None):
        st = step or self.step()
        iKs, lG, ntot = C.c_int(0), C.c_int(0), C.c_int(self.ntotGM)
        _chk(self.L.phb200_dev_solve_sparse(self.ctx, C.byref(st), C.byref(iKs), C.byref(lG), C.byref(ntot)),
             "dev_solve_sparse")
        self.iKs, self.lGMRES, self.ntotGM = iKs.value, lG.value, ntot.value
        return self.iKs

    def dev_sparseap(self, slot=0):
        _chk(self.L.phb200_dev_sparseap(self.ctx, int(slot)), "dev_sparseap")

    def i3LU(self, Diag, r, code):
        """i3lu.f:1-181; code 'LU_Fact'|'forward'|'backward'|'product'."""
        ic = {"LU_Fact": 0, "forward": 1, "backward": 2, "product": 3}[code.strip()]
        _chk(self.L.phb200_i3lu(self.ctx, _p(Diag), _p(r), ic), "i3lu")

    def i3pre(self, want_egmass=False):
        EG = np.zeros((self.part.numel, self.nedof, self.nedof), order="F") if want_egmass else None
        _chk(self.L.phb200_i3pre(self.ctx, _p(EG)), "i3pre")
        return EG

    def Au1GMR(self, uBrg):
        """au1gmr.f:1-106, in place on uBrg(nshg,5)."""
        assert uBrg.flags.f_contiguous and uBrg.shape == (self.part.nshg, 5)
        _chk(self.L.phb200_au1gmr(self.ctx, _p(uBrg)), "au1gmr")
        return uBrg

    def bc3per(self, r):
        _chk(self.L.phb200_bc3per(self.ctx, _p(r)), "bc3per")
        return r

    def commu(self, global_, n, code):
        """commu.f:1-297; code 'in ' | 'out'."""
        ic = {"in": 0, "out": 1}[code.strip()]
        _chk(self.L.phb200_commu(self.ctx, _p(global_), int(n), ic), "commu")
        return global_

    def sumgat(self, u, n):
        out = C.c_double(0.0)
        _chk(self.L.phb200_sumgat(self.ctx, _p(np.asfortranarray(u)), int(n), C.byref(out)), "sumgat")
        return out.value

    def aerfrc(self, zero=False):
        """COMMON /aerfrc/: (Force(3), HFlux, flxID(10,0:MAXSURF))."""
        F = np.zeros(3)
        H = C.c_double(0)
        fl = np.zeros((10, 1001), order="F")
        _chk(self.L.phb200_get_aerfrc(self.ctx, _p(F), C.byref(H), _p(fl), int(zero)), "get_aerfrc")
        return F, H.value, fl

    # --------------------------------------------------- HBM-resident path
    def set_state(self, y, ac):
        self._y = np.asfortranarray(y, dtype=np.float64)
        self._ac = np.asfortranarray(ac, dtype=np.float64)
        _chk(self.L.phb200_set_state(self.ctx, _p(self._y), _p(self._ac)), "set_state")

    # ---- incompressible flavour (incompressible/elmgmr.f ElmGMR, lesSparse.f) ----
    def IncElmGMR(self, y, ac, ip, *, want_lhs=True, **over):
        """ElmGMR(u, y, ac, x, shp, shgl, iBC, BC, shpb, shglb, res, iper, ilwork, rowp, colm, lhsK, lhsP, ...)
        of the incompressible code (elmgmr.f:1-6): returns dict(res (nshg,4)[, lhsK (9,nnz_tot), lhsP (4,nnz_tot)]).
        genadj / set_sparse must have been called (itrdrv does it once)."""
        s = PhbIncomp.from_params(ip, **over)
        y = np.asfortranarray(y, dtype=np.float64)
        ac = np.asfortranarray(ac, dtype=np.float64)
        out = {"res": np.zeros((self.part.nshg, 4), order="F")}
        lhs = bool(s.lhs) and want_lhs
        if lhs:
            out["lhsK"] = np.zeros((9, self.nnz_tot), order="F")
            out["lhsP"] = np.zeros((4, self.nnz_tot), order="F")
        _chk(self.L.phb200_inc_elmgmr(self.ctx, _p(y), _p(ac), C.byref(s), _p(out["res"]),
                                      _p(out.get("lhsK")), _p(out.get("lhsP"))), "inc_elmgmr")
        return out

    def dev_inc_elmgmr(self, ip, **over):
        s = PhbIncomp.from_params(ip, **over)
        _chk(self.L.phb200_inc_dev_elmgmr(self.ctx, C.byref(s)), "inc_dev_elmgmr")

    def LesAp(self, kind, p):
        """fLesSparseAp{G,KG,NGt,NGtC,Full}(col, row, kLhs, pLhs, p, q, nNodes, nnz_tot) (lesSparse.f:204-492)
        on the device-resident lhsK/lhsP of the last IncElmGMR."""
        k = {"G": 0, "KG": 1, "NGt": 2, "NGtC": 3, "Full": 4}[kind]
        n = self.part.nshg
        p = np.asfortranarray(p, dtype=np.float64)
        shape = {0: (n, 3), 1: (n, 3), 2: (n,), 3: (n,), 4: (n, 4)}[k]
        q = np.zeros(shape, order="F")
        _chk(self.L.phb200_les_ap(self.ctx, k, _p(p), _p(q)), "les_ap")
        return q

    def dev_inc_apfull(self):
        _chk(self.L.phb200_inc_dev_apfull(self.ctx), "inc_dev_apfull")

    # ------------------------------------- Newton / time-step shell (timestep.cu)
    def set_old_state(self, yold, acold):
        yo = np.asfortranarray(yold, dtype=np.float64)
        ao = np.asfortranarray(acold, dtype=np.float64)
        _chk(self.L.phb200_set_old_state(self.ctx, _p(yo), _p(ao)), "set_old_state")

    def get_state(self, old=False):
        """(y, ac) or, with old=True, (y, ac, yold, acold) copied back from the device."""
        arrs = [self._vec() for _ in range(4 if old else 2)]
        ptrs = [_p(a) for a in arrs] + [None] * (4 - len(arrs))
        _chk(self.L.phb200_get_state(self.ctx, *ptrs), "get_state")
        return tuple(arrs)

    def itrPredict(self, ipred=1, step=None):
        """itrPC.f:54-119 on the resident state."""
        _chk(self.L.phb200_itrpredict(self.ctx, C.byref(step or self.step()), int(ipred)), "itrpredict")

    def itrBC(self, ires=1):
        """itrbc.f:1-199 on the resident y/ac (incl. commu 'out')."""
        _chk(self.L.phb200_itrbc(self.ctx, int(ires)), "itrbc")

    def itrCorrect(self, step=None):
        """itrPC.f:127-150 with the Dy of the last solve."""
        _chk(self.L.phb200_itrcorrect(self.ctx, C.byref(step or self.step())), "itrcorrect")

    def itrUpdate(self, step=None):
        """itrPC.f:205-210."""
        _chk(self.L.phb200_itrupdate(self.ctx, C.byref(step or self.step())), "itrupdate")

    def rstat(self, nshgt=None):
        """rstat.f:94-112: totres(1:2) of the last solve."""
        out = np.zeros(2)
        _chk(self.L.phb200_rstat(self.ctx, C.c_longlong(int(nshgt or self.part.nshg)), _p(out)), "rstat")
        return out

    def TimeStep(self, nitr=2, ipred=1, sparse=False, LHSupd=1, nshgt=None, step=None):
        """One step of itrdrv.f's flow sequence on the resident state (set_state +
        set_old_state first).  Returns stats (nitr,6): totres(1), totres(2), iKs, lGMRES, lhs, 0."""
        st = step or self.step()
        stats = np.zeros((nitr, 6))
        ntot = C.c_int(self.ntotGM)
        _chk(self.L.phb200_timestep(self.ctx, C.byref(st), int(ipred), int(nitr), int(bool(sparse)), int(LHSupd),
                                    C.c_longlong(int(nshgt or self.part.nshg)), C.byref(ntot), _p(stats)),
             "timestep")
        self.ntotGM = ntot.value
        return stats

    def dev_elmgmre(self, step=None):
        st = step or self.step()
        _chk(self.L.phb200_dev_elmgmre(self.ctx, C.byref(st)), "dev_elmgmre")

    def dev_solve(self, step=None):
        st = step or self.step()
        iKs, lG, ntot = C.c_int(0), C.c_int(0), C.c_int(self.ntotGM)
        _chk(self.L.phb200_dev_solve(self.ctx, C.byref(st), C.byref(iKs), C.byref(lG), C.byref(ntot)), "dev_solve")
        self.iKs, self.lGMRES, self.ntotGM = iKs.value, lG.value, ntot.value
        return self.iKs

    def dev_ap(self, slot=0):
        _chk(self.L.phb200_dev_ap(self.ctx, int(slot)), "dev_ap")

    def get(self, what):
        nshg = self.part.nshg
        if what == "res":
            a = self._vec()
            _chk(self.L.phb200_get_res(self.ctx, _p(a)), "get_res")
        elif what == "Dy":
            a = self._vec()
            _chk(self.L.phb200_get_dy(self.ctx, _p(a)), "get_dy")
        elif what == "BDiag":
            a = np.zeros((nshg, 5, 5), order="F")
            _chk(self.L.phb200_get_bdiag(self.ctx, _p(a)), "get_bdiag")
        elif what == "EGmass":
            a = np.zeros((self.part.numel, self.nedof, self.nedof), order="F")
            _chk(self.L.phb200_get_egmass(self.ctx, _p(a)), "get_egmass")
        else:
            raise KeyError(what)
        return a

    def get_egmass_range(self, e0, n):
        """EGmass(e0:e0+n, :, :) (0-based element range of the reference's order) without moving the rest"""
        a = np.zeros((int(n), self.nedof, self.nedof), order="F")
        _chk(self.L.phb200_get_egmass_range(self.ctx, C.c_longlong(int(e0)), int(n), _p(a)), "get_egmass_range")
        return a

    def get_lhsk_range(self, k0, n):
        """lhsK(:, k0:k0+n) (0-based CSR block range)"""
        a = np.zeros((25, int(n)), order="F")
        _chk(self.L.phb200_get_lhsk_range(self.ctx, C.c_longlong(int(k0)), C.c_longlong(int(n)), _p(a)),
             "get_lhsk_range")
        return a

    # ------------------------------------------------------ instrumentation
    def sync(self):
        _chk(self.L.phb200_sync(self.ctx), "sync")

    def event(self, slot):
        _chk(self.L.phb200_event_record(self.ctx, slot), "event_record")

    def elapsed_ms(self, a, b):
        ms = C.c_float(0)
        _chk(self.L.phb200_event_elapsed_ms(self.ctx, a, b, C.byref(ms)), "event_elapsed")
        return ms.value

    def launches(self):
        return int(self.L.phb200_launch_count(self.ctx))

    def profile(self, on):
        _chk(self.L.phb200_profile(self.ctx, int(on)), "profile")

    def profile_get(self):
        out = {}
        for name, k in KCLASS.items():
            ms, n = C.c_float(0), C.c_longlong(0)
            self.L.phb200_profile_get(self.ctx, k, C.byref(ms), C.byref(n))
            out[name] = (ms.value, n.value)
        return out

    def profile_reset(self):
        self.L.phb200_profile_reset(self.ctx)

    def fp64_peak(self):
        t = C.c_double(0)
        _chk(self.L.phb200_fp64_peak(self.ctx, C.byref(t)), "fp64_peak")
        return t.value

    def set_deterministic(self, on=True):
        """ordered gather instead of FP64 atomics for qres / res / BDiag (lhs=1 assemblies of linear tets)"""
        _chk(self.L.phb200_set_deterministic(self.ctx, int(bool(on))), "set_deterministic")

    def dmma_peak(self):
        t = C.c_double(0)
        _chk(self.L.phb200_dmma_peak(self.ctx, C.byref(t)), "dmma_peak")
        return t.value

    def red_peak(self, nblk):
        """G FP64 scatter-adds per second into nblk random 25-double blocks (microbenchmark)"""
        t = C.c_double(0)
        _chk(self.L.phb200_red_peak(self.ctx, C.c_longlong(int(nblk)), C.byref(t)), "red_peak")
        return t.value

    def flush_l2(self):
        _chk(self.L.phb200_flush_l2(self.ctx), "flush_l2")

    # multi-GPU
    def comm_init(self, id128: bytes):
        buf = C.create_string_buffer(id128, 128)
        _chk(self.L.phb200_comm_init(self.ctx, buf), "comm_init")

    def local_group_join(self, nranks):
        _chk(self.L.phb200_local_group_join(self.ctx, int(nranks)), "local_group_join")


def nccl_unique_id() -> bytes:
    L = _lib.load()
    buf = C.create_string_buffer(128)
    _chk(L.phb200_nccl_unique_id(buf), "nccl_unique_id")
    return buf.raw
