"""ctypes binding of libphb200.so (include/phb200.h).

This is the same binding surface the Fortran ISO_C_BINDING shim uses
(INTEGRATION.md).  The library is CUDA-only: if it is missing or no GPU is
visible the product path raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

MAXTOP, MAXSH, MAXQPT = 6, 32, 125
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libphb200.so")


class PhbCommon(C.Structure):
    """struct phb200_common (include/phb200.h)."""
    _fields_ = [
        *[(n, C.c_int) for n in (
            "nshg", "numnp", "numel", "numelb", "nflow", "ndof", "ndofBC", "nshape", "nedof",
            "nelblk", "nelblb", "nlwork", "numpe", "myrank",
            "ipord", "idiff", "itau", "iremoveStabTimeTerm", "EntropyPressure",
            "iDC", "Navier", "Kspace", "nGMRES", "minIters",
            "matflg2", "matflg3")],
        *[(n, C.c_double) for n in (
            "Rgas", "gamma", "gamma1", "pr", "datmat121", "datmat221", "datmat321", "datmat131",
            "epsM", "dtsfct", "taucfct", "temper")],
        ("nint", C.c_int * MAXTOP), ("nintb", C.c_int * MAXTOP),
        ("Qwt", C.c_double * (MAXTOP * MAXQPT)), ("Qwtb", C.c_double * (MAXTOP * MAXQPT)),
    ]


class PhbStep(C.Structure):
    """struct phb200_step."""
    _fields_ = [*[(n, C.c_int) for n in ("lhs", "iprec", "iter", "nitr", "lstep", "istep")],
                *[(n, C.c_double) for n in ("Dtgl", "almi", "alfi", "gami", "etol")]]


class PhbIncomp(C.Structure):
    """struct phb200_incomp."""
    _fields_ = [*[(n, C.c_int) for n in ("iconvflow", "itau", "idiff", "ipord", "lhs", "matflg5")],
                ("rho", C.c_double), ("rmu", C.c_double), ("bf", C.c_double * 3),
                *[(n, C.c_double) for n in ("flmpl", "flmpr", "Delt", "Dtgl", "almi", "alfi", "gami",
                                            "dtsfct", "taucfct")],
                ("iviscflux", C.c_int), ("itwmod", C.c_int), ("nsrflist", C.POINTER(C.c_int))]

    @classmethod
    def from_params(cls, ip, **over):
        s = cls()
        for n in ("iconvflow", "itau", "idiff", "ipord", "lhs", "matflg5"):
            setattr(s, n, int(over.get(n, getattr(ip, n))))
        for n in ("rho", "rmu", "flmpl", "flmpr", "Delt", "Dtgl", "almi", "alfi", "gami", "dtsfct", "taucfct"):
            setattr(s, n, float(over.get(n, getattr(ip, n))))
        for i in range(3):
            s.bf[i] = float(ip.bf[i])
        s.iviscflux, s.itwmod = int(ip.iviscflux), int(ip.itwmod)
        s._nsrf = (C.c_int * 1001)()            # nsrflist(0:MAXSURF); lives as long as the struct
        for k in ip.surfaces:
            s._nsrf[int(k)] = 1
        s.nsrflist = C.cast(s._nsrf, C.POINTER(C.c_int))
        return s


# every symbol include/phb200.h declares (tests/test_abi.py checks the .so exports them all)
SYMBOLS = [
    "phb200_init", "phb200_finalize", "phb200_nccl_unique_id", "phb200_comm_init",
    "phb200_solgmre", "phb200_elmgmre", "phb200_i3lu", "phb200_i3pre", "phb200_au1gmr",
    "phb200_bc3per", "phb200_commu", "phb200_sumgat", "phb200_set_state", "phb200_dev_elmgmre",
    "phb200_dev_solve", "phb200_dev_ap", "phb200_get_res", "phb200_get_dy", "phb200_get_bdiag",
    "phb200_get_egmass", "phb200_get_egmass_range", "phb200_get_lhsk_range", "phb200_get_aerfrc", "phb200_event_record", "phb200_event_elapsed_ms", "phb200_sync",
    "phb200_launch_count", "phb200_profile", "phb200_profile_get", "phb200_profile_reset",
    "phb200_set_deterministic", "phb200_fp64_peak", "phb200_dmma_peak", "phb200_red_peak", "phb200_flush_l2", "phb200_version", "phb200_sizeof_common",
    "phb200_sizeof_step", "phb200_local_group_join", "phb200_genadj", "phb200_set_sparse",
    "phb200_elmgmrs", "phb200_spsi3pre", "phb200_sparseap", "phb200_solgmrs", "phb200_dev_elmgmrs",
    "phb200_dev_solve_sparse", "phb200_dev_sparseap",
    "phb200_set_old_state", "phb200_get_state", "phb200_itrpredict", "phb200_itrbc", "phb200_itrcorrect",
    "phb200_itrupdate", "phb200_rstat", "phb200_timestep",
    "phb200_solmfg", "phb200_elmmfg", "phb200_itrres", "phb200_au1mfg", "phb200_dev_elmmfg",
    "phb200_dev_solve_mfg", "phb200_dev_au1mfg", "phb200_egmres",
    "phb200_inc_elmgmr", "phb200_inc_dev_elmgmr", "phb200_les_ap", "phb200_inc_dev_apfull", "phb200_sizeof_incomp",
]

_LIB = None


def build(verbose=False):
    """Compile libphb200.so in-tree for sm_100a (nvcc cross-compiles on CPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")], stdout=out)
    return LIB_PATH


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libphb200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                "phasta_b200 has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.phb200_version.restype = C.c_char_p
        lib.phb200_launch_count.restype = C.c_longlong
        lib.phb200_finalize.restype = None
        if lib.phb200_sizeof_common() != C.sizeof(PhbCommon) or lib.phb200_sizeof_step() != C.sizeof(PhbStep) \
                or lib.phb200_sizeof_incomp() != C.sizeof(PhbIncomp):
            raise RuntimeError('phb200 struct layout mismatch between include/phb200.h and lib.py')
        _LIB = lib
    return _LIB
