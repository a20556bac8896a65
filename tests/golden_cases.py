"""The cases behind tests/golden/f77_*.npz: seeded make_case() arguments and the
reference routines run on each (tests/golden/make_golden_f77.py executes the
reference's own Fortran on them; tests/test_golden_f77.py pins the oracle and
the CUDA path against the stored outputs)."""
import numpy as np

# name -> (make_case args, kwargs, runs)
CASES = {
    "tet_channel_bnd": ((3, 2, 2), dict(bc="channel", ibksiz=16, boundary=True, natural="mixed", etol=1e-6),
                        ("elmgmre", "solgmre", "solgmrs")),
    "tet_allbc": ((3, 3, 2), dict(bc="mixed", ibksiz=16, etol=1e-6), ("elmgmre", "solgmre")),
    # every velocity code 1..7, density and pressure+temperature codes on interior nodes (bc3LHS/bc3Res/bc3BDg branches)
    "tet_allcodes": ((4, 4, 3), dict(bc="allcodes", ibksiz=50, etol=1e-6), ("elmgmre", "solgmre", "solgmrs")),
    # discontinuity capturing (e3dc.f): iDC = 1 (the DC viscosity of e3dc.f:208-222), 2 and 3
    "tet_dc1": ((3, 3, 2), dict(bc="channel", ibksiz=16, iDC=1, etol=1e-6), ("elmgmre", "solgmre")),
    "tet_dc2": ((3, 2, 2), dict(bc="channel", ibksiz=64, iDC=2), ("elmgmre",)),
    "tet_dc3": ((3, 2, 2), dict(bc="channel", ibksiz=64, iDC=3), ("elmgmre", "elmgmre0")),
    "hex_dc1": ((2, 2, 2), dict(bc="channel", topo="hex", ibksiz=8, iDC=1), ("elmgmre",)),
    "wedge_dc3": ((2, 3, 2), dict(bc="channel", topo="wedge", ibksiz=16, iDC=3), ("elmgmre", "elmgmre0")),
    "mixed_dc1": ((2, 4, 2), dict(bc="channel", topo="mixed", ibksiz=16, iDC=1, etol=1e-6), ("elmgmre", "solgmrs")),
    "tet_1pt_nodiff": ((3, 2, 2), dict(bc="channel", ibksiz=64, rule=1, idiff=0, etol=1e-6), ("elmgmre", "solgmre")),
    "tet_sutherland": ((2, 2, 2), dict(bc="channel", ibksiz=64, matflg2=1, etol=1e-6), ("elmgmre",)),
    "tet_resonly": ((3, 2, 2), dict(bc="channel", ibksiz=16, boundary=True, natural="mixed"), ("elmgmre0",)),
    "hex_channel": ((3, 2, 2), dict(bc="channel", topo="hex", ibksiz=8, etol=1e-6), ("elmgmre", "solgmre", "solgmrs")),
    "wedge_allbc": ((2, 3, 2), dict(bc="mixed", topo="wedge", ibksiz=16, etol=1e-6), ("elmgmre", "solgmre")),
    # SolGMRe is not run on the mixed mesh: i3pre.f:53 passes the strided section BDiagl(iel:inum,:,:,:) of an
    # (numel,nshape,5,5) array to local's (npro,nshl,25) dummy, which scrambles the tet blocks when nshl<nshape
    # (the reference's EBE solver is only well defined on single-topology meshes; its default SolGMRs is fine)
    "mixed_channel": ((2, 4, 2), dict(bc="channel", topo="mixed", ibksiz=16, etol=1e-6), ("elmgmre", "solgmrs")),
    # boundary elements on quadrilateral faces of hexes (lcsyst 2), triangular (3) and quadrilateral (4) faces of
    # wedges (e3bvar.f:139-176 per-topology normals / WdetJb, getbnodes lnode), every natural-BC code; z faces too
    "hex_bnd": ((3, 2, 2), dict(bc="channel", topo="hex", ibksiz=8, boundary=True, natural="mixed", periodic_z=False,
                                etol=1e-6), ("elmgmre", "solgmre", "solgmrs")),
    "wedge_bnd": ((2, 3, 2), dict(bc="channel", topo="wedge", ibksiz=16, boundary=True, natural="mixed",
                                  periodic_z=False, etol=1e-6), ("elmgmre", "elmgmre0", "solgmrs")),
    "mixed_bnd": ((2, 4, 2), dict(bc="channel", topo="mixed", ibksiz=16, boundary=True, natural="mixed", etol=1e-6),
                  ("elmgmre", "solgmrs")),
    # matrix-free flavour (SolMFG): acoustic units (see common.nondimensional) and a state that has been through
    # itrBC, as in itrdrv.f:394 -- Au1MFG applies itrBC to the perturbed state
    "tet_nd_mfg": ((3, 2, 2), dict(bc="channel", ibksiz=16, boundary=True, natural="mixed", etol=1e-4, nd=True),
                   ("solmfg",)),
    "hex_nd_mfg": ((2, 2, 2), dict(bc="channel", topo="hex", ibksiz=8, etol=1e-4, nd=True), ("solmfg",)),
    # the matrix-free flavour with discontinuity capturing: ElmMFG runs e3dc with ires=3 (incl. the rmi(:,11) statement
    # of e3dc.f:262), ItrRes / Au1MFG with ires=2
    "tet_nd_mfg_dc1": ((3, 2, 2), dict(bc="channel", ibksiz=16, boundary=True, natural="mixed", etol=1e-4, nd=True,
                                       iDC=1), ("solmfg",)),
    "hex_nd_mfg_dc3": ((2, 2, 2), dict(bc="channel", topo="hex", ibksiz=8, etol=1e-4, nd=True, iDC=3), ("solmfg",)),
    # the same without essential BCs, periodicity or boundary elements: residuals and block diagonal are then the raw
    # element sums, which tests/test_bnd_kernel_host.py compares with the kernels run on the host
    "tet_nd_mfg_dc1_raw": ((3, 2, 2), dict(bc="none", periodic_z=False, ibksiz=16, etol=1e-4, nd=True, iDC=1), ("solmfg",)),
    "hex_nd_mfg_dc3_raw": ((2, 2, 2), dict(bc="none", periodic_z=False, topo="hex", ibksiz=8, etol=1e-4, nd=True, iDC=3),
                           ("solmfg",)),
    # the matrix-free flavour through wedge boundary faces (ElmMFG -> AsBMFG on lcsyst 3 and 4)
    "wedge_bnd_nd_mfg": ((2, 3, 2), dict(bc="channel", topo="wedge", ibksiz=16, boundary=True, natural="mixed",
                                         periodic_z=False, etol=1e-4, nd=True), ("solmfg",)),
}


def build_case(name):
    from common import make_case, nondimensional
    a, kw, runs = CASES[name]
    kw = dict(kw)
    nd = kw.pop("nd", False)
    case = make_case(*a, **kw)
    if nd:
        case = nondimensional(case)
    return case, runs


def input_digest(case):
    """guards the fixtures against drift of the seeded generators"""
    params, tables, parts, states = case
    mp = parts[0]
    y, ac = states[0]
    return np.array([float(np.sum(mp.x)), float(np.sum(np.abs(y))), float(np.sum(np.abs(ac))),
                     float(sum(int(np.sum(b.astype(np.int64))) for b in mp.mien)), float(np.sum(mp.iBC)),
                     float(np.sum(mp.BC))])
