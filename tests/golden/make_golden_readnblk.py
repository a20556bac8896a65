"""Pins the phastaIO file format of this repo (phasta_b200/phio.py: geombc.dat.<n>, restart.<step>.<n>) to the
reference's own READER: common/readnblk.f -- with genblkPosix.f, gensav.f, genbkbPosix.f, gensvb.f underneath -- is
executed UNMODIFIED by f77np, its phio_* calls (the `use phio` interface of common/phio.f90: phio_openfile,
phio_readheader, phio_readdatablock, phio_closefile) bound to phasta_b200.phio.PhioFile on the files that
phio.write_geombc / write_restart produced.  What the Fortran ends up holding -- the /conpar/ scalars, x, nBC,
iBCtmp, BCinp, iper, the element and boundary blocks, qold, acold, lstep -- is stored in
tests/golden/f77_readnblk.npz and compared in tests/test_phio.py with what phio.read_geombc / read_restart return
from the same files.

    python tests/golden/make_golden_readnblk.py [--check]
"""
import argparse
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from f77np import Program, ScalarRef, scan_functions  # noqa: E402

REF = "/root/reference/phSolver/common"
# name -> (make_case arguments, IBKSZ, lstep, rank read)
CASES = {
    "tet_bnd": (dict(nx=5, ny=4, nz=3, topo="tet", boundary=True, natural="mixed", bc="channel"), 32, 120, 0),
    "mixed_allcodes": (dict(nx=6, ny=6, nz=4, topo="mixed", boundary=True, natural="mixed", bc="allcodes"), 20, 7, 0),
    # the middle part of three: the numpe > 1 branch ('size of ilwork array', 'ilwork', ctypes), split segments
    "tet_part2of3": (dict(nx=6, ny=3, nz=3, topo="tet", boundary=True, natural="mixed", bc="channel", nparts=3, max_seg=7),
                     32, 40, 1),
}


def build_case(name):
    from common import make_case
    kw, ibksz, lstep, rank = CASES[name]
    kw = dict(kw)
    nx, ny, nz = kw.pop("nx"), kw.pop("ny"), kw.pop("nz")
    return make_case(nx, ny, nz, ibksiz=ibksz, **kw), ibksz, lstep, rank


def run_readnblk(root, ibksz, lstep, rank=0, numpe=1):
    """readnblk.f as rank `rank` of `numpe` on <root>/<numpe>-procs_case/{geombc.dat.<rank+1>, restart.<lstep>.<rank+1>}:
    returns a dict of what it read"""
    from phasta_b200 import phio

    d = phio.case_dir(root, numpe)
    files = {"geombc": os.path.join(d, "geombc.dat.%d" % (rank + 1)),
             "restart": os.path.join(d, "restart.%d.%d" % (lstep, rank + 1))}
    state = {"next": None, "f": None, "opened": []}

    def construct(prog, fh, kind, fname):
        state["next"] = kind.rstrip("\0").strip()

    def openfile(prog, fname, fh):
        state["f"] = phio.PhioFile(files[state["next"]], "r")
        state["opened"].append(state["next"])

    def closefile(prog, fh):
        state["f"].close()
        state["f"] = None

    def readheader(prog, fh, phrase, target, n, dtype, iotype):
        h = state["f"].readheader(phrase.rstrip("\0").strip(), int(n), "integer")
        if isinstance(target, ScalarRef):
            target.set(int(h[0]) if len(h) else 0)
        else:
            target[:len(h)] = h

    def readdatablock(prog, fh, phrase, arr, n, dtype, iotype):
        if int(n) <= 0:
            return
        data = state["f"].readdatablock(phrase.rstrip("\0").strip(), int(n), dtype.rstrip("\0").strip())
        arr.reshape(-1, order="F")[:int(n)] = data

    def error(prog, *a):
        raise RuntimeError("reference called error(%r)" % (a,))

    def ctypes_stub(prog, il):
        # ctypes.f:36-47 (executed for real by make_golden_commu.py): iother becomes 0-based; the MPI datatypes it
        # also builds do not exist outside an MPI run
        pos = 1
        for _ in range(int(il[0])):
            il[pos + 2] -= 1
            pos += 4 + 2 * int(il[pos + 3])

    noop = lambda prog, *a: None  # noqa: E731
    mods = {"fhandle": 0, "iotype": "binary", "c_null_char": "\0", "nsynciofieldsreadgeombc": 0,
            "geomrestartstream": 0, "geombc_read": 1, "restart_read": 2, "cname2": lambda i: ".%d" % int(i),
            "mien": [], "mmat": [], "mxmudmi": [], "mieng": [], "mienb": [], "mibcb": [], "mbcb": [], "mmatb": [],
            "_comp_dtype": {"mien": np.int64, "mmat": np.int64, "mieng": np.int64, "mxmudmi": np.float64,
                            "mienb": np.int64, "mibcb": np.int64, "mmatb": np.int64, "mbcb": np.float64}}
    prog = Program([REF], modules=mods,
                   stubs={"phio_readheader": readheader, "phio_readdatablock": readdatablock, "phio_openfile": openfile,
                          "phio_closefile": closefile, "phio_constructname": construct, "phastaio_setfile": noop,
                          "posixio_setup": noop, "streamio_setup_read": noop, "syncio_setup_read": noop,
                          "phstr_appendint": noop, "phstr_appendstr": noop, "error": error, "ctypes": ctypes_stub,
                          "drvallreducemaxint": lambda prog, a, b: {1: lstep}})
    for fn in ("gensav.f", "genblkPosix.f", "gensvb.f", "genbkbPosix.f", "readnblk.f"):
        scan_functions(os.path.join(REF, fn))
        prog.load(os.path.join(REF, fn))
    G = prog.G
    # what input.f / the input.config plumbing has set before readnblk is called (compressible, no scalars, posix)
    G.update(myrank=rank, master=0, numpe=numpe, input_mode=0, ibksiz=ibksz, usingpetsc=0, svlsflag=0, istretchoutlet=0,
             iles=0, itwmod=0, nohomog=0, ideformwall=0, nsynciofiles=1, melcat=8, nsd=3, zero=0.0, one=1.0,
             npro=0, nshl=0, nshlb=0, nenbl=0, mattyp=0, ndofl=0, nsymdl=0, lcsyst=0, nenl=0, nfacel=0, maxsh=32,
             numnp=0, nshg=0, numel=0, numelb=0, nen=0, nelblk=0, nelblb=0, numpbc=0, ntopsh=0, nlwork=0, nshg0=0,
             lstep=0, nfath=0, nsonmax=0)
    # mnodeb / nenCat as common.f:104-109 sets them: nenCat(i,nsd): number of element nodes per category
    nencat = np.zeros((8, 3), dtype=np.int64, order="F")
    nencat[:, 0] = (2, 2, 2, 2, 3, 3, 3, 3)
    nencat[:, 1] = (3, 4, 3, 4, 6, 9, 6, 9)
    nencat[:, 2] = (4, 8, 6, 5, 10, 27, 18, 14)
    G["nencat"] = nencat
    G["matflg"] = np.zeros((6, 100), dtype=np.int64, order="F")
    G["impl"] = np.zeros(100, dtype=np.int64)
    G["lcblk"] = np.zeros((10, 50001), dtype=np.int64, order="F")
    G["lcblkb"] = np.zeros((10, 50001), dtype=np.int64, order="F")
    L = prog.call("readnblk")
    nelblk, nelblb = int(G["nelblk"]), int(G["nelblb"])
    out = {"scalars": np.array([int(G[k]) for k in ("numnp", "nshg", "numel", "numelb", "nen", "nelblk", "nelblb", "numpbc",
                                                   "nflow", "ndof", "ndofbc", "ndibcb", "ndbcb", "nsymdf", "nenb", "lstep",
                                                   "nlwork", "nshg0")], dtype=np.int64),
           "x": np.array(L["point2x"]), "nBC": np.array(L["nbc"], dtype=np.int32),
           "iBCtmp": np.array(L["ibctmp"], dtype=np.int32), "BCinp": np.array(L["bcinp"]),
           "iper": np.array(L["point2iper"], dtype=np.int32), "qold": np.array(L["qold"]),
           "acold": np.array(L["acold"]), "uold": np.array(L["uold"]),
           "lcblk": np.array(G["lcblk"][:, :nelblk + 1], dtype=np.int32),
           "lcblkb": np.array(G["lcblkb"][:, :nelblb + 1], dtype=np.int32),
           "ilwork": np.array(L["point2ilwork"], dtype=np.int32) if numpe > 1 else np.zeros(1, dtype=np.int32)}
    M = prog.M
    for i in range(nelblk):
        out["mien_%d" % i] = np.array(M["mien"][i].p, dtype=np.int32)
    for i in range(nelblb):
        out["mienb_%d" % i] = np.array(M["mienb"][i].p, dtype=np.int32)
        out["mibcb_%d" % i] = np.array(M["mibcb"][i].p, dtype=np.int32)
        out["mbcb_%d" % i] = np.array(M["mbcb"][i].p)
    assert state["opened"] == ["geombc", "restart"]
    return out


def generate():
    from phasta_b200 import phio
    out = {}
    for name in CASES:
        (params, tables, parts, states), ibksz, lstep, rank = build_case(name)
        y, ac = states[rank]
        with tempfile.TemporaryDirectory() as d:
            phio.write_geombc(parts[rank], d)
            phio.write_restart(d, rank, len(parts), lstep, y, ac)
            r = run_readnblk(d, ibksz, lstep, rank, len(parts))
        for k, v in r.items():
            out["%s_%s" % (name, k)] = v
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    new = generate()
    dst = os.path.join(HERE, "f77_readnblk.npz")
    if a.check:
        old = np.load(dst)
        assert set(old.files) == set(new), "fixture keys differ"
        for k in new:
            assert np.array_equal(old[k], new[k]), k
        print("f77_readnblk.npz reproduced bit for bit (%d arrays)" % len(new))
    else:
        np.savez_compressed(dst, **new)
        print("wrote", dst, len(new), "arrays")
