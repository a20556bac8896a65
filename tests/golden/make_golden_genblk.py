"""Pins SURVEY row a28 (genblk / gensav / lcblk / mien) to the reference: common/genblkPosix.f and common/gensav.f --
and the boundary variant common/genbkbPosix.f + gensvb.f (lcblkb, mienb, miBCB, mBCB) -- are
executed UNMODIFIED by f77np; the two phio calls they make (phio_readheader / phio_readdatablock, the `use phio`
interface of common/phio.f90) are bound to phasta_b200.phio.PhioFile reading a geombc file written by
phasta_b200.phio.write_geombc.  So the chain checked is: the file this repo writes -> the reference's own block
generator -> lcblk(10,nelblk+1) and mien(iblk)%p, which tests/test_phio.py compares bit for bit with what
phasta_b200.phio.read_geombc builds from the same file.

    python tests/golden/make_golden_genblk.py [--check]      -> tests/golden/f77_genblk.npz
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

from f77np import Program, scan_functions  # noqa: E402

REF = "/root/reference/phSolver/common"
# boundary blocks (genbkbPosix.f + gensvb.f): name -> (make_case arguments, IBKSZ)
BCASES = {
    "tet_bnd": (dict(nx=5, ny=4, nz=3, topo="tet", boundary=True, natural="mixed"), 32),
    "mixed_bnd": (dict(nx=6, ny=6, nz=4, topo="mixed", boundary=True, natural="mixed"), 20),   # tets, wedge tri + quad faces
    "hex_bnd": (dict(nx=4, ny=3, nz=3, topo="hex", boundary=True, natural="mixed"), 4096),
}
# name -> (make_case arguments, IBKSZ)
CASES = {
    "tet": (dict(nx=5, ny=4, nz=3, topo="tet"), 64),             # 360 tets: five full blocks and a ragged one
    "mixed": (dict(nx=6, ny=6, nz=4, topo="mixed"), 50),         # tets then wedges: two 'connectivity interior' blocks
    "hex_one_block": (dict(nx=4, ny=3, nz=3, topo="hex"), 4096),  # fewer elements than IBKSZ
}


def build_case(name):
    from common import make_case
    kw, ibksz = (CASES[name] if name in CASES else BCASES[name])
    kw = dict(kw)
    nx, ny, nz = kw.pop("nx"), kw.pop("ny"), kw.pop("nz")
    return make_case(nx, ny, nz, bc="channel", ibksiz=ibksz, **kw), ibksz


def run_genblk(path, ntopo, ibksz):
    """genblkPosix.f on the geombc file `path`: returns lcblk(10,nelblk+1) and the list of mien(iblk)%p"""
    from phasta_b200 import phio

    f = phio.PhioFile(path, "r")

    def readheader(prog, fh, phrase, ints, n, dtype, iotype):
        h = f.readheader(phrase.rstrip("\0").strip(), int(n), "integer")
        ints[:len(h)] = h

    def readdatablock(prog, fh, phrase, arr, n, dtype, iotype):
        data = f.readdatablock(phrase.rstrip("\0").strip(), int(n), "integer")
        arr.reshape(-1, order="F")[:int(n)] = data

    readheader.array_args = (2,)
    readdatablock.array_args = (2,)
    prog = Program([REF], modules={"fhandle": 0, "iotype": "binary", "c_null_char": "\0", "mien": [], "mmat": [], "mxmudmi": [], "mieng": [],
                                   "_comp_dtype": {"mien": np.int64, "mmat": np.int64, "mieng": np.int64,
                                                   "mxmudmi": np.float64}},
                   stubs={"phio_readheader": readheader, "phio_readdatablock": readdatablock})
    for fn in ("gensav.f", "genblkPosix.f"):
        scan_functions(os.path.join(REF, fn))
        prog.load(os.path.join(REF, fn))
    G = prog.G
    # what readnblk.f has set before it calls genblk (readnblk.f:150-215): nelblk = number of topology blocks in the
    # file, ndof, nsymdf; /elmpar/ nfacel is never assigned on this path and stays 0; maxsh, usingpetsc, numpe
    G.update(nelblk=ntopo, ndof=5, nsymdf=15, nfacel=0, maxsh=32, usingpetsc=0, numpe=1, myrank=0, npro=0, nshl=0,
             mattyp=0, ndofl=0, nsymdl=0, lcsyst=0, nenl=0)
    G["lcblk"] = np.zeros((10, 50001), dtype=np.int64, order="F")
    prog.call("genblkposix", ibksz)
    f.close()
    nelblk = int(G["nelblk"])
    return np.array(G["lcblk"][:, :nelblk + 1], dtype=np.int32, order="F"), \
        [np.array(p.p, dtype=np.int32, order="F") for p in prog.M["mien"][:nelblk]], \
        [np.array(p.p, dtype=np.int32) for p in prog.M["mmat"][:nelblk]]


def run_genbkb(path, ntopo, ibksz):
    """genbkbPosix.f + gensvb.f on the geombc file `path`: lcblkb(10,nelblb+1), mienb, miBCB, mBCB per block"""
    from phasta_b200 import phio

    f = phio.PhioFile(path, "r")

    def readheader(prog, fh, phrase, ints, n, dtype, iotype):
        h = f.readheader(phrase.rstrip("\0").strip(), int(n), "integer")
        ints[:len(h)] = h

    def readdatablock(prog, fh, phrase, arr, n, dtype, iotype):
        data = f.readdatablock(phrase.rstrip("\0").strip(), int(n), dtype.rstrip("\0").strip())
        arr.reshape(-1, order="F")[:int(n)] = data

    readheader.array_args = (2,)
    readdatablock.array_args = (2,)
    prog = Program([REF], modules={"fhandle": 0, "iotype": "binary", "c_null_char": "\0", "nsynciofieldsreadgeombc": 0,
                                   "mienb": [], "mibcb": [], "mbcb": [], "mmatb": [],
                                   "_comp_dtype": {"mienb": np.int64, "mibcb": np.int64, "mmatb": np.int64,
                                                   "mbcb": np.float64}},
                   stubs={"phio_readheader": readheader, "phio_readdatablock": readdatablock})
    for fn in ("gensvb.f", "genbkbPosix.f"):
        scan_functions(os.path.join(REF, fn))
        prog.load(os.path.join(REF, fn))
    G = prog.G
    # what readnblk.f has set before genbkb: nelblb = boundary topology blocks in the file, ndof, ndiBCB, ndBCB
    G.update(nelblb=ntopo, ndof=5, ndibcb=2, ndbcb=6, numpe=1, myrank=0, npro=0, nshl=0, nshlb=0, nenbl=0, mattyp=0,
             ndofl=0, lcsyst=0, nenl=0, zero=0.0)
    G["lcblkb"] = np.zeros((10, 50001), dtype=np.int64, order="F")
    prog.call("genbkbposix", ibksz)
    f.close()
    n = int(G["nelblb"])
    M = prog.M
    return (np.array(G["lcblkb"][:, :n + 1], dtype=np.int32, order="F"),
            [np.array(p.p, dtype=np.int32, order="F") for p in M["mienb"][:n]],
            [np.array(p.p, dtype=np.int32, order="F") for p in M["mibcb"][:n]],
            [np.array(p.p, dtype=np.float64, order="F") for p in M["mbcb"][:n]])


def generate():
    from phasta_b200 import phio
    out = {}
    for name in BCASES:
        (params, tables, parts, states), ibksz = build_case(name)
        with tempfile.TemporaryDirectory() as d:
            path = phio.write_geombc(parts[0], d)
            # one 'connectivity boundary' block per boundary kind in the file (genbkbPosix.f loops nelblb times)
            kinds = []
            for b in range(parts[0].nelblb):
                k = (int(parts[0].lcblkb[2, b]), int(parts[0].lcblkb[5, b]))
                if k not in kinds:
                    kinds.append(k)
            lcblkb, mienb, mibcb, mbcb = run_genbkb(path, len(kinds), ibksz)
        out["%s_lcblkb" % name] = lcblkb
        out["%s_nelblb" % name] = np.int32(len(mienb))
        for i in range(len(mienb)):
            out["%s_mienb_%d" % (name, i)] = mienb[i]
            out["%s_mibcb_%d" % (name, i)] = mibcb[i]
            out["%s_mbcb_%d" % (name, i)] = mbcb[i]
    for name in CASES:
        (params, tables, parts, states), ibksz = build_case(name)
        with tempfile.TemporaryDirectory() as d:
            path = phio.write_geombc(parts[0], d)
            ntopo = len(set(int(t) for t in parts[0].lcblk[2, :-1]))
            lcblk, mien, mmat = run_genblk(path, ntopo, ibksz)
        out["%s_lcblk" % name] = lcblk
        out["%s_nelblk" % name] = np.int32(len(mien))
        for i, (b, m) in enumerate(zip(mien, mmat)):
            out["%s_mien_%d" % (name, i)] = b
            assert (m == 1).all()                 # genblkPosix.f: `mater=1 ! all one material for now`
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    new = generate()
    dst = os.path.join(HERE, "f77_genblk.npz")
    if a.check:
        old = np.load(dst)
        assert set(old.files) == set(new), "fixture keys differ"
        for k in new:
            assert np.array_equal(old[k], new[k]), k
        print("f77_genblk.npz reproduced bit for bit (%d arrays)" % len(new))
    else:
        np.savez_compressed(dst, **new)
        print("wrote", dst, len(new), "arrays")
