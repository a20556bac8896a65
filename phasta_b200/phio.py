"""phastaIO POSIX files (one part per file): geombc.dat.<rank+1>, restart.<step>.<rank+1>, numstart.dat.

Independent reader/writer of the reference's fixture format, so that a synthetic MeshPart can be written
as the files `readnblk` consumes and read back through the same header/key sequence
(BASELINE.json configs[0] "synthetic phastaIO geombc/restart").

Format facts followed (reference file:line):
  * header line  "<key phrase> : < nbytes > i1 i2 ...\\n", then `nbytes` raw bytes = data + '\\n'
    (phastaIO/phastaIO.cc:1624-1644 writeHeader, :230-305 readHeader);
  * key matching is a case- and blank-insensitive PREFIX match that stops at '?' in the searched
    phrase (phastaIO.cc:186-204 cscompare), searching forward from the current position and
    wrapping to the top of the file once (phastaIO.cc:239-296);
  * "byteorder magic number" carries the int 362436 (phastaIO.cc:268-272); '#' starts a comment;
  * file names geombc.dat.<rank+1>, restart.<step>.<rank+1> in <N>-procs_case/
    (phSolver/common/phio_posix.cc:13-17,97-106, readnblk.f:72-84,481-484), start step in numstart.dat;
  * key order and integer payloads of geombc: phSolver/common/readnblk.f:114-376,
    genblkPosix.f:42-56 (7 ints), genbkbPosix.f:47-83 (8 ints); restart: readnblk.f:491-549 with
    `restar` reordering {p,u,v,w,T} (file) <-> {u,v,w,p,T} (y) (compressible/restar.f:39-45,74-80).
Host-side file I/O only: nothing here is on the device path.
"""
from __future__ import annotations

import os

import numpy as np

from .mesh import MeshPart, NDOF, NDOFBC

MAGIC = 362436
_DT = {"integer": np.dtype("<i4"), "double": np.dtype("<f8")}
# (lcsyst, nshl) -> the phrase tail the reference's pre-processor writes after "connectivity interior"
_TOPO_NAME = {1: "linear tetrahedron", 2: "linear hexahedron", 3: "linear wedge", 4: "linear wedge quadface",
              5: "linear pyramid", 6: "linear pyramid triface"}
# nen -> nenb (readnblk.f:148-151 through nenCat); nenl, nenbl, nshlb per interior topology
_NENBL = {1: 3, 2: 4, 3: 3, 4: 4, 5: 4, 6: 3}


def cscompare(test: str, target: str) -> bool:
    """phastaIO.cc:186-204: `test` (the phrase asked for) against `target` (the token in the file)."""
    s1 = test.replace(" ", "").lower()
    s2 = target.replace(" ", "").lower()
    i = 0
    while i < len(s1) and i < len(s2) and s2[i] != "?" and s1[i] == s2[i]:
        i += 1
    return i >= len(s1) or s1[i] == "?"


class PhioFile:
    """One POSIX phasta file; mirrors phio_openfile / readheader / readdatablock / writeheader /
    writedatablock / closefile (phSolver/common/phIO.h, phio_posix.cc) for iotype "binary"."""

    def __init__(self, path: str, mode: str):
        assert mode in ("r", "w")
        self.path, self.mode = path, mode
        self.f = open(path, "rb" if mode == "r" else "wb")
        self.wrong_endian = False
        self._last = None           # (phrase, ndata, datatype) of the last header
        if mode == "w":
            self.f.write(b"# PHASTA Input File Version 2.0\n# Byte Order Magic Number : 362436 \n")
            self.writeheader("byteorder magic number", [1], 1, "integer")
            self.writedatablock("byteorder magic number", np.array([MAGIC], dtype=np.int32), "integer")

    # ------------------------------------------------------------------ write
    def writeheader(self, phrase, ints, ndata, datatype):
        size = _DT[datatype].itemsize * int(ndata) + (1 if ndata > 0 else 0)
        line = "%s : < %d > " % (phrase, size) + "".join("%d " % int(v) for v in ints) + "\n"
        self.f.write(line.encode())
        self._last = (phrase, int(ndata), datatype)

    def writedatablock(self, phrase, array, datatype):
        last = self._last
        if last is None or not cscompare(last[0], phrase):
            raise IOError("phio: header not consistent with data block (%r after %r)" % (phrase, last and last[0]))
        a = np.asarray(array, dtype=_DT[datatype]).ravel(order="F")
        if a.size != last[1] or datatype != last[2]:
            raise IOError("phio: header and datablock differ for %r" % phrase)
        self._last = None
        if a.size:
            self.f.write(a.tobytes())
            self.f.write(b"\n")

    # ------------------------------------------------------------------- read
    def _line(self):
        ln = self.f.readline(1023)
        return ln.decode("latin-1") if ln else None

    def readheader(self, phrase, expect, datatype="integer"):
        """-> list of `expect` ints, or None when the phrase is not in the file (the reference prints a
        warning and leaves the caller's integers untouched, phastaIO.cc:298-302)."""
        rewinds = 0
        line = self._line()
        if line is None:
            self.f.seek(0)
            rewinds += 1
            line = self._line()
        while rewinds < 2:
            if line and line[0] != "\n":
                text = line.split("#", 1)[0]
                if text and ":" in text:
                    token, rest = text.split(":", 1)
                    toks = rest.replace(",", " ").replace(";", " ").replace("<", " ").replace(">", " ").split()
                    if cscompare(phrase, token):
                        vals = [int(t) for t in toks[1:1 + expect]]
                        if len(vals) < expect:
                            raise IOError("phio: expected %d ints for %r" % (expect, phrase))
                        nbytes = int(toks[0])
                        isz = _DT[datatype].itemsize
                        self._last = (phrase, max(0, nbytes - 1) // isz if nbytes else 0, datatype)
                        return vals
                    if cscompare(token, "byteorder magic number"):
                        v = np.frombuffer(self.f.read(4), dtype="<i4")[0]
                        self.f.read(1)
                        self.wrong_endian = int(v) != MAGIC
                    else:
                        self.f.seek(int(toks[0]), os.SEEK_CUR)
            line = self._line()
            if line is None:
                self.f.seek(0)
                rewinds += 1
                line = self._line()
        self._last = None
        return None

    def readdatablock(self, phrase, n, datatype):
        if self._last is None or not cscompare(self._last[0], phrase):
            raise IOError("phio: data block %r read without its header" % phrase)
        self._last = None
        n = int(n)
        if n == 0:
            return np.zeros(0, dtype=_DT[datatype])
        dt = _DT[datatype]
        raw = self.f.read(dt.itemsize * n)
        self.f.read(1)     # the trailing newline
        a = np.frombuffer(raw, dtype=dt).copy()
        if self.wrong_endian:
            a = a.byteswap()
        if a.size != n:
            raise IOError("phio: short read of %r" % phrase)
        return a

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ------------------------------------------------------------------------------------------------
# essential-BC attribute form <-> solver form
def genBC1(BCtmp, iBC):
    """phSolver/common/genbc1.f:1-250 (nsd=3, nsclr=0): per-node attribute rows BCtmp(nshg,ndof+7) ->
    BC(nshg,ndofBC).  Statement order kept (flip test, normalisation, elimination)."""
    T = np.array(BCtmp, dtype=np.float64, order="F", copy=True)
    iBC = np.asarray(iBC).astype(np.int64)
    n = T.shape[0]
    BC = np.zeros((n, NDOFBC), order="F")
    c = (iBC >> 3) & 7
    m = (iBC & 1) != 0
    BC[m, 0] = T[m, 0]
    m = (iBC & 2) != 0
    BC[m, 1] = T[m, 1]
    m = (iBC & 4) != 0
    BC[m, 0] = T[m, 2]

    def single(code, piv, o1, o2):
        m = c == code
        tmp = T[m, 3] ** 2 + T[m, 4] ** 2 + T[m, 5] ** 2
        BC[m, 2] = tmp * T[m, 6] / T[m, 3 + piv]
        BC[m, 3] = T[m, 3 + o1] / T[m, 3 + piv]
        BC[m, 4] = T[m, 3 + o2] / T[m, 3 + piv]

    def flip(code, ia, ib):
        m = (c == code) & ((T[:, ia] == 0) | (T[:, ib] == 0))
        t = T[m, 3:7].copy()
        T[m, 3:7] = T[m, 7:11]
        T[m, 7:11] = t

    def normalise(m):
        for o in (3, 7):
            tmp = np.sqrt(T[m, o] ** 2 + T[m, o + 1] ** 2 + T[m, o + 2] ** 2)
            T[m, o] /= tmp
            T[m, o + 1] /= tmp
            T[m, o + 2] /= tmp
            T[m, o + 3] *= tmp

    single(1, 0, 1, 2)
    single(2, 1, 0, 2)
    # code 3: u1, u2 in terms of u3 (genbc1.f:36-76); 0-based columns: 3..6 first vector, 7..10 second
    flip(3, 3, 8)
    m = c == 3
    normalise(m)
    T[m, 3] = T[m, 8] * T[m, 3] - T[m, 4] * T[m, 7]
    T[m, 5] = T[m, 8] * T[m, 5] - T[m, 4] * T[m, 9]
    T[m, 6] = T[m, 8] * T[m, 6] - T[m, 4] * T[m, 10]
    BC[m, 2] = T[m, 6] / T[m, 3]
    BC[m, 3] = T[m, 5] / T[m, 3]
    T[m, 8] = T[m, 3] * T[m, 8]
    T[m, 9] = T[m, 3] * T[m, 9] - T[m, 7] * T[m, 5]
    T[m, 10] = T[m, 3] * T[m, 10] - T[m, 7] * T[m, 6]
    BC[m, 4] = T[m, 10] / T[m, 8]
    BC[m, 5] = T[m, 9] / T[m, 8]
    single(4, 2, 0, 1)
    # code 5: u1, u3 in terms of u2 (genbc1.f:87-130)
    flip(5, 3, 9)
    m = c == 5
    normalise(m)
    T[m, 3] = T[m, 9] * T[m, 3] - T[m, 5] * T[m, 7]
    T[m, 4] = T[m, 9] * T[m, 4] - T[m, 5] * T[m, 8]
    T[m, 6] = T[m, 9] * T[m, 6] - T[m, 5] * T[m, 10]
    BC[m, 2] = T[m, 6] / T[m, 3]
    BC[m, 3] = T[m, 4] / T[m, 3]
    T[m, 8] = T[m, 3] * T[m, 8] - T[m, 7] * T[m, 4]
    T[m, 9] = T[m, 3] * T[m, 9]
    T[m, 10] = T[m, 3] * T[m, 10] - T[m, 7] * T[m, 6]
    BC[m, 4] = T[m, 10] / T[m, 9]
    BC[m, 5] = T[m, 8] / T[m, 9]
    # code 6: u2, u3 in terms of u1 (genbc1.f:133-176)
    flip(6, 4, 9)
    m = c == 6
    normalise(m)
    T[m, 3] = T[m, 9] * T[m, 3] - T[m, 5] * T[m, 7]
    T[m, 4] = T[m, 9] * T[m, 4] - T[m, 5] * T[m, 8]
    T[m, 6] = T[m, 9] * T[m, 6] - T[m, 5] * T[m, 10]
    BC[m, 2] = T[m, 6] / T[m, 4]
    BC[m, 3] = T[m, 3] / T[m, 4]
    T[m, 7] = T[m, 4] * T[m, 7] - T[m, 8] * T[m, 3]
    T[m, 9] = T[m, 4] * T[m, 9]
    T[m, 10] = T[m, 4] * T[m, 10] - T[m, 8] * T[m, 6]
    BC[m, 4] = T[m, 10] / T[m, 9]
    BC[m, 5] = T[m, 7] / T[m, 9]
    m = c == 7
    for k in range(3):
        BC[m, 2 + k] = T[m, 6] * T[m, 3 + k]
    return BC


def active_bc_mask(iBC):
    """(nshg, ndofBC) mask of the BC entries the codes in iBC make the solver read
    (bc3res.f:30-153, itrbc.f:60-177): col 1 rho|p, col 2 T, cols 3..6 by velocity code."""
    iBC = np.asarray(iBC).astype(np.int64)
    m = np.zeros((iBC.size, NDOFBC), dtype=bool)
    c = (iBC >> 3) & 7
    m[:, 0] = (iBC & 5) != 0
    m[:, 1] = (iBC & 2) != 0
    m[:, 2:5] = (c != 0)[:, None]
    m[:, 5] = (c == 3) | (c == 5) | (c == 6)
    return m


def bcinp_from_BC(iBC, BC):
    """Attribute rows whose genBC1 image is `BC` (exact for codes 0 and 7 and the scalar values, to
    round-off for the slope codes 1..6): what a pre-processor would have written for this part."""
    iBC = np.asarray(iBC).astype(np.int64)
    n = iBC.size
    T = np.zeros((n, NDOF + 7), order="F")
    c = (iBC >> 3) & 7
    dens, pres = (iBC & 1) != 0, (iBC & 4) != 0
    T[dens, 0] = BC[dens, 0]
    T[:, 1] = np.where((iBC & 2) != 0, BC[:, 1], 0.0)
    T[pres, 2] = BC[pres, 0]
    m = c == 7
    T[m, 3:6] = BC[m, 2:5]
    T[m, 6] = 1.0
    # one constraint: u_piv = BC3 - BC4 u_o1 - BC5 u_o2
    for code, piv, o1, o2 in ((1, 0, 1, 2), (2, 1, 0, 2), (4, 2, 0, 1)):
        m = c == code
        T[m, 3 + piv] = 1.0
        T[m, 3 + o1] = BC[m, 3]
        T[m, 3 + o2] = BC[m, 4]
        T[m, 6] = BC[m, 2] / (1.0 + BC[m, 3] ** 2 + BC[m, 4] ** 2)
    # two constraints: (pa = BC3 - BC4 u_f, pb = BC5 - BC6 u_f); free component f
    for code, pa, pb, fr in ((3, 0, 1, 2), (5, 0, 2, 1), (6, 1, 2, 0)):
        m = c == code
        T[m, 3 + pa] = 1.0
        T[m, 3 + fr] = BC[m, 3]
        T[m, 6] = BC[m, 2] / (1.0 + BC[m, 3] ** 2)
        T[m, 7 + pb] = 1.0
        T[m, 7 + fr] = BC[m, 5]
        T[m, 10] = BC[m, 4] / (1.0 + BC[m, 5] ** 2)
    return T


# ------------------------------------------------------------------------------------------------
def case_dir(root, numpe):
    return os.path.join(root, "%d-procs_case" % numpe)


def _groups(part: MeshPart):
    """consecutive blocks of one (lcsyst, nshl, ipord) = one 'connectivity interior' topology block"""
    out = []
    for b in range(part.nelblk):
        key = tuple(int(v) for v in part.lcblk[[2, 3, 4, 9], b])
        if out and out[-1][0] == key:
            out[-1][1].append(np.asarray(part.mien[b]))
        else:
            out.append((key, [np.asarray(part.mien[b])]))
    return [(k, np.concatenate(v, axis=0)) for k, v in out]


def _groups_b(part: MeshPart):
    out = []
    for b in range(part.nelblb):
        key = tuple(int(v) for v in part.lcblkb[[2, 3, 4, 5, 8, 9], b])
        rec = (np.asarray(part.mienb[b]), np.asarray(part.miBCB[b]), np.asarray(part.mBCB[b]))
        if out and out[-1][0] == key:
            out[-1][1].append(rec)
        else:
            out.append((key, [rec]))
    return [(k, tuple(np.concatenate([r[i] for r in v], axis=0) for i in range(3))) for k, v in out]


def write_geombc(part: MeshPart, root: str) -> str:
    """Write geombc.dat.<rank+1> under <root>/<numpe>-procs_case/ with the key sequence readnblk.f reads.
    ilwork is stored as in the file (iother 1-based; ctypes.f:47 subtracts 1 on read)."""
    d = case_dir(root, part.numpe)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "geombc.dat.%d" % (part.rank + 1))
    groups, groups_b = _groups(part), _groups_b(part)
    has_bc = np.flatnonzero(np.asarray(part.iBC) != 0)
    numpbc = int(has_bc.size)
    nBC = np.zeros(part.nshg, dtype=np.int32)
    nBC[has_bc] = np.arange(1, numpbc + 1, dtype=np.int32)
    BCinp = bcinp_from_BC(part.iBC, part.BC)[has_bc]
    nen = max(int(k[2]) for k, _ in groups)
    with PhioFile(path, "w") as f:
        def scalar(key, v):
            f.writeheader(key, [v], 0, "integer")
        scalar("number of nodes", part.numnp)
        scalar("number of modes", part.nshg)
        scalar("number of interior elements", part.numel)
        scalar("number of boundary elements", sum(int(g[1][0].shape[0]) for g in groups_b))
        scalar("maximum number of element nodes", nen)
        scalar("number of interior tpblocks", len(groups))
        scalar("number of boundary tpblocks", len(groups_b))
        scalar("number of nodes with Dirichlet BCs", numpbc)
        scalar("number of shape functions", max(int(k[3]) for k, _ in groups))
        if part.numpe > 1:
            il = np.array(part.ilwork, dtype=np.int32)
            pos = 1
            for _ in range(int(il[0])):
                il[pos + 2] += 1
                pos += 4 + 2 * int(il[pos + 3])
            scalar("size of ilwork array", il.size)
            f.writeheader("ilwork", [il.size], il.size, "integer")
            f.writedatablock("ilwork", il, "integer")
        f.writeheader("co-ordinates", [part.numnp, 3], part.numnp * 3, "double")
        f.writedatablock("co-ordinates", part.x, "double")
        for (lcsyst, ipord, nenl, nshl), ien in groups:
            key = "connectivity interior " + _TOPO_NAME[lcsyst]
            nenbl = _NENBL[lcsyst]
            f.writeheader(key, [ien.shape[0], nenl, ipord, nshl, nenbl, nenbl, lcsyst], ien.size, "integer")
            f.writedatablock(key, ien, "integer")
        f.writeheader("bc mapping array", [part.nshg], part.nshg, "integer")
        f.writedatablock("bc mapping array", nBC, "integer")
        f.writeheader("bc codes array", [numpbc], numpbc, "integer")
        f.writedatablock("bc codes array", np.asarray(part.iBC)[has_bc], "integer")
        f.writeheader("boundary condition array", [BCinp.size], BCinp.size, "double")
        f.writedatablock("boundary condition array", BCinp, "double")
        f.writeheader("periodic masters array", [part.nshg], part.nshg, "integer")
        # the file holds 0 for "not periodic"; readnblk keeps it and the solver maps 0 -> self (genini/perprep)
        iper = np.asarray(part.iper, dtype=np.int32)
        f.writedatablock("periodic masters array", np.where(iper == np.arange(1, part.nshg + 1), 0, iper), "integer")
        for (lcsyst, ipord, nenl, nenbl, nshl, nshlb), (ienb, ibcb, bcb) in groups_b:
            tail = _TOPO_NAME[lcsyst]
            n = ienb.shape[0]
            hdr = [n, nenl, ipord, nshl, nshlb, nenbl, lcsyst, NDOF + 1]
            f.writeheader("connectivity boundary " + tail, hdr, ienb.size, "integer")
            f.writedatablock("connectivity boundary " + tail, ienb, "integer")
            f.writeheader("nbc codes " + tail, hdr, n * 2, "integer")
            f.writedatablock("nbc codes " + tail, ibcb, "integer")
            # the file stores one value row per element (genbkbPosix.f:60-63); gensvb spreads it over nshlb
            f.writeheader("nbc values " + tail, hdr, n * (NDOF + 1), "double")
            f.writedatablock("nbc values " + tail, np.asarray(bcb)[:, 0, :], "double")
    return path


def _blocked(ien, ibksiz):
    return [np.asfortranarray(ien[n:n + ibksiz]) for n in range(0, ien.shape[0], ibksiz)]


def read_geombc(root: str, rank: int, numpe: int, ibksiz: int) -> MeshPart:
    """readnblk.f:114-376 + genblkPosix.f + genbkbPosix.f + geniBC (genibc.f:17-19) + genBC/genBC1 +
    ctypes.f:47, producing the in-memory arrays the solver entry points take."""
    path = os.path.join(case_dir(root, numpe), "geombc.dat.%d" % (rank + 1))
    with PhioFile(path, "r") as f:
        def scalar(key):
            v = f.readheader(key, 1)
            if v is None:
                raise IOError("phio: %r not found in %s" % (key, path))
            return v[0]
        numnp = scalar("number of nodes")
        nshg = scalar("number of modes")
        numel = scalar("number of interior elements")
        scalar("number of boundary elements")
        scalar("maximum number of element nodes")
        ntp = scalar("number of interior tpblocks")
        ntpb = scalar("number of boundary tpblocks")
        numpbc = scalar("number of nodes with Dirichlet BCs")
        scalar("number of shape functions")
        if numpe > 1:
            nlwork = scalar("size of ilwork array")
            f.readheader("ilwork", 1)
            ilwork = f.readdatablock("ilwork", nlwork, "integer").astype(np.int32)
            pos = 1                                     # ctypes.f:36-47
            for _ in range(int(ilwork[0])):
                ilwork[pos + 2] -= 1
                pos += 4 + 2 * int(ilwork[pos + 3])
        else:
            ilwork = np.array([0], dtype=np.int32)
        numnp = f.readheader("co-ordinates", 2, "double")[0]
        x = np.asfortranarray(f.readdatablock("co-ordinates", numnp * 3, "double").reshape((numnp, 3), order="F"))
        # genblkPosix.f:36-105
        cols, mien, iel = [], [], 1
        for _ in range(ntp):
            neltp, nenl, ipordl, nshl, _, _, lcsyst = f.readheader("connectivity interior?", 7)
            ien = f.readdatablock("connectivity interior?", neltp * nshl, "integer").reshape((neltp, nshl), order="F")
            for blk in _blocked(ien.astype(np.int32), ibksiz):
                cols.append([iel, 0, lcsyst, ipordl, nenl, 0, 0, NDOF, NDOF * (NDOF + 1) // 2, nshl])
                mien.append(blk)
                iel += blk.shape[0]
        cols.append([iel] + [0] * 9)
        lcblk = np.asfortranarray(np.array(cols, dtype=np.int32).T)
        f.readheader("bc mapping array", 1)
        nBC = f.readdatablock("bc mapping array", nshg, "integer")
        f.readheader("bc codes array", 1)
        iBCtmp = f.readdatablock("bc codes array", numpbc, "integer")
        nin = f.readheader("boundary condition array", 1, "double")[0]
        BCinp = f.readdatablock("boundary condition array", nin, "double")
        f.readheader("periodic masters array", 1)
        iper = f.readdatablock("periodic masters array", nshg, "integer").astype(np.int32)
        # genbkbPosix.f:41-131
        colsb, mienb, miBCB, mBCB, iel = [], [], [], [], 1
        for _ in range(ntpb):
            neltp, nenl, ipordl, nshl, nshlb, nenbl, lcsyst, _ = f.readheader("connectivity boundary?", 8)
            ienb = f.readdatablock("connectivity boundary?", neltp * nshl, "integer").reshape((neltp, nshl), order="F")
            f.readheader("nbc codes?", 8)
            ibcb = f.readdatablock("nbc codes?", neltp * 2, "integer").reshape((neltp, 2), order="F")
            f.readheader("nbc values?", 8, "double")
            bcb = f.readdatablock("nbc values?", neltp * (NDOF + 1), "double").reshape((neltp, NDOF + 1), order="F").copy()
            c = ibcb[:, 0]
            bcb[(c & 1) == 0, 0] = 0.0                  # genbkbPosix.f:86-100
            bcb[(c & 2) == 0, 1] = 0.0
            bcb[(c & 8) == 0, 5] = 0.0
            bcb[(c & 4) == 0, 2:5] = 0.0
            for n in range(0, neltp, ibksiz):
                npro = min(ibksiz, neltp - n)
                colsb.append([iel, 0, lcsyst, ipordl, nenl, nenbl, 0, NDOF, nshl, nshlb])
                mienb.append(np.asfortranarray(ienb[n:n + npro].astype(np.int32)))
                miBCB.append(np.asfortranarray(ibcb[n:n + npro].astype(np.int32)))
                mBCB.append(np.asfortranarray(np.repeat(bcb[n:n + npro, None, :], nshlb, axis=1)))   # gensvb.f
                iel += npro
    # geniBC (genibc.f:17-19), genBC (genbc.f:20-25) + genBC1
    iBC = np.zeros(nshg, dtype=np.int32)
    BCtmp = np.zeros((nshg, NDOF + 7), order="F")
    if numpbc:
        sel = nBC != 0
        iBC[sel] = iBCtmp[nBC[sel] - 1]
        nsec = nin // numpbc
        BCtmp[sel, :] = BCinp.reshape((numpbc, nsec), order="F")[nBC[sel] - 1, :NDOF + 7]
    BC = genBC1(BCtmp, iBC)
    iper = np.where(iper == 0, np.arange(1, nshg + 1, dtype=np.int32), iper).astype(np.int32)
    part = MeshPart(rank=rank, numpe=numpe, nshg=nshg, numnp=numnp, numel=numel, x=x, lcblk=lcblk, mien=mien,
                    iBC=iBC, BC=BC, iper=iper, ilwork=ilwork)
    if colsb:
        colsb.append([iel] + [0] * 9)
        part.lcblkb = np.asfortranarray(np.array(colsb, dtype=np.int32).T)
        part.mienb, part.miBCB, part.mBCB = mienb, miBCB, mBCB
    return part


# ------------------------------------------------------------------------------------------------
def write_restart(root, rank, numpe, lstep, y, ac=None) -> str:
    """restart.<lstep>.<rank+1>: "solution" and "time derivative of solution" (nshg, ndof, lstep) in the
    file's {p,u,v,w,T} order (restar.f:74-80, new_interface.c:192-300) + numstart.dat (readnblk.f:72-80)."""
    d = case_dir(root, numpe)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "restart.%d.%d" % (lstep, rank + 1))
    perm = [3, 0, 1, 2, 4]
    with PhioFile(path, "w") as f:
        for key, a in (("solution", y), ("time derivative of solution", ac)):
            if a is None:
                continue
            a = np.asarray(a)
            f.writeheader(key, [a.shape[0], a.shape[1], lstep], a.size, "double")
            f.writedatablock(key, np.asfortranarray(a[:, perm]), "double")
    if rank == 0:
        with open(os.path.join(d, "numstart.dat"), "w") as g:
            g.write("%d\n" % lstep)
    return path


def read_restart(root, rank, numpe, nshg):
    """-> (y, ac, lstep) as `restar('in')` leaves them (restar.f:39-45); missing fields follow
    readnblk.f:516-549 (solution zero with p = T = 1; time derivative zero)."""
    d = case_dir(root, numpe)
    lstep = int(open(os.path.join(d, "numstart.dat")).read().split()[0])
    path = os.path.join(d, "restart.%d.%d" % (lstep, rank + 1))
    inv = [1, 2, 3, 0, 4]
    out = []
    with PhioFile(path, "r") as f:
        for key in ("solution", "time derivative of solution"):
            h = f.readheader(key, 3)
            if h is None or h[0] == 0:
                q = np.zeros((nshg, NDOF), order="F")
                if key == "solution":
                    q[:, 0] = 1.0
                    q[:, NDOF - 1] = 1.0
            else:
                if h[0] != nshg:
                    raise IOError("restar: nshg %d in %s, expected %d" % (h[0], path, nshg))
                lstep = h[2]
                q = f.readdatablock(key, h[0] * h[1], "double").reshape((h[0], h[1]), order="F")[:, :NDOF]
            out.append(np.asfortranarray(q[:, inv]))
    return out[0], out[1], lstep
