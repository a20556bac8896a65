"""Golden vectors for the halo exchange from the reference's own Fortran: common/ctypes.f builds the MPI derived
datatypes from ilwork, common/commu.f posts MPI_ISEND / MPI_IRECV / MPI_WAITALL and accumulates.  Both are executed
UNMODIFIED by f77np, one interpreter instance per rank, on top of a small in-process MPI emulation:

  * MPI_TYPE_INDEXED / MPI_TYPE_HVECTOR / MPI_TYPE_COMMIT build python descriptors of the typemaps;
  * an exchange is run in two passes over all ranks -- pass 1 executes every rank's commu up to MPI_WAITALL and
    keeps the bytes each MPI_ISEND would put on the wire (the typemap order of the datatype), pass 2 executes commu
    again and completes the receives at MPI_WAITALL, after which the routine's own accumulation loop runs.

Writes tests/golden/f77_commu.npz: per rank the input vectors and the results of commu(...,'in ') followed by
commu(...,'out') for n = 1, 5 (ndof), 25 (nflow*nflow), 12 ((nflow-1)*nsd) on a 3-part slab mesh with split segments,
and the ilwork array as ctypes.f leaves it (iother 0-based).

    python tests/golden/make_golden_commu.py [--check]
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

from f77np import Program, scan_functions  # noqa: E402

REF = "/root/reference/phSolver/common"
MPI_DOUBLE, MPI_LONG = 9001, 9002
NPARTS, NX, NY, NZ, MAXSEG = 3, 6, 3, 3, 7
NS = (1, 5, 25, 12)


class _WaitAll(Exception):
    pass


class World:
    def __init__(self):
        self.mail = {}       # (src, dst, tag) -> message (1-D float array in typemap order)
        self.phase = 1
        self.pending = []    # receives posted by the rank currently running (pass 2)


class Rank:
    """one MPI rank: its own interpreter (COMMON storage), its own datatype table"""

    def __init__(self, world, part):
        self.world, self.part, self.rank = world, part, part.rank
        self.types = {MPI_DOUBLE: ("base", 1), MPI_LONG: ("base", 1)}
        self.next_type = 100
        stubs = {"mpi_type_extent": self._extent, "mpi_type_indexed": self._indexed, "mpi_type_hvector": self._hvector,
                 "mpi_type_commit": lambda p, *a: None, "mpi_isend": self._isend, "mpi_irecv": self._irecv,
                 "mpi_waitall": self._waitall, "mpi_barrier": lambda p, *a: None, "error": self._error}
        self._isend.__func__.array_args = (0,)
        self._irecv.__func__.array_args = (0,)
        self.prog = Program([REF], modules={}, stubs=stubs)
        for f in ("ctypes.f", "commu.f"):
            scan_functions(os.path.join(REF, f))
            self.prog.load(os.path.join(REF, f))
        G = self.prog.G
        G.update(nshg=part.nshg, numnp=part.numnp, nsd=3, ndof=5, nflow=5, nlwork=part.nlwork, numpe=part.numpe,
                 myrank=part.rank, master=0, impistat=0, impistat2=0, maxtask=200, maxfront=0, nshg0=0,
                 mpi_double_precision=MPI_DOUBLE, mpi_long_long_int=MPI_LONG, mpi_comm_world=0, mpi_status_size=6)
        G["sevsegtype"] = np.zeros((200, 16), dtype=np.int64, order="F")
        # ilwork as it is read from geombc (iother 1-based); ctypes.f:47 makes it 0-based
        il = np.array(part.ilwork, dtype=np.int64)
        pos = 1
        for _ in range(int(il[0])):
            il[pos + 2] += 1
            pos += 4 + 2 * int(il[pos + 3])
        self.ilwork = il
        self.prog.call("ctypes", self.ilwork)

    # ---- datatype constructors ------------------------------------------------------------------
    def _error(self, prog, *a):
        raise RuntimeError("reference called error(%r)" % (a,))

    def _extent(self, prog, dtype, extent, ierr):
        return {1: 8}

    def _new(self, desc):
        self.next_type += 1
        self.types[self.next_type] = desc
        return self.next_type

    def _indexed(self, prog, count, lens, offs, old, new, ierr):
        segs = [(int(offs[i]), int(lens[i])) for i in range(int(count))]
        return {4: self._new(("indexed", segs, int(old)))}

    def _hvector(self, prog, count, blocklen, stride, old, new, ierr):
        assert int(blocklen) == 1
        return {4: self._new(("hvector", int(count), int(stride), int(old)))}

    def offsets(self, dtype):
        """element offsets (in doubles) of the typemap, in MPI's order"""
        d = self.types[int(dtype)]
        if d[0] == "base":
            return np.array([0])
        if d[0] == "indexed":
            base = self.offsets(d[2])
            return np.concatenate([o + k + base for o, ln in d[1] for k in range(ln)])
        base = self.offsets(d[3])
        assert d[2] % 8 == 0
        return np.concatenate([k * (d[2] // 8) + base for k in range(d[1])])

    # ---- point-to-point ----------------------------------------------------------------------------
    def _isend(self, prog, buf, count, dtype, dest, tag, comm, req, ierr):
        if self.world.phase != 1:
            return
        off = self.offsets(dtype)
        off = np.concatenate([off + c * 0 for c in range(int(count))]) if int(dtype) != MPI_DOUBLE else np.arange(int(count))
        self.world.mail[(self.rank, int(dest), int(tag))] = np.array(buf[off], dtype=np.float64)

    def _irecv(self, prog, buf, count, dtype, src, tag, comm, req, ierr):
        if self.world.phase != 2:
            return
        off = np.arange(int(count)) if int(dtype) == MPI_DOUBLE else self.offsets(dtype)
        self.world.pending.append((buf, off, int(src), int(tag)))

    def _waitall(self, prog, m, req, stat, ierr):
        if self.world.phase == 1:
            raise _WaitAll()
        for buf, off, src, tag in self.world.pending:
            msg = self.world.mail[(src, self.rank, tag)]
            assert msg.size == off.size, (msg.size, off.size)
            buf[off] = msg
        self.world.pending = []

    def commu(self, g, n, code):
        try:
            self.prog.call("commu", g, self.ilwork, int(n), code)
        except Exception as e:       # f77np re-raises with call-site context: look for our marker
            if not isinstance(e, _WaitAll) and "_WaitAll" not in repr(e) and not isinstance(e.__cause__, _WaitAll):
                raise


def exchange(world, ranks, vecs, n, code):
    world.mail.clear()
    world.phase = 1
    for r, g in zip(ranks, vecs):
        r.commu(g, n, code)
    world.phase = 2
    for r, g in zip(ranks, vecs):
        r.commu(g, n, code)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    from common import make_case, make_oracle
    case = make_case(NX, NY, NZ, nparts=NPARTS, bc="channel", max_seg=MAXSEG)
    parts = case[2]
    world = World()
    ranks = [Rank(world, p) for p in parts]
    out = {}
    for r, p in zip(ranks, parts):
        out["ilwork_ctypes_%d" % p.rank] = r.ilwork.astype(np.int32)      # after ctypes.f
        assert np.array_equal(r.ilwork, p.ilwork), "ctypes.f: iother must come out 0-based"
        out["maxfront_%d" % p.rank] = np.array(int(r.prog.G["maxfront"]))
    ok = True
    for n in NS:
        vecs = [np.asfortranarray(np.random.default_rng(1000 * n + p.rank).standard_normal((p.nshg, n))) for p in parts]
        for p, v in zip(parts, vecs):
            out["in_n%d_r%d" % (n, p.rank)] = v.copy(order="F")
        exchange(world, ranks, vecs, n, "in ")
        for p, v in zip(parts, vecs):
            out["afterin_n%d_r%d" % (n, p.rank)] = v.copy(order="F")
        exchange(world, ranks, vecs, n, "out")
        for p, v in zip(parts, vecs):
            out["afterout_n%d_r%d" % (n, p.rank)] = v.copy(order="F")
        if args.check:
            o = make_oracle(case)
            w = [out["in_n%d_r%d" % (n, p.rank)].copy(order="F") for p in parts]
            o.commu(w, n, "in")
            a = all(np.array_equal(w[p.rank], out["afterin_n%d_r%d" % (n, p.rank)]) for p in parts)
            o.commu(w, n, "out")
            b = all(np.array_equal(w[p.rank], out["afterout_n%d_r%d" % (n, p.rank)]) for p in parts)
            changed = any(not np.array_equal(out["in_n%d_r%d" % (n, p.rank)], out["afterin_n%d_r%d" % (n, p.rank)])
                          for p in parts)
            print("n=%2d  oracle == commu.f  'in' %s  'out' %s   (exchange changed data: %s)" % (n, a, b, changed))
            ok &= a and b and changed
    np.savez_compressed(os.path.join(HERE, "f77_commu.npz"), **out)
    if args.check:
        print("ALL OK" if ok else "MISMATCH")


if __name__ == "__main__":
    main()
