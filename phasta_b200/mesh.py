"""Synthetic PHASTA mesh parts: structured box, 6 Kuhn tets
per hex, slab partition along x with ilwork master/slave lists.

Every array uses the reference's in-memory layout (column-major, 1-based node
ids) so the same buffers feed the oracle, the C-ABI and a phastaIO writer:
  x(numnp,3), ien per block (npro,nshl), lcblk(10,nelblk+1)
  (phSolver/common/genblkPosix.f:62-72), iBC bits (compressible/bc3res.f:30-153),
  BC(nshg,ndofBC=ndof+1) (common/readnblk.f:171), iper (au1gmr.f:35-37),
  ilwork (common/commu.f:131-143, iother 0-based as after ctypes.f:47).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from itertools import permutations

import numpy as np

NFLOW = 5
NDOF = 5
NDOFBC = NDOF + 1


@dataclass
class MeshPart:
    """One partition (= one MPI rank of the reference / one GPU here)."""
    rank: int
    numpe: int
    nshg: int
    numnp: int
    numel: int
    x: np.ndarray                 # (numnp,3) F
    lcblk: np.ndarray             # (10, nelblk+1) F int32
    mien: list                    # per block (npro,nshl) F int32, 1-based
    iBC: np.ndarray               # (nshg,) int32
    BC: np.ndarray                # (nshg, NDOFBC) F
    iper: np.ndarray              # (nshg,) int32 1-based
    ilwork: np.ndarray            # (nlwork,) int32
    lcblkb: np.ndarray = None     # (10, nelblb+1)
    mienb: list = field(default_factory=list)
    miBCB: list = field(default_factory=list)
    mBCB: list = field(default_factory=list)
    gnode: np.ndarray = None      # local -> global node id (0-based), tests only
    gelem: np.ndarray = None      # local -> global element id (0-based)

    @property
    def nelblk(self):
        return self.lcblk.shape[1] - 1

    @property
    def nelblb(self):
        return 0 if self.lcblkb is None else self.lcblkb.shape[1] - 1

    @property
    def nlwork(self):
        return int(self.ilwork.size)

    def ien_all(self):
        """(numel, nshl) connectivity, 1-based, file order (single-topology meshes)."""
        return np.concatenate([np.asarray(b) for b in self.mien], axis=0)

    def elem_nodes(self):
        """list over elements (file order) of 0-based node-id arrays; any topology mix."""
        return [np.asarray(r) - 1 for b in self.mien for r in np.asarray(b)]


_KUHN = list(permutations(range(3)))


def _box_tets(nx, ny, nz, node_id):
    """Connectivity (6*nx*ny*nz, 4) of the Kuhn split, hex-major order.
    node_id(i,j,k) -> 0-based id (vectorised)."""
    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    nh = I.size
    ien = np.empty((nh, 6, 4), dtype=np.int64)
    for t, perm in enumerate(_KUHN):
        d = np.zeros((4, 3), dtype=np.int64)
        for s, ax in enumerate(perm):
            d[s + 1:, ax] += 1
        # parity of the permutation decides orientation; keep det>0 for
        # x = x4 + r(x1-x4) + s(x2-x4) + t(x3-x4)  (uniformP.c:18-35)
        verts = [node_id(I + d[v, 0], J + d[v, 1], K + d[v, 2]) for v in range(4)]
        e = np.array([d[0] - d[3], d[1] - d[3], d[2] - d[3]], dtype=float)
        if np.linalg.det(e) < 0:
            verts[0], verts[1] = verts[1], verts[0]
        for v in range(4):
            ien[:, t, v] = verts[v]
    return ien.reshape(nh * 6, 4)


# topology -> (lcsyst, nenl, nshl, nfacel)  (genblkPosix.f:62-72, common.h:111)
_TOPO = {"tet": (1, 4, 4, 4), "hex": (2, 8, 8, 6), "wedge": (3, 6, 6, 5)}


def _blocks(groups, ibksiz, ipord=1):
    """genblkPosix.f:52-96: per topology (file order = order of `groups`),
    consecutive runs of <= ibksiz elements.  groups: list of (topo, ien1)
    with ien1 (n, nshl) 1-based.  Returns lcblk(10,nelblk+1) and mien."""
    cols, mien = [], []
    first = 1
    for topo, ien1 in groups:
        lcsyst, nenl, nshl, nfacel = _TOPO[topo]
        n = ien1.shape[0]
        for a in range(0, n, ibksiz):
            b = min(a + ibksiz, n)
            # rows 6 and 9 as genblkPosix.f:66-69 leaves them: nfacel is never assigned on the read path
            # (COMMON /elmpar/ stays 0), nsymdl = nsymdf = ndof(ndof+1)/2 (readnblk.f:175)
            cols.append([first + a, 0, lcsyst, ipord, nenl, 0, 0, NDOF, NDOF * (NDOF + 1) // 2, nshl])
            mien.append(np.asfortranarray(ien1[a:b], dtype=np.int32))
        first += n
    cols.append([first, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    lcblk = np.asfortranarray(np.array(cols, dtype=np.int32).T)
    return lcblk, mien


def _box_hexes(I, J, K, node_id):
    """8-node hexes in HexShapeAndDrv order (newshape.cc:431-432): node a at
    (xi,eta,zeta) = n[a], x = i, y = j, z = k."""
    off = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    return np.stack([node_id(I + a, J + b, K + c) for a, b, c in off], axis=1)


def _box_wedges(I, J, K, node_id):
    """Two 6-node wedges per hex, prism axis along y (wall normal).  The
    triangles use the (0,0)-(1,1) diagonal of the x-z face, which is the
    diagonal the Kuhn tets put on that face, so wedge layers are conforming
    with the tets above them.  Node order (newshape.cc:735-740): 1-3 the
    zeta=-1 triangle {1-r-s, r, s}, 4-6 the zeta=+1 triangle; ordered so
    that det(dx/dxi) > 0."""
    tris = [((0, 0), (1, 1), (1, 0)), ((0, 0), (0, 1), (1, 1))]   # (di,dk); x_r cross x_s points to +y
    out = []
    for tri in tris:
        lo = [node_id(I + di, J, K + dk) for di, dk in tri]
        hi = [node_id(I + di, J + 1, K + dk) for di, dk in tri]
        out.append(np.stack(lo + hi, axis=1))
    return np.stack(out, axis=1).reshape(-1, 6)


# boundary-element kinds: lcsyst of the boundary block (genbkbPosix.f:52-59; a wedge with a quadrilateral boundary
# face is lcsyst 4, elmgmr.f:191), nenl, nenbl, nshl, nshlb, lnode (getbnodes, hierarchic.f:90-190, 0-based) and
# the proper rotations of the reference element that bring each of its faces of that kind onto lnode
_BKIND = {
    "hex": dict(lcsyst=2, nenl=8, nenbl=4, lnode=(0, 1, 2, 3),
                rots=((0, 1, 2, 3, 4, 5, 6, 7), (0, 4, 5, 1, 3, 7, 6, 2), (0, 3, 7, 4, 1, 2, 6, 5),
                      (1, 5, 6, 2, 0, 4, 7, 3), (2, 6, 7, 3, 1, 5, 4, 0), (4, 7, 6, 5, 0, 3, 2, 1))),
    "wedge3": dict(lcsyst=3, nenl=6, nenbl=3, lnode=(0, 1, 2),
                   rots=((0, 1, 2, 3, 4, 5), (3, 5, 4, 0, 2, 1))),
    "wedge4": dict(lcsyst=4, nenl=6, nenbl=4, lnode=(0, 3, 4, 1),
                   rots=((0, 1, 2, 3, 4, 5), (1, 2, 0, 4, 5, 3), (2, 0, 1, 5, 3, 4))),
}


def _natural_codes(face_gnodes, natural, seed, nshlb):
    """iBCB(nb,2) / BCB(nb,nshlb,6): "none" -> all fluxes floating; "mixed" -> deterministic sprinkle of
    mass-flux / pressure / traction / heat-flux codes + values, keyed on the GLOBAL ids of the face nodes."""
    nb = face_gnodes.shape[0]
    iBCB = np.zeros((nb, 2), dtype=np.int32)
    BCB = np.zeros((nb, nshlb, NDOF + 1))
    if natural == "mixed":
        gkey = np.sort(face_gnodes, axis=1)
        h = (gkey[:, 0] * 73856093 ^ gkey[:, 1] * 19349663 ^ gkey[:, 2] * 83492791) % 11
        r = np.random.default_rng(seed + 11)
        vals = r.uniform(0.5, 1.5, size=(11, 6)) * np.array([30.0, 1.0e5, 2.0, 2.0, 2.0, 50.0])
        for code, bit in ((1, 1), (2, 2), (3, 4), (4, 8), (5, 1 | 2), (6, 4 | 8)):
            s_ = h == code
            iBCB[s_, 0] = bit
            iBCB[s_, 1] = code
            BCB[s_, :, :] = vals[code][None, None, :]
        # genbkbPosix.f:86-100: values of unset codes are zeroed on read
        BCB[(iBCB[:, 0] & 1) == 0, :, 0] = 0.0
        BCB[(iBCB[:, 0] & 2) == 0, :, 1] = 0.0
        BCB[(iBCB[:, 0] & 4) == 0, :, 2:5] = 0.0
        BCB[(iBCB[:, 0] & 8) == 0, :, 5] = 0.0
    return iBCB, BCB


def _boundary_tets(ien0, x, on_boundary_planes):
    """each boundary tet is the volume tet re-ordered so that local nodes 1..3 are the boundary triangle
    with outward normal (v1 x v2, e3bvar.f:120-140) and node 4 is interior"""
    faces_of = ((1, 2, 3), (0, 3, 2), (0, 1, 3), (0, 2, 1))   # face opposite node k
    out = []
    for mask in on_boundary_planes:
        m = mask[ien0]                                         # (numel,4)
        for k in range(4):
            f = faces_of[k]
            sel = m[:, f[0]] & m[:, f[1]] & m[:, f[2]] & ~m[:, k]
            if not sel.any():
                continue
            t = ien0[sel]
            out.append(np.stack([t[:, f[0]], t[:, f[1]], t[:, f[2]], t[:, k]], axis=1))
    if not out:
        return None
    b = np.concatenate(out, axis=0)
    # orient: normal (b-a)x(c-a) must point away from d
    a_, b_, c_, d_ = (x[b[:, i]] for i in range(4))
    nrm = np.cross(b_ - a_, c_ - a_)
    flip = np.einsum("ij,ij->i", nrm, d_ - a_) > 0
    b[flip, 1], b[flip, 2] = b[flip, 2].copy(), b[flip, 1].copy()
    return b


def _boundary_rotated(ien0, kind, on_boundary_planes):
    """boundary hexes / wedges: the volume element rotated (a proper rotation of the reference element, so
    det(dx/dxi) stays positive) until the boundary face sits on lnode; the normals of e3bvar.f:139-152 then
    point outward for positively oriented elements"""
    K = _BKIND[kind]
    ln = list(K["lnode"])
    rest = [a for a in range(K["nenl"]) if a not in ln]
    out = []
    for mask in on_boundary_planes:
        for rot in K["rots"]:
            t = ien0[:, list(rot)]
            m = mask[t]
            sel = m[:, ln].all(axis=1) & ~m[:, rest].any(axis=1)
            if sel.any():
                out.append(t[sel])
    return np.concatenate(out, axis=0) if out else None


def _boundary_elements(groups0, x, on_boundary_planes, gnode, ibksiz, natural, seed):
    """Boundary elements of a part (genbkbPosix.f:47-123), one run of blocks per kind in the order tets,
    hexes, wedges with a triangular face, wedges with a quadrilateral face.
    groups0: list of (topo, ien0) with ien0 (n,nshl) 0-based; on_boundary_planes: boolean node masks."""
    runs = []
    for topo, ien0 in groups0:
        if topo == "tet":
            b = _boundary_tets(ien0, x, on_boundary_planes)
            if b is not None:
                runs.append((1, 4, 3, (0, 1, 2), b))
        else:
            for kind in (("hex",) if topo == "hex" else ("wedge3", "wedge4")):
                b = _boundary_rotated(ien0, kind, on_boundary_planes)
                if b is not None:
                    K = _BKIND[kind]
                    runs.append((K["lcsyst"], K["nenl"], K["nenbl"], K["lnode"], b))
    if not runs:
        return None, [], [], []
    runs.sort(key=lambda r: r[0])
    cols, mienb, miBCB, mBCB = [], [], [], []
    first = 1
    for lcsyst, nenl, nenbl, lnode, b in runs:
        nb = b.shape[0]
        ienb = (b + 1).astype(np.int32)
        iBCB, BCB = _natural_codes(gnode[b[:, list(lnode)]], natural, seed, nenbl)
        for a in range(0, nb, ibksiz):
            q = slice(a, min(a + ibksiz, nb))
            #            iel       -  lcsyst  ipord nenl  nenbl  mattyp ndofl nshl  nshlb   (genbkbPosix.f:103-114)
            cols.append([first + a, 0, lcsyst, 1, nenl, nenbl, 0, NDOF, nenl, nenbl])
            mienb.append(np.asfortranarray(ienb[q]))
            miBCB.append(np.asfortranarray(iBCB[q]))
            mBCB.append(np.asfortranarray(BCB[q]))
        first += nb
    cols.append([first, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    lcblkb = np.asfortranarray(np.array(cols, dtype=np.int32).T)
    return lcblkb, mienb, miBCB, mBCB


def make_box(nx, ny, nz, *, L=(1.0, 0.5, 0.5), nparts=1, ibksiz=64, perturb=0.15,
             bc="channel", periodic_z=True, seed=1234, max_seg=0, only_rank=None, boundary=False, natural="none",
             topo="tet", wedge_layers=1):
    """Build `nparts` MeshPart objects for an nx*ny*nz-hex box (6 tets/hex).

    bc: "channel"  x-min inflow (velocity code 7 + T), x-max pressure,
                   y walls no-slip isothermal, z periodic (or free)
        "none"     no essential BCs (iBC=0)
        "mixed"    channel + a few nodes with every velocity code 1..6 and
                   density BC, random slopes (exercises bc3* branches)
    max_seg: if >0 split ilwork segments to at most this length.
    only_rank: build (and return a 1-list with) just that rank's part.
    topo: "tet" (6 Kuhn tets per hex), "hex", "wedge" (2 wedges per hex), or
          "mixed" = `wedge_layers` hex layers at each y wall split into wedges,
          tets in between (BASELINE.json configs[2]); blocks are grouped by
          topology, tets first.
    boundary: also generate boundary elements (x-min, x-max, y walls; z faces
              unless periodic) with natural-BC codes per `natural`.
    """
    assert nx % nparts == 0 or nparts == 1, "nx must be divisible by nparts"
    rng = np.random.default_rng(seed)
    parts = []
    hx, hy, hz = L[0] / nx, L[1] / ny, L[2] / nz
    nyp, nzp = ny + 1, nz + 1

    def gid(i, j, k):
        return (i * nyp + j) * nzp + k

    # global iBC/BC pattern as functions of (i,j,k) so parts agree
    nxs = nx // nparts
    for p in range(nparts):
        if only_rank is not None and p != only_rank:
            continue
        i0, i1 = p * nxs, (p + 1) * nxs
        nxl = i1 - i0

        def lid(i, j, k, i0=i0):
            return ((i - i0) * nyp + j) * nzp + k

        I, J, K = np.meshgrid(np.arange(i0, i1 + 1), np.arange(nyp), np.arange(nzp),
                              indexing="ij")
        I, J, K = I.ravel(), J.ravel(), K.ravel()
        nn = I.size
        x = np.empty((nn, 3), order="F")
        X, Y, Z = I * hx, J * hy, K * hz
        if perturb:
            # interior-only smooth perturbation; vanishes on all faces so the
            # box stays a box and periodic planes stay congruent
            s = (np.sin(np.pi * X / L[0]) * np.sin(np.pi * Y / L[1]) *
                 np.sin(np.pi * Z / L[2]))
            X = X + perturb * hx * s * np.sin(7.0 * Y / L[1] + 3.0 * Z / L[2])
            Y = Y + perturb * hy * s * np.sin(5.0 * X / L[0] + 2.0 * Z / L[2])
            Z = Z + perturb * hz * s * np.sin(4.0 * X / L[0] + 6.0 * Y / L[1])
        x[:, 0], x[:, 1], x[:, 2] = X, Y, Z
        gnode = gid(I, J, K)

        nid = lambda a, b, c: lid(a + i0, b, c)   # noqa: E731
        hexid = ((np.arange(nxl)[:, None, None] + i0) * ny * nz +
                 np.arange(ny)[None, :, None] * nz + np.arange(nz)[None, None, :]).ravel()
        HI, HJ, HK = (a.ravel() for a in np.meshgrid(np.arange(nxl), np.arange(ny), np.arange(nz), indexing="ij"))
        ien0 = None
        if topo == "tet":
            ien0 = _box_tets(nxl, ny, nz, nid)
            groups = [("tet", ien0 + 1)]
            gelem = (hexid[:, None] * 6 + np.arange(6)[None, :]).ravel()
        elif topo == "hex":
            groups = [("hex", _box_hexes(HI, HJ, HK, nid) + 1)]
            gelem = hexid * 6
        elif topo == "wedge":
            groups = [("wedge", _box_wedges(HI, HJ, HK, nid) + 1)]
            gelem = (hexid[:, None] * 6 + np.arange(2)[None, :]).ravel()
        elif topo == "mixed":
            assert 2 * wedge_layers < ny, "wedge layers must leave tets in between"
            nearwall = (HJ < wedge_layers) | (HJ >= ny - wedge_layers)
            tets = _box_tets(nxl, ny, nz, nid).reshape(-1, 6, 4)[~nearwall].reshape(-1, 4)
            wdg = _box_wedges(HI[nearwall], HJ[nearwall], HK[nearwall], nid)
            groups = [("tet", tets + 1), ("wedge", wdg + 1)]
            gelem = np.concatenate([(hexid[~nearwall][:, None] * 6 + np.arange(6)[None, :]).ravel(),
                                    (hexid[nearwall][:, None] * 6 + np.arange(2)[None, :]).ravel()])
        else:
            raise ValueError("topo must be tet, hex, wedge or mixed")
        numel = int(sum(g[1].shape[0] for g in groups))
        lcblk, mien = _blocks(groups, ibksiz)

        iBC = np.zeros(nn, dtype=np.int32)
        BC = np.zeros((nn, NDOFBC), order="F")
        iper = np.arange(1, nn + 1, dtype=np.int32)
        if bc in ("channel", "mixed", "allcodes"):
            inflow = I == 0
            outflow = I == nx
            wall = (J == 0) | (J == ny)
            iBC[inflow] |= (7 << 3) | (1 << 1)
            iBC[outflow & ~wall] |= (1 << 2)
            iBC[wall] |= (7 << 3) | (1 << 1)
            # essential-BC values as itrBC applies them (itrbc.f:60-177): BC(:,1) pressure, BC(:,2)
            # temperature, BC(:,3:5) the velocity of code 7; no-slip isothermal walls
            BC[:, 0] = 1.0e5
            BC[:, 1] = 300.0
            BC[inflow, 2], BC[inflow, 3], BC[inflow, 4] = 30.0, 3.0, 1.5
            BC[wall, 2:5] = 0.0
            if periodic_z:
                slave = K == nz
                iBC[slave] |= (1 << 10)
                iper[slave] = lid(I[slave], J[slave], 0) + 1
        if bc == "allcodes":
            # every interior node carries a code: velocity codes 1..7 round-robin in the GLOBAL node id,
            # density / pressure+temperature on the eighth; random slopes (small meshes still see every
            # branch of bc3LHS / bc3Res / bc3BDg / itrBC)
            interior = (I > 0) & (I < nx) & (J > 0) & (J < ny) & (K > 0) & (K < nz)
            h = gnode % 9
            for code in range(1, 8):
                iBC[interior & (h == code)] |= (code << 3)
            iBC[interior & (h == 0)] |= 1
            iBC[interior & (h == 8)] |= (1 << 2) | (1 << 1)
            r = np.random.default_rng(seed + 11)
            slopes = r.uniform(-0.5, 0.5, size=((nx + 1) * nyp * nzp, NDOFBC))
            BC[:, :] = slopes[gnode]
            BC[:, 0] = 1.1 + 0.1 * slopes[gnode, 0]
        if bc == "mixed":
            # sprinkle the remaining velocity codes / density BC on interior
            # nodes, deterministic in the GLOBAL node id
            interior = (I > 0) & (I < nx) & (J > 0) & (J < ny) & (K > 0) & (K < nz)
            h = (gnode * 2654435761) % 97
            for code in range(1, 7):
                sel = interior & (h == code)
                iBC[sel] |= (code << 3)
            iBC[interior & (h == 7)] |= 1          # density
            iBC[interior & (h == 8)] |= (1 << 2) | (1 << 1)
            r = np.random.default_rng(seed + 7)
            slopes = r.uniform(-0.5, 0.5, size=((nx + 1) * nyp * nzp, NDOFBC))
            BC[:, :] = slopes[gnode]
            BC[:, 0] = 1.1 + 0.1 * slopes[gnode, 0]

        # ilwork: slab neighbours; lower rank is master
        tasks = []
        plane = nyp * nzp

        def segs(first):
            if max_seg and max_seg < plane:
                out = []
                a = first
                while a < first + plane:
                    ln = min(max_seg, first + plane - a)
                    out.append((a, ln))
                    a += ln
                return out
            return [(first, plane)]

        if nparts > 1:
            if p > 0:      # my left plane is owned by p-1: I am slave, send
                tasks.append((1000 + p - 1, 0, p - 1, segs(1)))
            if p < nparts - 1:   # my right plane: I am master, receive from p+1
                tasks.append((1000 + p, 1, p + 1, segs(nxl * plane + 1)))
        il = [len(tasks)]
        for tag, iacc, iother, sg in tasks:
            il += [tag, iacc, iother, len(sg)]
            for a, ln in sg:
                il += [a, ln]
        ilwork = np.array(il, dtype=np.int32)

        mp = MeshPart(rank=p, numpe=nparts, nshg=nn, numnp=nn, numel=numel,
                      x=x, lcblk=lcblk, mien=mien, iBC=iBC, BC=BC, iper=iper,
                      ilwork=ilwork, gnode=gnode, gelem=gelem)
        if boundary:
            planes = [J == 0, J == ny]
            if i0 == 0:
                planes.append(I == 0)
            if i1 == nx:
                planes.append(I == nx)
            if not (periodic_z and bc in ("channel", "mixed", "allcodes")):
                planes += [K == 0, K == nz]
            lcb, ienb, ibcb, bcb = _boundary_elements([(t, g1 - 1) for t, g1 in groups], x, planes, gnode, ibksiz,
                                                       natural, seed)
            if lcb is not None:
                mp.lcblkb, mp.mienb, mp.miBCB, mp.mBCB = lcb, ienb, ibcb, bcb
        parts.append(mp)
    return parts


def extract_slab(part: MeshPart, ia: int, ib: int, plane: int):
    """The x-slab of hex columns [ia, ib) of a box part as a stand-alone single-rank MeshPart (node planes ia..ib,
    renumbered from 1; essential BCs, periodicity and coordinates carried over; no boundary elements, no halo).
    `plane` = nodes per x-plane ((ny+1)(nz+1)).  Returns (sub, node_offset, elems) where elems[i] is the 0-based
    position, in `part`'s element order, of the slab's element i.  Used to spot-check meshes too large for a full
    CPU oracle pass: away from its two cut faces the slab sees exactly the elements the whole mesh does."""
    lo, hi = ia * plane, (ib + 1) * plane          # 0-based node range [lo, hi)
    assert 0 <= lo < hi <= part.nshg
    by_topo, order = {}, []
    for b, ien in enumerate(part.mien):
        ien = np.asarray(ien)
        lcsyst = int(part.lcblk[2, b])
        keep = ((ien.min(axis=1) - 1) >= lo) & ((ien.max(axis=1) - 1) < hi)
        if lcsyst not in by_topo:
            by_topo[lcsyst] = ([], [])
            order.append(lcsyst)
        by_topo[lcsyst][0].append(ien[keep] - lo)
        by_topo[lcsyst][1].append(int(part.lcblk[0, b]) - 1 + np.nonzero(keep)[0])
    name = {v[0]: k for k, v in _TOPO.items()}
    groups, elems = [], []
    for lcsyst in order:
        rows = np.concatenate(by_topo[lcsyst][0], axis=0)
        if rows.shape[0]:
            groups.append((name[lcsyst], rows))
            elems.append(np.concatenate(by_topo[lcsyst][1]))
    ibksiz = int(max(np.asarray(b).shape[0] for b in part.mien))
    lcblk, mien = _blocks(groups, ibksiz)
    nn = hi - lo
    iper = np.asarray(part.iper[lo:hi]) - lo
    assert iper.min() >= 1 and iper.max() <= nn, "periodic partners must lie in the same x-plane"
    sub = MeshPart(rank=0, numpe=1, nshg=nn, numnp=nn, numel=int(sum(g[1].shape[0] for g in groups)),
                   x=np.asfortranarray(part.x[lo:hi]), lcblk=lcblk, mien=mien,
                   iBC=np.ascontiguousarray(part.iBC[lo:hi]), BC=np.asfortranarray(part.BC[lo:hi]),
                   iper=iper.astype(np.int32), ilwork=np.array([0], dtype=np.int32),
                   gnode=None if part.gnode is None else part.gnode[lo:hi], gelem=None)
    return sub, lo, np.concatenate(elems)


def make_state(part: MeshPart, nglobal_nodes: int, seed=1234):
    """SURVEY 8(d) synthetic state, deterministic in the GLOBAL node id so
    partitioned and serial runs see identical fields.
    y = {u1,u2,u3,p,T} (nshg,5) F; ac same layout."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(-1.0, 1.0, size=(nglobal_nodes, 10))
    g = part.gnode
    y = np.empty((part.nshg, NDOF), order="F")
    ac = np.empty((part.nshg, NDOF), order="F")
    y[:, 0] = 30.0 * 1.0 + 3.0 * r[g, 0]
    y[:, 1] = 30.0 * 0.1 + 3.0 * r[g, 1]
    y[:, 2] = 30.0 * 0.05 + 3.0 * r[g, 2]
    y[:, 3] = 1.0e5 * (1.0 + 0.01 * r[g, 3])
    y[:, 4] = 300.0 * (1.0 + 0.01 * r[g, 4])
    ac[:, :] = 1.0e2 * r[g, 5:10]
    # periodic slaves carry the master's state (itrbc.f:170-178)
    m = part.iper - 1
    y[:, :] = y[m, :]
    ac[:, :] = ac[m, :]
    return y, ac


def make_smooth_state(part: MeshPart, L=(1.0, 0.5, 0.5), amp=1.0e-3):
    """Smooth subsonic channel state for time-stepping runs: Poiseuille-like
    u1(y) that vanishes on the y walls, small smooth u2/u3/p/T perturbations,
    ac = 0.  A function of the coordinates only, so parts agree on shared nodes;
    periodic in z.  Also writes the matching essential-BC values into part.BC
    (itrbc.f:60-177: BC(:,1) pressure, BC(:,2) temperature, BC(:,3:5) velocity
    of code 7) so that itrBC leaves the initial state unchanged."""
    x, yy, z = part.x[:, 0] / L[0], part.x[:, 1] / L[1], part.x[:, 2] / L[2]
    two_pi = 2.0 * np.pi
    prof = 4.0 * yy * (1.0 - yy)
    y = np.empty((part.nshg, NDOF), order="F")
    y[:, 0] = 30.0 * prof * (1.0 + 10 * amp * np.sin(two_pi * x) * np.cos(two_pi * z))
    y[:, 1] = 30.0 * amp * prof * np.sin(two_pi * x) * np.sin(two_pi * z)
    y[:, 2] = 30.0 * amp * prof * np.cos(two_pi * x) * np.sin(two_pi * z)
    y[:, 3] = 1.0e5 * (1.0 + amp * np.cos(two_pi * x) * np.cos(two_pi * z) * (0.5 + yy))
    y[:, 4] = 300.0 * (1.0 + amp * prof * np.sin(two_pi * x) * np.cos(two_pi * z))
    m = part.iper - 1
    y[:, :] = y[m, :]
    ac = np.zeros_like(y)
    part.BC[:, 0] = y[:, 3]
    part.BC[:, 1] = y[:, 4]
    part.BC[:, 2:5] = y[:, 0:3]
    return y, ac


def global_node_count(nx, ny, nz):
    return (nx + 1) * (ny + 1) * (nz + 1)


def nondimensional(case, mu=2.0e-4):
    """The same case in acoustic units (rho0 = c0 = T0 = L = 1): an exact
    similarity transform of state, BC values and parameters, so every variable
    is O(1).  The reference's matrix-free solver sizes its finite-difference
    interval for such variables (itrfdi.f:96-139: eGMRES from epsM alone);
    with SI magnitudes (p ~ 1e5) the perturbation y + eGMRES*u falls below one
    ulp of y and Au1MFG returns round-off."""
    import copy
    params, tables, parts, states = case
    P = copy.deepcopy(params)
    T0 = 300.0
    c0 = float(np.sqrt(P.gamma * P.Rgas * T0))
    rho0 = 1.0e5 / (P.Rgas * T0)
    p0 = rho0 * c0 * c0
    P.Rgas = P.Rgas * T0 / (c0 * c0)
    P.datmat121 = mu
    P.Dtgl = P.Dtgl / c0
    sy = np.array([1.0 / c0, 1.0 / c0, 1.0 / c0, 1.0 / p0, 1.0 / T0])
    nparts, nstates = [], []
    for mp, (y, ac) in zip(parts, states):
        q = copy.deepcopy(mp)
        q.BC[:, 0] *= 1.0 / p0            # pressure (itrbc.f:60-177); density BCs are not used by "channel"
        q.BC[:, 1] *= 1.0 / T0
        q.BC[:, 2:5] *= 1.0 / c0          # velocity of code 7
        for B in q.mBCB:                  # natural BCs (e3bvar.f:290-340): mass, pressure, traction(3), heat
            B[:, :, 0] *= 1.0 / (rho0 * c0)
            B[:, :, 1:5] *= 1.0 / p0
            B[:, :, 5] *= 1.0 / (rho0 * c0 ** 3)
        nparts.append(q)
        nstates.append((np.asfortranarray(y * sy), np.asfortranarray(ac * sy / c0)))
    return P, tables, nparts, nstates
