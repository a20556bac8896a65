"""Parity at benchmark size (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

A full CPU oracle pass over a 4 M- or 32 M-tet part is minutes of one core and, for the EBE flavour, 13-100 GB of
host EGmass.  Instead the oracle runs on x-slabs cut out of the very part the GPU assembled
(phasta_b200.mesh.extract_slab: same coordinates, state, BC codes, periodicity).  Two planes in from its cut faces
a slab sees exactly the elements the whole mesh does -- qres (the global L2 projection of AsIq, asiq.f / qpbc.f)
is complete one plane in, so the elements between planes ia+1..ib-1 get the inputs they get in the whole mesh, and
res / BDiag / lhsK rows are complete for the nodes of planes ia+2..ib-2.  Slabs that touch a true domain face
(x-min of rank 0, x-max of the last rank) are compared right up to that face.  The first, middle and last slab of a
part cover element 0, the last tile, the last node and the last CSR block, i.e. the size_t indexing of the big
layouts.  Used by tests/test_gpu_at_size.py and bench.py's parity leg.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from phasta_b200.mesh import extract_slab
from . import oracle_py


def _rel(a, b):
    nb = float(np.linalg.norm(np.asarray(b).ravel()))
    return float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel())) / (nb if nb > 0 else 1.0)


def default_slabs(nxl, width=5):
    """first, middle and last slab of a part that is nxl hex columns long"""
    width = min(width, nxl)
    mid = max(0, min(nxl - width, nxl // 2 - width // 2))
    out = []
    for ia in (0, mid, nxl - width):
        if (ia, ia + width) not in out:
            out.append((ia, ia + width))
    return out


def _runs(idx, chunk):
    """split a sorted index array into contiguous runs no longer than chunk: [(start, n, offset_in_idx)]"""
    out, i, n = [], 0, len(idx)
    while i < n:
        j = i + 1
        while j < n and j - i < chunk and idx[j] == idx[j - 1] + 1:
            j += 1
        out.append((int(idx[i]), j - i, i))
        i = j
    return out


def slab_check(g, part, params, tables, y, ac, plane, slabs=None, *, flavour="ebe", max_chunks=6, chunk=8192,
               threads=None):
    """Compare what `g` (a PhastaGPU that has just assembled (y, ac) with the default step: lhs=1, iprec=1) holds in
    HBM with the oracle on every slab.  flavour "ebe": res, BDiag, EGmass tiles (ElmGMRe); "csr": res, BDiag,
    colm/rowp rows (bit-exact) and lhsK blocks (ElmGMRs; g.genadj() must have run).  Returns the largest relative
    L2 error per quantity over the slabs plus what was compared."""
    nxl = part.nshg // plane - 1
    slabs = slabs or default_slabs(nxl)
    first_rank, last_rank = part.rank == 0, part.rank == part.numpe - 1
    res_g, bd_g = g.get("res"), g.get("BDiag")
    if flavour == "csr":
        colm_g, rowp_g = np.asarray(g.colm), np.asarray(g.rowp)

    def one(sl):
        ia, ib = sl
        sub, off, elems = extract_slab(part, ia, ib, plane)
        o = oracle_py.Oracle([sub], params, tables, [(y[off:off + sub.nshg], ac[off:off + sub.nshg])])
        if flavour == "csr":
            o.genadj()
            o.ElmGMRs()
        else:
            o.ElmGMRe()
        op = o.parts[0]
        at_lo = ia == 0 and first_rank
        at_hi = ib == nxl and last_rank
        pa, pb = (ia if at_lo else ia + 2), (ib if at_hi else ib - 2)          # node planes that are complete
        ea, eb = (ia if at_lo else ia + 1), (ib if at_hi else ib - 1)          # elements between these planes
        assert pa <= pb and ea < eb, "slab too thin"
        n0, n1 = (pa - ia) * plane, (pb - ia + 1) * plane                       # slab-local node range
        out = {"slab": [ia, ib], "nodes": n1 - n0}
        out["res"] = _rel(res_g[off + n0:off + n1], op.res[n0:n1])
        out["BDiag"] = _rel(bd_g[off + n0:off + n1], op.BDiag[n0:n1])
        if flavour == "csr":
            ks0, ks1 = int(op.colm[n0]) - 1, int(op.colm[n1]) - 1
            kg0, kg1 = int(colm_g[off + n0]) - 1, int(colm_g[off + n1]) - 1
            out["csr_rows_bit_exact"] = bool(
                ks1 - ks0 == kg1 - kg0
                and np.array_equal(np.diff(op.colm[n0:n1 + 1]), np.diff(colm_g[off + n0:off + n1 + 1]))
                and np.array_equal(op.rowp[ks0:ks1] + off, rowp_g[kg0:kg1]))
            out["blocks"] = kg1 - kg0
            out["lhsK"] = _rel(g.get_lhsk_range(kg0, kg1 - kg0), op.lhsK[:, ks0:ks1]) if out["csr_rows_bit_exact"] else 1.0
        else:
            lo_n, hi_n = (ea - ia) * plane, (eb - ia + 1) * plane
            good = []
            pos = 0
            for ien in sub.mien:
                ien = np.asarray(ien)
                ok = ((ien.min(axis=1) - 1) >= lo_n) & ((ien.max(axis=1) - 1) < hi_n)
                good.append(pos + np.nonzero(ok)[0])
                pos += ien.shape[0]
            good = np.concatenate(good)                 # slab-local element ids whose inputs are complete
            order = np.argsort(elems[good], kind="stable")
            good = good[order]
            runs = _runs(elems[good], chunk)
            if len(runs) > max_chunks:                  # first, last and evenly spaced chunks in between
                pick = sorted(set(np.linspace(0, len(runs) - 1, max_chunks).round().astype(int).tolist()))
                runs = [runs[i] for i in pick]
            num = den = 0.0
            cnt = 0
            for start, n, at in runs:
                eg = g.get_egmass_range(start, n)
                ref = op.EGmass[good[at:at + n]]
                num += float(np.sum((eg - ref) ** 2))
                den += float(np.sum(ref ** 2))
                cnt += n
            out["EGmass"] = (num / den) ** 0.5 if den > 0 else 0.0
            out["elements"] = cnt
            out["element_range"] = [int(elems[good[0]]), int(elems[good[-1]])]
        return out

    with ThreadPoolExecutor(threads or len(slabs)) as ex:
        per = list(ex.map(one, slabs))
    keys = ["res", "BDiag"] + (["lhsK"] if flavour == "csr" else ["EGmass"])
    summary = {k: max(p[k] for p in per) for k in keys}
    if flavour == "csr":
        summary["csr_rows_bit_exact"] = all(p["csr_rows_bit_exact"] for p in per)
        summary["blocks"] = int(sum(p["blocks"] for p in per))
    else:
        summary["elements"] = int(sum(p["elements"] for p in per))
    summary["nodes"] = int(sum(p["nodes"] for p in per))
    summary["slabs"] = [p["slab"] for p in per]
    summary["last_element_checked"] = max((p.get("element_range", [0, 0])[1] for p in per), default=0)
    summary["per_slab"] = per
    return summary
