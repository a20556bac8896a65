"""The halo exchange against the reference's own ctypes.f + commu.f, executed by f77np on top of an in-process MPI
emulation (tests/golden/make_golden_commu.py -> tests/golden/f77_commu.npz): 3 x-slab parts with split segments,
commu(...,'in ') then commu(...,'out') for n = 1, ndof, nflow^2, (nflow-1)*nsd."""
import os
import sys

import numpy as np
import pytest

from common import make_case, make_oracle, rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
from make_golden_commu import MAXSEG, NPARTS, NS, NX, NY, NZ  # noqa: E402


def _load():
    with np.load(os.path.join(GOLD, "f77_commu.npz")) as f:
        z = {k: f[k] for k in f.files}          # a plain dict: NpzFile must not be indexed from several threads
    case = make_case(NX, NY, NZ, nparts=NPARTS, bc="channel", max_seg=MAXSEG)
    return z, case


def test_ilwork_as_ctypes_leaves_it():
    """ctypes.f:47 turns iother 0-based; that is the in-memory ilwork every entry point takes (a29)"""
    z, case = _load()
    for p in case[2]:
        assert np.array_equal(z["ilwork_ctypes_%d" % p.rank], p.ilwork)
        # maxfront = the longest front of the part (ctypes.f:49-54): what sizes the halo buffers
        il, pos, fronts = p.ilwork, 1, []
        for _ in range(int(il[0])):
            nseg = int(il[pos + 3])
            fronts.append(int(sum(il[pos + 5 + 2 * s] for s in range(nseg))))
            pos += 4 + 2 * nseg
        assert int(z["maxfront_%d" % p.rank]) == max(fronts)


@pytest.mark.parametrize("n", NS)
def test_oracle_commu_matches_reference_fortran_bit_for_bit(n):
    z, case = _load()
    parts = case[2]
    o = make_oracle(case)
    w = [z["in_n%d_r%d" % (n, p.rank)].copy(order="F") for p in parts]
    o.commu(w, n, "in")
    for p in parts:
        assert np.array_equal(w[p.rank], z["afterin_n%d_r%d" % (n, p.rank)])
    o.commu(w, n, "out")
    for p in parts:
        assert np.array_equal(w[p.rank], z["afterout_n%d_r%d" % (n, p.rank)])
    # the middle part is master of one plane and slave of the other: both roles are in the fixture
    assert any(not np.array_equal(z["in_n%d_r1" % n], z["afterin_n%d_r1" % n]) for _ in (0,))
