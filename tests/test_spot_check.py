"""The slab spot check used at benchmark size (oracle/spot_check.py), validated on the CPU: the oracle on slabs cut out
of a mesh against the oracle on the whole mesh standing in for the GPU.  Agreement to round-off is expected in the slab
interiors (same elements, same inputs), and the check must notice a perturbed value."""
import numpy as np
import pytest

from common import make_case, make_oracle
from oracle.spot_check import default_slabs, slab_check


class WholeMeshOracle:
    """duck-types the four reads slab_check makes of a PhastaGPU"""

    def __init__(self, case, flavour):
        self.o = make_oracle(case)
        self.p = self.o.parts[0]
        if flavour == "csr":
            self.o.genadj()
            self.o.ElmGMRs()
            self.colm, self.rowp = self.p.colm, self.p.rowp
        else:
            self.o.ElmGMRe()

    def get(self, what):
        return getattr(self.p, what)

    def get_egmass_range(self, e0, n):
        return self.p.EGmass[e0:e0 + n]

    def get_lhsk_range(self, k0, n):
        return self.p.lhsK[:, k0:k0 + n]


@pytest.mark.parametrize("topo,flavour", [("tet", "ebe"), ("tet", "csr"), ("mixed", "ebe"), ("mixed", "csr")])
def test_slabs_reproduce_the_whole_mesh(topo, flavour):
    nx, ny, nz = 14, 5, 4
    case = make_case(nx, ny, nz, bc="channel", topo=topo, wedge_layers=1, ibksiz=64)
    params, tables, parts, states = case
    g = WholeMeshOracle(case, flavour)
    plane = (ny + 1) * (nz + 1)
    s = slab_check(g, parts[0], params, tables, *states[0], plane, flavour=flavour, chunk=37, max_chunks=4)
    assert s["slabs"] == [[0, 5], [5, 10], [9, 14]] == [list(t) for t in default_slabs(nx)]
    # (not bit for bit: the slab's blocks of ibksiz elements start elsewhere, which moves a few sums by an ulp)
    assert s["res"] < 1e-13 and s["BDiag"] < 1e-13
    if flavour == "csr":
        assert s["csr_rows_bit_exact"] and s["lhsK"] < 1e-13 and s["blocks"] > 0
    else:
        assert s["EGmass"] < 1e-13 and s["elements"] > 0
        assert s["last_element_checked"] == parts[0].numel - 1      # the last element of the part is covered
    # the first and last slabs reach the domain faces: 4 + 2 + 4 planes
    assert s["nodes"] == 10 * plane


def test_a_wrong_value_is_noticed():
    case = make_case(10, 4, 3, bc="channel")
    params, tables, parts, states = case
    g = WholeMeshOracle(case, "ebe")
    plane = 5 * 4
    g.p.res[3 * plane + 7, 2] += 1e-6 * np.linalg.norm(g.p.res)
    g.p.EGmass[parts[0].numel // 2, 3, 4] += 1.0
    s = slab_check(g, parts[0], params, tables, *states[0], plane, slabs=[(0, 5), (2, 8), (5, 10)], flavour="ebe")
    assert s["res"] > 1e-10 and s["EGmass"] > 1e-10 and s["BDiag"] < 1e-13
