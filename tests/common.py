"""Shared case builders for the parity tests."""
import numpy as np

from phasta_b200 import SolverParams, make_box, make_state, make_tables, global_node_count


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def make_case(nx, ny, nz, *, nparts=1, bc="channel", periodic_z=True, ibksiz=64, rule=2, seed=1234,
              max_seg=0, boundary=False, natural="none", topo="tet", wedge_layers=1, **pkw):
    params = SolverParams(intg=rule, **pkw)
    tables = make_tables(rule, 2)
    parts = make_box(nx, ny, nz, nparts=nparts, bc=bc, periodic_z=periodic_z, ibksiz=ibksiz, seed=seed,
                     max_seg=max_seg, boundary=boundary, natural=natural, topo=topo, wedge_layers=wedge_layers)
    ng = global_node_count(nx, ny, nz)
    states = [make_state(p, ng, seed=seed) for p in parts]
    return params, tables, parts, states


def make_oracle(case, **flags):
    from oracle.oracle_py import Oracle
    params, tables, parts, states = case
    o = Oracle(parts, params, tables, states)
    if flags:
        o.set_flags(**flags)
    return o


from phasta_b200.mesh import nondimensional  # noqa: E402,F401
