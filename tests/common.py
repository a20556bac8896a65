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
