"""Golden vectors for genBC1 (phSolver/common/genbc1.f) from the reference's own Fortran, executed by the
f77np interpreter: random attribute rows BCtmp(nshg,ndof+7) with every velocity code 0..7 (including rows
that take the "flip them" branches) -> BC(nshg,ndofBC).  Writes tests/golden/f77_genbc1.npz.

    python tests/golden/make_golden_genbc.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from f77np import Program, scan_functions  # noqa: E402

REF = "/root/reference/phSolver"


def main():
    prog = Program([os.path.join(REF, "common")], modules={}, stubs={})
    src = os.path.join(REF, "common", "genbc1.f")
    scan_functions(src)
    prog.load(src)
    n = 400
    rng = np.random.default_rng(20261017)
    BCtmp = np.asfortranarray(rng.uniform(-1.0, 1.0, size=(n, 12)))
    BCtmp[:, 0] = rng.uniform(0.9, 1.3, n)       # density
    BCtmp[:, 1] = rng.uniform(280.0, 320.0, n)   # temperature
    BCtmp[:, 2] = rng.uniform(0.9e5, 1.1e5, n)   # pressure
    BCtmp[:, 11] = 0.0
    code = np.arange(n) % 8
    iBC = (code << 3).astype(np.int64)
    iBC |= rng.integers(0, 2, n) * 1             # density
    iBC |= rng.integers(0, 2, n) * 2             # temperature
    iBC |= np.where((iBC & 1) == 0, rng.integers(0, 2, n) * 4, 0)   # pressure (not with density)
    # rows that take the flip branches of codes 3, 5, 6 (genbc1.f:36-43,87-94,133-140)
    for k in range(n):
        if k % 24 in (3, 5, 6):
            BCtmp[k, 3 if code[k] != 6 else 4] = 0.0
        if k % 48 in (27, 29, 30):
            BCtmp[k, {3: 8, 5: 9, 6: 9}[code[k]]] = 0.0
    G = prog.G
    G.update(nshg=n, ndof=5, ndofbc=6, nsd=3, nsclr=0)
    BC = np.zeros((n, 6), order="F")
    work = BCtmp.copy(order="F")
    with np.errstate(all="ignore"):
        prog.call("genbc1", work, iBC.copy(), BC)
    np.savez_compressed(os.path.join(HERE, "f77_genbc1.npz"), BCtmp=BCtmp, iBC=iBC.astype(np.int32), BC=BC)
    print("wrote f77_genbc1.npz; finite:", np.isfinite(BC).all(), " |BC| max", np.abs(BC[np.isfinite(BC)]).max())


if __name__ == "__main__":
    main()
