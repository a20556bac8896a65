"""Snapshot of the COMMON-block scalars the hot path reads
(phSolver/common/common.h:35-268; defaults from common/input.config and
common/common.f:104-109, common/input.f:30,163-182)."""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import numpy as np


def _rgas():
    Rh = 8.31441
    Msh = (2.8e-2, 3.2e-2)
    xN2, xO2 = 0.79, 0.21
    Rs = (Rh / Msh[0], Rh / Msh[1])
    return 1.0 / (xN2 / Rs[0] + xO2 / Rs[1])      # input.f:182


@dataclass
class SolverParams:
    # /genpar/ (common.h:184-189)
    ipord: int = 1
    idiff: int = 1                # input_fform.cc:706-711
    itau: int = 0                 # input.config:233
    iremoveStabTimeTerm: int = 0
    EntropyPressure: int = 0
    dtsfct: float = 1.0           # input.config:238
    taucfct: float = 1.0          # input.config:239
    # /solpar/, /incomp/
    iDC: int = 0
    Navier: int = 1
    Kspace: int = 50              # input.config:195-198
    nGMRES: int = 1
    minIters: int = 10
    # material (/matdat/ datmat, matflg; /mmatpar/)
    matflg2: int = 0              # 0 constant viscosity, else Sutherland
    matflg3: int = 0
    Rgas: float = _rgas()
    gamma: float = 1.4
    gamma1: float = 0.4
    pr: float = 0.72
    datmat121: float = 1.8e-5     # viscosity
    datmat221: float = 273.0
    datmat321: float = 110.4
    datmat131: float = 0.0
    epsM: float = math.sqrt(np.finfo(np.float64).eps)   # input.f:30
    temper: float = 1.0           # input.config:266
    # /timdat/ (per-step; itrPC.f:18-47): backward Euler defaults
    Dtgl: float = 1.0e4
    almi: float = 1.0
    alfi: float = 1.0
    gami: float = 1.0
    etol: float = 1.0e-3
    # per-call switches (itrdrv.f:456-457,511-512)
    lhs: int = 1
    iprec: int = 1
    # quadrature rule (input.config:253-254): 2 -> 4-pt tets, 3-pt tris
    intg: int = 2
    intgb: int = 2
    ibksiz: int = 64

    def as_dict(self):
        return asdict(self)


@dataclass
class IncompParams:
    """The COMMON scalars the incompressible assembly reads beyond SolverParams
    (incompressible/e3ivar.f, e3stab.f, e3res.f, e3lhs.f; defaults common/input.config:
    Flow Advection Form: Convective -> iconvflow=2 (:227, input_fform.cc:650-653),
    Tau Matrix: Diagonal-Shakib -> itau=0 (:233), viscous correction -> idiff=1 (:248),
    lumped mass fractions 0 (:250-251)); water-like defaults rho=1, mu=1e-3."""
    iconvflow: int = 2
    itau: int = 0
    idiff: int = 1
    ipord: int = 1
    lhs: int = 1
    matflg5: int = 0              # matflg(5,1): 0 none, 1 constant body force datmat(1:3,5,1)
    rho: float = 1.0              # datmat(1,1,1)
    rmu: float = 1.0e-3           # datmat(1,2,1)
    bf: tuple = (0.0, 0.0, 0.0)
    flmpl: float = 0.0
    flmpr: float = 0.0
    Delt: float = 1.0e-2          # Delt(itseq); Dtgl = 1/Delt (itrdrv.f)
    almi: float = 1.0
    alfi: float = 1.0
    gami: float = 1.0
    dtsfct: float = 1.0
    taucfct: float = 1.0
    # boundary integral (incompressible/e3b.f, e3bvar.f): viscous flux flag (/nomodule/ iviscflux), wall-model
    # switch whose |1| turns the Force integral on (/turbvari/ itwmod), the surface IDs whose nsrflist entry is 1
    # (/aerfrc/ nsrflist(0:MAXSURF): flux through / force on these surfaces is integrated)
    iviscflux: int = 1
    itwmod: int = 0
    surfaces: tuple = ()

    @property
    def Dtgl(self):
        return 1.0 / self.Delt

    def with_rhoinf(self, rhoinf):
        """generalized-alpha scalars of itrPC.f:18-26 (itrSetup)"""
        if 0.0 <= rhoinf <= 1.0:
            self.almi = (3.0 - rhoinf) / (1.0 + rhoinf) / 2.0
            self.alfi = 1.0 / (1.0 + rhoinf)
            self.gami = 0.5 + self.almi - self.alfi
        else:
            self.almi = self.alfi = self.gami = 1.0
        return self
