"""FlowCondType: the slice of the reference's global `flow` (FlowCondition.f90:11-27) that the hot
path reads."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence


@dataclass
class FlowCondType:
    nu: float = 0.1                                    # flow%nu (Solidbody.f90:282)
    denIn: float = 1.0
    uvwIn: Sequence[float] = (0.0, 0.0, 0.0)
    shearRateIn: Sequence[float] = (0.0, 0.0, 0.0)
    velocityKind: int = 0                              # 0 shear, 2 oscillatory (FluidDomain.f90:1795-1799)
    volumeForceIn: Sequence[float] = (0.0, 0.0, 0.0)
    volumeForceAmp: float = 0.0
    volumeForceFreq: float = 0.0
    volumeForcePhi: float = 0.0
    Uref: float = 1.0
    ntolLBM: int = 1                                   # flow%ntolLBM (Solidbody.f90:345)
    dtolLBM: float = 1e-10                             # flow%dtolLBM
    numsubstep: int = 1                                # flow%numsubstep (LBMBlockComm.f90:326)
