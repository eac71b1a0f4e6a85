"""plugin_navierstokes_b200 -- B200-native (sm_100a) FV1 / FVCR defect + Jacobian assembly of the incompressible
Navier-Stokes system behind the UG4 NavierStokes plugin's element-disc surface.

csrc/      CUDA kernels + the C ABI (include/nsb200.h) -> libnsb200.so (built in-tree by build.py)
disc.py    host-side mirror of NavierStokesFV1 / NavierStokesFVCR (names, setters, errors of the reference)
meshgen.py synthetic grids / states of the BASELINE.json configurations
partition.py  element partition + interface lists for the multi-GPU path
cr_reorder.py host mirror of OrderCRCuthillMcKee (FVCR dof ordering, solver-side preprocessing)
cr_ilut.py    host mirror of the CRILUT preconditioner (reference algorithm on the host, not a device solver)
"""
from . import _capi as capi                                   # noqa: F401
from .disc import (NavierStokes, NavierStokesFV1, NavierStokesFVCR, UGError,            # noqa: F401
                   CreateNavierStokesUpwind, CreateNavierStokesStabilization,
                   NavierStokesNoUpwind, NavierStokesFullUpwind, NavierStokesSkewedUpwind,
                   NavierStokesLinearProfileSkewedUpwind, NavierStokesPositiveUpwind, NavierStokesRegularUpwind,
                   NavierStokesFIELDSStabilization, NavierStokesFLOWStabilization,
                   NavierStokesFV1WithoutStabilization, NavierStokesWall, NavierStokesInflowFV1, NavierStokesNoNormalStressOutflowFV1,
                   NavierStokesNoNormalStressOutflow, FV1SmagorinskyTurbViscData, DiscConstraintFVCR, ThetaTimeStep)
from .cr_reorder import OrderCRCuthillMcKee                    # noqa: F401,E402
from .cr_ilut import CRILUTPreconditioner as CRILUT             # noqa: F401,E402
