"""Drop-in for ``pyaxisymflow.elasto_kernels.div_tau``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    update_vorticity_from_solid_stress,
    update_vorticity_from_solid_stress_periodic,
)
