"""Drop-in for ``pyaxisymflow.kernels.compute_vorticity_from_velocity``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    compute_vorticity_from_velocity_unb,
    compute_vorticity_from_velocity_periodic,
)
