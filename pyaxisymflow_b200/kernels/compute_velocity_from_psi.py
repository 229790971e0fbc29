"""Drop-in for ``pyaxisymflow.kernels.compute_velocity_from_psi``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    compute_velocity_from_psi_unb,
    compute_velocity_from_psi_periodic,
)
