"""Drop-in for ``pyaxisymflow.kernels.compute_velocity_from_phi``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import compute_velocity_from_phi_unb  # noqa: F401
