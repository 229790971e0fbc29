"""Drop-in for ``pyaxisymflow.kernels.kill_boundary_vorticity_sine``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    kill_boundary_vorticity_sine_z,
    kill_boundary_vorticity_sine_r,
)
