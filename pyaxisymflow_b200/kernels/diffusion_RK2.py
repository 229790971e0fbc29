"""Drop-in for ``pyaxisymflow.kernels.diffusion_RK2``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    diffusion_RK2_unb,
    diffusion_RK2_periodic,
)
