"""Drop-in for ``pyaxisymflow.kernels.smooth_Heaviside``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    smooth_Heaviside,
    smooth_Heaviside_sphere,
)
