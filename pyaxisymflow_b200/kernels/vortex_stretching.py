"""Drop-in for ``pyaxisymflow.kernels.vortex_stretching``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    vortex_stretching,
)
