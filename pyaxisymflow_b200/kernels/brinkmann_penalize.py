"""Drop-in for ``pyaxisymflow.kernels.brinkmann_penalize``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    brinkmann_penalize,
)
