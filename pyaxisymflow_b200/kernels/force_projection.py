"""Drop-in for ``pyaxisymflow.kernels.force_projection``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    force_projection,
)
