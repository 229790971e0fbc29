"""Drop-in for ``pyaxisymflow.kernels.compute_forces``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    compute_force_on_body,
)
