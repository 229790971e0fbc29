"""Drop-in for ``pyaxisymflow.elasto_kernels.extrapolate_eta_using_least_squares_unb``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    extrapolate_eta_with_least_squares,
)
