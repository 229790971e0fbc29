"""Drop-in for ``pyaxisymflow.elasto_kernels.solid_sigma``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    solid_sigma,
    solid_sigma_periodic,
)
