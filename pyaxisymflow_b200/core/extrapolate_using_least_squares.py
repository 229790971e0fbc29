"""Drop-in for ``pyaxisymflow.core.extrapolate_using_least_squares``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    extrapolate_using_least_squares_till_first_order,
    extrapolate_using_least_squares_till_second_order,
)
