"""Drop-in for ``pyaxisymflow.pyst_kernels.elementwise_ops``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    gen_elementwise_sum_pyst_kernel,
    gen_set_fixed_val_pyst_kernel,
)
