"""Drop-in for ``pyaxisymflow.pyst_kernels.advection_flux``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    gen_advection_flux_conservative_eno3_pyst_kernel,
    gen_advection_flux_non_conservative_eno3_pyst_kernel,
)
