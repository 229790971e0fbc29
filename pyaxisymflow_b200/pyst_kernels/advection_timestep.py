"""Drop-in for ``pyaxisymflow.pyst_kernels.advection_timestep``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel,
    gen_advection_timestep_euler_forward_non_conservative_eno3_pyst_kernel,
)
