"""Drop-in for ``pyaxisymflow.kernels.periodic_boundary_ghost_comm``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    gen_periodic_boundary_ghost_comm,
    gen_periodic_boundary_ghost_comm_eta,
)
