"""Drop-in for ``pyaxisymflow.core.particles_to_mesh``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    particles_to_mesh_2D_unbounded_mp4,
    particles_to_mesh_2D_mp4,
)
