"""Drop-in for ``pyaxisymflow.core.particles_to_mesh`` (core/src/instantiate.yml:17-27).  The MP4 pair the configured
drivers call lives in :mod:`pyaxisymflow_b200.ops` (fused forms included); the rest of the family in
:mod:`pyaxisymflow_b200.particles`."""
from ..particles import P2M as _P2M
from ..particles import particles_to_mesh_1D_mp4  # noqa: F401

globals().update(_P2M)
from ..ops import (  # noqa: F401,E402  (the tuned MP4 entries win over the generic ones)
    particles_to_mesh_2D_unbounded_mp4,
    particles_to_mesh_2D_mp4,
)
