"""Drop-in for ``pyaxisymflow.core.mesh_to_particles`` (core/src/instantiate.yml:1-15); implemented in
:mod:`pyaxisymflow_b200.particles` on sm_100a kernels."""
from ..particles import (  # noqa: F401
    mesh_to_particles_1D_mp4,
    wrap_particles_around_1D_domain,
    wrap_particles_around_2D_domain,
)
from ..particles import M2P as _M2P

globals().update(_M2P)
