"""Drop-in for ``pyaxisymflow.kernels.advect_particle``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    advect_vorticity_via_particles,
    advect_vorticity_via_particles_periodic,
    advect_vorticity_via_lattice_particles,
)
