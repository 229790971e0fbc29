"""Drop-in for ``pyaxisymflow.elasto_kernels.advect_refmap_via_eno3``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    gen_advect_refmap_via_eno3,
    gen_advect_refmap_via_eno3_periodic,
)
