"""Drop-in for ``pyaxisymflow.kernels.update_baroclinic_vorticity``; implemented in :mod:`pyaxisymflow_b200.ops` on sm_100a kernels."""
from ..ops import (  # noqa: F401
    update_baroclinic_vorticity,
    update_baroclinic_vorticity_penal,
    update_baroclinic_vorticity_diff_penal,
)
