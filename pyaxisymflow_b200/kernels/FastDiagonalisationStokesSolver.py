"""Drop-in for ``pyaxisymflow.kernels.FastDiagonalisationStokesSolver``; implemented in :mod:`pyaxisymflow_b200.fd` (four FP64 tensor-core GEMMs)."""
from ..fd import (  # noqa: F401
    FastDiagonalisationStokesSolver,
)
