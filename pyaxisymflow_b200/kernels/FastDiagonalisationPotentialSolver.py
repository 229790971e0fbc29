"""Drop-in for ``pyaxisymflow.kernels.FastDiagonalisationPotentialSolver``; implemented in :mod:`pyaxisymflow_b200.fd` (four FP64 tensor-core GEMMs)."""
from ..fd import (  # noqa: F401
    FastDiagonalisationPotentialSolver,
)
