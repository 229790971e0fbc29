"""Drop-in for ``pyaxisymflow.kernels.FastDiagonalisationStokesSolver``; implemented in :mod:`pyaxisymflow_b200.fd` (cosine / real-FFT z transforms + tridiagonal r sweeps on large grids, the four FP64 DMMA GEMMs of the eigen-decomposition on small ones)."""
from ..fd import (  # noqa: F401
    FastDiagonalisationStokesSolver,
)
