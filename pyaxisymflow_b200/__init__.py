"""pyaxisymflow_b200 -- B200-native (sm_100a) per-timestep hot path of PyAxisymFlow.

The sub-packages ``kernels``, ``pyst_kernels``, ``elasto_kernels`` and ``core`` mirror the
reference's module tree, so a driver switches by replacing ``pyaxisymflow`` with
``pyaxisymflow_b200`` in its imports.  All arithmetic runs in the hand-written CUDA kernels of
``libaxisym_b200.so`` (C ABI in ``include/axisym_b200.h``); there is no CPU fallback.
"""
from ._lib import AxbError, LIB_PATH, launch_count  # noqa: F401
from .device import DeviceField  # noqa: F401

__version__ = "0.1.0"
