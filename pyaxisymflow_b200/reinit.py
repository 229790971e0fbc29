"""Narrow-band level-set re-initialisation on the GPU.

Stands in for the drivers' third-party call ``skfmm.distance(phi, dx=dx, narrow=band)``
(``examples/SoftSphereStreaming/soft_sphere_streaming.py:196-199``): same arguments, same return
type (a ``numpy.ma.MaskedArray`` whose mask marks the cells the march did not reach), same
exceptions (``ValueError`` without a zero contour, ``RuntimeError`` on a negative discriminant).
scikit-fmm is not vendored in the reference, so parity is unpinned; the checker is the restatement
``oracle.axisym_oracle.fmm_distance``.  The work runs in ``csrc/reinit.cu``; there is no CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .device import DeviceField, Stage, make_grid, ptr, stream_ptr

MAX_DOUBLE = float(np.finfo(np.float64).max)


class NarrowBandReinit:
    """Workspace + entry for one grid shape (device-resident drivers keep one of these)."""

    def __init__(self, nr, nz):
        if not torch.cuda.is_available():
            raise _lib.AxbError("NarrowBandReinit needs a CUDA device (no CPU fallback)")
        self.nr, self.nz = int(nr), int(nz)
        self.nbytes = int(_lib.call("axb_reinit_workspace_bytes", self.nr, self.nz))
        self.work = torch.empty(self.nbytes + 256, dtype=torch.uint8, device="cuda")
        off = (-self.work.data_ptr()) % 256
        self._work_ptr = ctypes.c_void_p(self.work.data_ptr() + off)
        self.sweeps = 0

    def __call__(self, phi, dx, narrow, order=2, mask_out=None, ld=None):
        """in place on a CUDA float64 tensor ``phi`` (nr, nz); optional uint8 ``mask_out`` (1 = masked)."""
        if tuple(phi.shape) != (self.nr, self.nz):
            raise ValueError(f"phi has shape {tuple(phi.shape)}, workspace is for {(self.nr, self.nz)}")
        g = make_grid(self.nr, self.nz, int(phi.stride(0)) if ld is None else ld, dx)
        info = (ctypes.c_int * 2)(0, 0)
        _lib.call("axb_reinit_distance", ctypes.byref(g), ptr(phi), float(narrow), int(order), ptr(mask_out),
                  self._work_ptr, self.nbytes, info, stream_ptr())
        self.sweeps = int(info[0])
        status = int(info[1])
        if status & 1:
            raise ValueError("the array phi contains no zero contour (no zero level set)")
        if status & 2:
            raise RuntimeError("Negative discriminant in distance marcher quadratic.")
        if status & 4:
            raise RuntimeError(f"narrow-band re-initialisation: no fixed point after {self.sweeps} sweeps "
                               "(level set too rough inside the band)")
        return phi


def distance(phi, dx=1.0, narrow=0.0, order=2):
    """``skfmm.distance(phi, dx=dx, narrow=narrow, order=order)`` for 2-D float64 fields.

    NumPy input (parity mode: staged to the GPU and back) -> ``numpy.ma.MaskedArray``.
    ``DeviceField`` / CUDA tensor input -> ``(DeviceField distance, torch.bool mask)``; masked cells hold
    ``numpy.finfo(float).max`` like scikit-fmm's raw result.  ``narrow`` must be > 0 (the drivers always
    pass the band; a whole-domain march is not a hot-path operation)."""
    if np.ndim(dx) != 0:
        raise ValueError("anisotropic dx is not supported on this path (the drivers pass a scalar)")
    if not narrow > 0:
        raise ValueError("narrow must be > 0 on this path")
    host = isinstance(phi, np.ndarray)
    st = Stage()
    src = st.dev(phi)
    if src.ndim != 2:
        raise ValueError("phi must be 2-D")
    d = src.contiguous().clone()
    mask = torch.empty(d.shape, dtype=torch.uint8, device="cuda")
    NarrowBandReinit(*d.shape)(d, dx, narrow, order, mask_out=mask)
    mask = mask.bool()
    d[mask] = MAX_DOUBLE
    if host:
        return np.ma.MaskedArray(d.cpu().numpy(), mask=mask.cpu().numpy())
    return DeviceField(d), mask
