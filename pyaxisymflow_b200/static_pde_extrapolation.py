"""``StaticPDEExtrapolation`` of ``examples/PeriodicSoftSlab/bounded_static_PDE_extrapolation.py:5-235`` on sm_100a
kernels (``csrc/pde_extrap.cu``): same constructor, same ``extrapolate(eta, phi)``, same in-place effect on ``eta``
(and, for periodic use, on the ghost columns of ``phi``), same ``ValueError``s.

The bounding box of the solid + extrapolation band is found on the device (one 4-integer read), the box is cut out
of the fields, and the set-up pass and the two Jacobi solves run as CUDA kernels; the Jacobi loops terminate on the
device.  NumPy arrays are staged through the GPU (parity mode); CUDA tensors / ``DeviceField`` are used in place.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .device import Stage, ptr, stream_ptr

_call = _lib.call


class StaticPDEExtrapolation:
    def __init__(self, dx, grid_size_r, grid_size_z, extrap_tol, extrap_band, periodic=False,
                 per_communicator_gen=None, per_communicator_eta=None):
        if not torch.cuda.is_available():
            raise _lib.AxbError("StaticPDEExtrapolation needs a CUDA device (no CPU fallback)")
        self.dx, self.grid_size_r, self.grid_size_z = dx, grid_size_r, grid_size_z
        self.extrap_tol, self.extrap_band = extrap_tol, extrap_band
        if periodic and (not per_communicator_gen) and (not per_communicator_eta):
            raise ValueError("Periodic commuicators cannot be NoneType for periodic BCs")
        self.periodic = periodic
        self.per_communicator_gen, self.per_communicator_eta = per_communicator_gen, per_communicator_eta
        self.eps = np.finfo(float).eps
        self.offset = np.sqrt(2) * dx
        self.r_start, self.r_end, self.z_start, self.z_end = 0, grid_size_r, 0, grid_size_z
        self.r_start_at_boundary = self.r_end_at_boundary = True
        self.z_start_at_boundary = self.z_end_at_boudary = True
        self.sweeps = (0, 0)
        self._work = None

    def extrapolate(self, eta, phi):
        """bounded_static_PDE_extrapolation.py:58-121"""
        if self.periodic:
            self.per_communicator_gen(phi)
            self.per_communicator_eta(eta)
        st = Stage()
        t_eta, t_phi = st.dev(eta, out=True), st.dev(phi)
        self._find_bounding_box(t_phi)
        box = (slice(self.r_start, self.r_end), slice(self.z_start, self.z_end))
        phi_b = (-t_phi[box]).contiguous()
        eta_b = t_eta[box].clone(memory_format=torch.contiguous_format)   # a copy even when the box spans whole rows
        n0, n1 = phi_b.shape
        ni = (n0 - 4, n1 - 4)
        nrp, nrn, nzp, nzn, den = (torch.empty(ni, dtype=torch.float64, device="cuda") for _ in range(5))
        zone = torch.empty(ni, dtype=torch.uint8, device="cuda")
        gn = torch.empty((n0, n1), dtype=torch.float64, device="cuda")
        s = stream_ptr()
        _call("axb_pde_extrap_setup", n0, n1, ptr(phi_b), ptr(eta_b), float(self.dx), float(self.offset),
              float(self.extrap_band), float(self.eps), ptr(nrp), ptr(nrn), ptr(nzp), ptr(nzn), ptr(den), ptr(zone),
              ptr(gn), s)
        nbytes = int(_call("axb_pde_extrap_workspace_bytes", n0, n1))
        if self._work is None or self._work.numel() < nbytes:
            self._work = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        k1, k2 = ctypes.c_int(0), ctypes.c_int(0)
        # H(psi) grad_phi . grad(grad_eta_n) = 0, then H(psi) (grad_phi . grad_eta - grad_eta_n) = 0   (:95-119)
        for soln, rhs, k in ((gn, None, k1), (eta_b, gn, k2)):
            _call("axb_pde_extrap_jacobi", n0, n1, ptr(soln), ptr(rhs) if rhs is not None else None, ptr(zone), ptr(den),
                  ptr(nrp), ptr(nrn), ptr(nzp), ptr(nzn), float(self.dx), float(self.extrap_tol), 0, ptr(self._work),
                  nbytes, ctypes.byref(k), s)
        self.sweeps = (k1.value, k2.value)
        self._restore_eta(t_eta, eta_b)
        st.finish()

    def _find_bounding_box(self, phi):
        """:126-153 -- the any() reductions run on the device, four integers come back"""
        inside = phi + self.extrap_band >= 0
        r_axis, z_axis = torch.any(inside, dim=1), torch.any(inside, dim=0)
        rr, zz = torch.nonzero(r_axis)[:, 0], torch.nonzero(z_axis)[:, 0]
        if rr.numel() == 0:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")      # what np.where(...)[0][[0, -1]] raises
        b = torch.stack([rr[0], rr[-1], zz[0], zz[-1]]).cpu().tolist()
        self.r_start, self.r_end, self.z_start, self.z_end = b[0], b[1] + 1, b[2], b[3] + 1
        if self.r_start >= 2:
            self.r_start -= 2
            self.r_start_at_boundary = False
        if self.r_end <= self.grid_size_r - 2:
            self.r_end += 2
            self.r_end_at_boundary = False
        if self.z_start >= 2:
            self.z_start -= 2
            self.z_start_at_boundary = False
        if self.z_end <= self.grid_size_z - 2:
            self.z_end += 2
            self.z_end_at_boudary = False
        if min(self.r_end - self.r_start, self.z_end - self.z_start) < 6:
            raise ValueError("Too few grid points. Using higher resolution!")

    def _restore_eta(self, eta, bounded_eta):
        """:220-235"""
        r0 = 0 if self.r_start_at_boundary else 2
        r1 = None if self.r_end_at_boundary else -2
        z0 = 0 if self.z_start_at_boundary else 2
        z1 = None if self.z_end_at_boudary else -2
        eta.zero_()
        eta[self.r_start + r0:self.r_end + (0 if r1 is None else r1),
            self.z_start + z0:self.z_end + (0 if z1 is None else z1)] = bounded_eta[r0:r1, z0:z1]
