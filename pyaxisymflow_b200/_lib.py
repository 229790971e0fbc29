"""ctypes binding of ``libaxisym_b200.so`` (the C ABI declared in ``include/axisym_b200.h``).

There is no CPU fallback: if the shared library is missing, or an entry point returns a
non-zero code, an exception is raised.  The library is built in-tree by
``make -C pyaxisymflow_b200/csrc`` (``__graft_entry__.build()`` does that).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_double, c_int, c_int16, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaxisym_b200.so")


class AxbError(RuntimeError):
    pass


class AxbGrid(Structure):
    _fields_ = [("nr", c_int32), ("nz", c_int32), ("ld", c_int64), ("dx", c_double),
                ("kz0", c_int32), ("nz_global", c_int32), ("ku0", c_int32), ("ku1", c_int32),
                ("ju0", c_int32), ("ju1", c_int32),
                ("batch", c_int32), ("scalar_stride", c_int32), ("batch_stride", c_int64)]


class AxbFdPlan(Structure):
    _fields_ = [("nr", c_int32), ("nz", c_int32),
                ("Lr", c_void_p), ("Rz", c_void_p), ("Rzb", c_void_p), ("Lrb", c_void_p),
                ("lam_r", c_void_p), ("lam_z", c_void_p),
                ("c0", c_double), ("c1", c_double), ("work", c_void_p),
                ("n_leaves", c_int32), ("n_folds", c_int32),
                ("leaf_n", c_int32 * 8), ("leaf_off", c_int32 * 8), ("fold_len", c_int32 * 8),
                ("leaf_fwd", c_void_p * 8), ("leaf_bwd", c_void_p * 8),
                ("r_tridiagonal", c_int32),
                ("r_sub", c_void_p), ("r_diag", c_void_p), ("r_sup", c_void_p), ("r_scale", c_void_p),
                ("z_fft", c_int32), ("z_tables", c_void_p), ("r_inv_pivots", c_void_p),
                ("r_row_coef", c_void_p), ("nz_spec", c_int32)]


_G = POINTER(AxbGrid)
_P = c_void_p   # device pointer
_D = c_double
_I = c_int
_S = c_void_p   # cudaStream_t

# name -> argtypes (every function returns int unless listed in _RESTYPE)
_SIGNATURES = {
    "axb_version": [],
    "axb_launch_count": [],
    "axb_set_stencil_path": [_I],
    "axb_set_solid_march": [_I],
    "axb_set_tridiag_sweep": [_I],
    "axb_debug_edge_blocks": [_G, _I, _I, _I, _I, _P, _I, _P],
    "axb_kill_boundary_vorticity_sine_z": [_G, _P, _P, _I, _S],
    "axb_kill_boundary_vorticity_sine_r": [_G, _P, _P, _I, _S],
    "axb_kill_boundary_vorticity_sine_r_parts": [_G, _P, _P, _I, _I, _S],
    "axb_periodic_ghost_comm": [_G, _P, _I, _D, _D, _S],
    "axb_velocity_from_psi": [_G, _P, _P, _P, _P, _D, _D, _P, _P, _S],
    "axb_brinkmann_penalize": [_G, _D, _D, _P, _D, _D, _P, _P, _P, _P, _P, _P, _S],
    "axb_vorticity_from_velocity": [_G, _P, _P, _P, _P, _P, _I, _S],
    "axb_penalise_update_vorticity": [_G, _P, _P, _P, _P, _P, _P, _D, _D, _P, _D, _D, _P, _P, _P, _S],
    "axb_advect_vorticity_eno3": [_G, _P, _P, _P, _P, _D, _P, _S],
    "axb_advect_refmap_eno3": [_G, _P, _P, _P, _P, _P, _P, _D, _P, _S],
    "axb_eno3_flux": [_G, _P, _P, _P, _P, _D, _I, _S],
    "axb_eno3_euler_step": [_G, _P, _P, _P, _P, _D, _I, _S],
    "axb_elementwise_sum": [_G, _P, _P, _P, _S],
    "axb_set_fixed_val": [_G, _P, _D, _S],
    "axb_diffusion_rk2_stage1": [_G, _P, _P, _P, _D, _D, _P, _S],
    "axb_diffusion_rk2_stage2": [_G, _P, _P, _P, _P, _D, _D, _P, _S],
    "axb_diffusion_rk2_fused": [_G, _P, _P, _P, _P, _D, _D, _P, _S],
    "axb_smooth_heaviside": [_G, _P, _P, _D, _S],
    "axb_smooth_heaviside_sphere": [_G, _P, _P, _P, _P, _D, _D, _D, _D, _S],
    "axb_vortex_stretching": [_G, _P, _P, _P, _D, _S],
    "axb_velocity_from_phi": [_G, _P, _P, _P, _S],
    "axb_baroclinic_vorticity_update": [_G, _P, _P, _P, _P, _P, _P, _P, _P, _P, _D, _D, _I, _S],
    "axb_reinit_workspace_bytes": [_I, _I],
    "axb_reinit_distance": [_G, _P, _D, _I, _P, _P, c_int64, POINTER(c_int), _S],
    "axb_reduce_max_abs_sum": [_G, _P, _P, _P, _S],
    "axb_reduce_max": [_G, _P, _P, _S],
    "axb_reduce_weighted_sum": [_G, _P, _P, _P, _D, _P, _S],
    "axb_fill_scalars": [_P, _I, _D, _S],
    "axb_rigid_flow_scalars": [_I, _P, _D, _D, _D, _D, _D, _S],
    "axb_axpy": [_G, _P, _P, _D, _P, _S],
    "axb_pin_level_set": [_G, _P, _P, _P, _P, _D, _D, _D, _D, _S],
    "axb_smooth_heaviside_mask": [_G, _P, _P, _P, _D, _D, _I, _S],
    "axb_add_bubble_flow": [_G, _P, _P, _P, _P, _P, _D, _D, _D, _D, _D, _S],
    "axb_add_bubble_flow_dev": [_G, _P, _P, _P, _P, _P, _D, _D, _D, _D, _P, _P, _S],
    "axb_bubble_flow_geometry": [_G, _P, _P, _P, _P, _D, _D, _S],
    "axb_add_bubble_flow_geom": [_G, _P, _P, _P, c_int64, _P, _P, _D, _D, _D, _D, _P, _P, _S],
    "axb_soft_sphere_scalars": [_I, _P, _D, _D, _D, _D, _D, _D, _D, _D, _D, _D, _S],
    "axb_particle_scalars_batched": [_I, _I, _I, _P, _P, _I, _D, _D, _D, _D, _D, _D, _D, _S],
    "axb_diffusion_rk2_stage1_dev": [_G, _P, _P, _P, _P, _P, _S],
    "axb_diffusion_rk2_stage2_dev": [_G, _P, _P, _P, _P, _P, _P, _S],
    "axb_smooth_heaviside_sphere_dev": [_G, _P, _P, _P, _P, _P, _D, _D, _D, _S],
    "axb_particle_scalars": [_I, _P, _P, _I, _D, _D, _D, _D, _D, _D, _D, _D, _D, _D, _S],
    "axb_cycle_average3": [_G, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _S],
    "axb_solid_sigma": [_G, _P, _P, _P, _D, _P, _P, _P, _P, _P, _P, _P, _S],
    "axb_solid_tau": [_G, _P, _P, _P, _P, _P, _P, _S],
    "axb_solid_vorticity_update": [_G, _P, _P, _P, _D, _P, _S],
    "axb_solid_stress_vorticity_update": [_G, _P, _P, _P, _P, _P, _D, _D, _P, _I, _S],
    "axb_ls_workspace_bytes": [_I, _I],
    "axb_ls_extrapolate_order1": [_I, _I, _P, _P, _P, _P, _P, _P, _P, c_int64, _I, POINTER(c_int), _S],
    "axb_ls_extrapolate_order2": [_I, _I, _P, _P, _P, _P, _P, _P, _P, c_int64, _I, POINTER(c_int), _S],
    "axb_ls_extrapolate_eta": [_G, _P, _P, _P, _P, _D, _P, _P, _P, c_int64, _I, POINTER(c_int), _S],
    "axb_ls_extrapolate_eta_device": [_G, _P, _P, _P, _P, _P, _P, _D, _P, _P, _P, c_int64, _I, _P, _S],
    "axb_ls_extrapolate_eta_device_parts": [_G, _P, _P, _P, _P, _P, _P, _D, _P, _P, _P, c_int64, _I, _P, _I, _S],
    "axb_p2m_mp4_2d": [_I, _I, _P, _P, _P, _P, _D, _D, _I, _S],
    "axb_rfft_rows": [_I, _I, _P, c_int64, _P, c_int64, _I, _P, _D, _S],
    "axb_irfft_rows": [_I, _I, _P, c_int64, _P, c_int64, _P, _D, _S],
    "axb_rfft_supported": [_I],
    "axb_pde_extrap_workspace_bytes": [_I, _I],
    "axb_pde_extrap_setup": [_I, _I, _P, _P, _D, _D, _D, _D, _P, _P, _P, _P, _P, _P, _P, _S],
    "axb_pde_extrap_jacobi": [_I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _D, _D, _I, _P, c_int64, POINTER(c_int), _S],
    "axb_m2p_2d": [_I, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _D, _D, _I, _S],
    "axb_p2m_2d": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _D, _D, _I, _S],
    "axb_m2p_1d_mp4": [_I, _P, _I, _P, _P, _D, _S],
    "axb_p2m_1d_mp4": [_I, _I, _P, _P, _P, _D, _S],
    "axb_wrap_particles_2d": [_I, _I, _P, _P, _D, _D, _D, _D, _S],
    "axb_advect_vorticity_particles": [_G, _P, _P, _P, _P, _P, _P, _D, _P, _I, _S],
    "axb_advect_vorticity_particles_flagged": [_G, _P, _P, _P, _P, _P, _P, _D, _P, _P, _S],
    "axb_set_p2m_atomic": [_I],
    "axb_fd_solve": [POINTER(AxbFdPlan), _P, c_int64, _P, c_int64, _S],
    "axb_dgemm": [_I, _I, _I, _P, c_int64, _P, c_int64, _P, c_int64, _P, _P, _D, _D, _S],
    "axb_dgemm_set_path": [_I],
    "axb_fd_fold": [_I, _I, _P, c_int64, _I, _S],
    "axb_fd_fold2": [_I, _I, _P, c_int64, _P, c_int64, _I, _S],
    "axb_dct2_rows": [_I, _I, _P, c_int64, _P, c_int64, _P, _D, _D, _S],
    "axb_dct3_rows": [_I, _I, _P, c_int64, _P, c_int64, _P, _S],
    "axb_tridiag_factor_columns": [_I, _I, _P, _P, _P, _P, _P, _D, _D, _P, _P, _S],
    "axb_tridiag_solve_factored": [_I, _I, _P, c_int64, _P, _P, _S],
    "axb_tridiag_partition_correct": [_I, _I, _P, c_int64, _P, _P, _P, _P, _P, _I, _S],
    "axb_tridiag_partition_correct_banded": [_I, _I, _P, c_int64, _P, _P, _P, _P, _P, _I, _P, _P, _S],
    "axb_tridiag_solve_columns": [_I, _I, _P, c_int64, _P, _P, _P, _P, _P, _D, _D, _P, _S],
    "axb_halo_pack": [_G, _P, _P, _P, _I, _S],
    "axb_halo_unpack": [_G, _P, _P, _P, _I, _D, _S],
    "axb_slab_to_blocks": [_I, _I, c_int64, _I, _P, _P, _S],
    "axb_blocks_to_rows": [_I, _I, _I, _P, _P, c_int64, _S],
    "axb_halo_put": [_G, _P, _P, _P, _I, _D, _S],
    "axb_row_halo_put": [_I, _P, _P, _P, c_int64, _I, _I, _I, _I, _S],
    "axb_row_halo_get": [_I, _P, _P, _P, c_int64, _I, _I, _I, _I, _S],
    "axb_peer_sync": [_I, _I, _P, _P, _S],
    "axb_peer_allreduce_max": [_I, _I, _P, _P, _P, _S],
    "axb_row_halo_exchange": [_I, _P, _P, _P, c_int64, _I, _I, _I, _I, _P, _P, _P, _P, _S],
    "axb_peer_block_put": [_I, _I, _P, c_int64, c_int64, _P, c_int64, c_int64, _I, _I, _S],
    "axb_rows_to_blocks": [_I, _I, _I, _P, c_int64, _P, _S],
    "axb_blocks_to_slab": [_I, _I, c_int64, _I, _P, _P, _S],
}
_RESTYPE = {"axb_launch_count": c_int64, "axb_ls_workspace_bytes": c_int64, "axb_reinit_workspace_bytes": c_int64,
            "axb_pde_extrap_workspace_bytes": c_int64}
_NO_CHECK = {"axb_version", "axb_launch_count", "axb_ls_workspace_bytes", "axb_reinit_workspace_bytes",
             "axb_pde_extrap_workspace_bytes", "axb_rfft_supported"}

_ERR = {-1: "AXB_EINVAL (null pointer / bad shape)", -2: "AXB_EALIGN (misaligned pointer)",
        -3: "AXB_ENOSUP (unsupported configuration)", -4: "AXB_EWORK (workspace too small)"}

_lib = None


def exported_names():
    """Entry points include/axisym_b200.h declares (used by the CPU-side symbol test)."""
    return sorted(_SIGNATURES)


def load():
    """dlopen the library and attach prototypes; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AxbError(
                f"{LIB_PATH} is missing: build it with `make -C pyaxisymflow_b200/csrc` "
                "(there is deliberately no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, c_int)
        _lib = lib
    return _lib


def call(name, *args):
    """Call an entry point and turn a non-zero return code into an exception."""
    rc = getattr(load(), name)(*args)
    if name in _NO_CHECK:
        return rc
    if rc != 0:
        what = _ERR.get(rc, f"cudaError_t {rc}" if rc > 0 else f"error {rc}")
        raise AxbError(f"{name} failed: {what}")
    return 0


def launch_count():
    return int(load().axb_launch_count())
