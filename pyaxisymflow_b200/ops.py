"""Drop-in kernel calls: same names, argument order and in-place semantics as the reference's
``pyaxisymflow.kernels`` / ``elasto_kernels`` / ``pyst_kernels`` / ``core`` callables, executed
by the sm_100a kernels of ``libaxisym_b200`` through ctypes.  The mirrored module tree
(``pyaxisymflow_b200.kernels.*`` ...) re-exports these functions under the reference's paths.

Arrays may be ``DeviceField`` / CUDA tensors (zero copy) or NumPy arrays (parity mode: staged
to the GPU and back).  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .device import DeviceField, Stage, coord_1d, make_grid, ptr, stream_ptr, unwrap

_call = _lib.call


def _is_field(x):
    return isinstance(x, (np.ndarray, torch.Tensor, DeviceField)) and getattr(unwrap(x), "ndim", 0) == 2


def _scalar(x):
    if isinstance(x, DeviceField):
        x = x.t
    if isinstance(x, torch.Tensor):
        return float(x)
    return float(x)


# --------------------------------------------------------------------------------------
# kernels/periodic_boundary_ghost_comm.py
# --------------------------------------------------------------------------------------
class _GhostComm:
    """Callable returned by gen_periodic_boundary_ghost_comm(_eta); carries its parameters so
    that the ``*_periodic`` kernels can be handed the communicator like in the reference."""

    def __init__(self, ghost_size, z_max=0.0, two_g_dx=0.0):
        self.ghost_size, self.z_max, self.two_g_dx = ghost_size, float(z_max), float(two_g_dx)

    def __call__(self, field):
        st = Stage()
        (f,), _, ld = st.fields(outs=[field])
        g = make_grid(f.shape[0], f.shape[1], ld, 1.0)
        _call("axb_periodic_ghost_comm", ctypes.byref(g), ptr(f), self.ghost_size, self.z_max, self.two_g_dx,
              stream_ptr())
        st.finish()


def gen_periodic_boundary_ghost_comm(ghost_size):
    """kernels/periodic_boundary_ghost_comm.py:4-15"""
    assert ghost_size > 0 and isinstance(ghost_size, int), "invalid ghost size"
    return _GhostComm(ghost_size)


def gen_periodic_boundary_ghost_comm_eta(ghost_size, Z_max, dx):
    """kernels/periodic_boundary_ghost_comm.py:18-33"""
    assert ghost_size > 0 and isinstance(ghost_size, int), "invalid ghost size"
    return _GhostComm(ghost_size, Z_max, 2 * ghost_size * dx)


# --------------------------------------------------------------------------------------
# kernels/kill_boundary_vorticity_sine.py
# --------------------------------------------------------------------------------------
def kill_boundary_vorticity_sine_z(vorticity, Z, width, dx):
    """kernels/kill_boundary_vorticity_sine.py:4-14"""
    st = Stage()
    (w,), _, ld = st.fields(outs=[vorticity])
    g = make_grid(w.shape[0], w.shape[1], ld, dx)
    z1 = coord_1d(st, Z, 1, w.shape[1])
    _call("axb_kill_boundary_vorticity_sine_z", ctypes.byref(g), ptr(w), ptr(z1), int(width), stream_ptr())
    st.finish()


def kill_boundary_vorticity_sine_r(vorticity, R, width, dx):
    """kernels/kill_boundary_vorticity_sine.py:17-27"""
    st = Stage()
    (w,), _, ld = st.fields(outs=[vorticity])
    g = make_grid(w.shape[0], w.shape[1], ld, dx)
    r1 = coord_1d(st, R, 0, w.shape[0])
    _call("axb_kill_boundary_vorticity_sine_r", ctypes.byref(g), ptr(w), ptr(r1), int(width), stream_ptr())
    st.finish()


# --------------------------------------------------------------------------------------
# kernels/compute_velocity_from_psi.py
# --------------------------------------------------------------------------------------
def compute_velocity_from_psi_unb(u_z, u_r, psi, R, dx):
    """kernels/compute_velocity_from_psi.py:4-17"""
    st = Stage()
    (uz, ur), (p,), ld = st.fields(outs=[u_z, u_r], ins=[psi])
    g = make_grid(p.shape[0], p.shape[1], ld, dx)
    r1 = coord_1d(st, R, 0, p.shape[0])
    _call("axb_velocity_from_psi", ctypes.byref(g), ptr(uz), ptr(ur), ptr(p), ptr(r1), 0.0, 0.0, None, None,
          stream_ptr())
    st.finish()


def compute_velocity_from_psi_periodic(u_z, u_r, psi, R, dx, per_communicator):
    """kernels/compute_velocity_from_psi.py:20-34"""
    per_communicator(psi)
    compute_velocity_from_psi_unb(u_z, u_r, psi, R, dx)


# --------------------------------------------------------------------------------------
# kernels/brinkmann_penalize.py
# --------------------------------------------------------------------------------------
def brinkmann_penalize(lam, dt, char_func, U_z, U_r, grid_u_z, grid_u_r, penalized_u_z, penalized_u_r):
    """kernels/brinkmann_penalize.py:4-16 (U_z / U_r: scalars or fields)"""
    st = Stage()
    fz, fr = _is_field(U_z), _is_field(U_r)
    ins = [char_func, grid_u_z, grid_u_r] + ([U_z] if fz else []) + ([U_r] if fr else [])
    (pz, pr), tin, ld = st.fields(outs=[penalized_u_z, penalized_u_r], ins=ins)
    chi, gz, gr = tin[:3]
    rest = tin[3:]
    Uzf = rest.pop(0) if fz else None
    Urf = rest.pop(0) if fr else None
    g = make_grid(chi.shape[0], chi.shape[1], ld, 1.0)
    _call("axb_brinkmann_penalize", ctypes.byref(g), float(lam), float(dt), ptr(chi),
          0.0 if fz else _scalar(U_z), 0.0 if fr else _scalar(U_r), ptr(Uzf), ptr(Urf), ptr(gz), ptr(gr),
          ptr(pz), ptr(pr), stream_ptr())
    st.finish()


# --------------------------------------------------------------------------------------
# kernels/compute_vorticity_from_velocity.py
# --------------------------------------------------------------------------------------
def compute_vorticity_from_velocity_unb(vort, u_z, u_r, dx):
    """kernels/compute_vorticity_from_velocity.py:4-13"""
    st = Stage()
    (v,), (uz, ur), ld = st.fields(outs=[vort], ins=[u_z, u_r])
    g = make_grid(v.shape[0], v.shape[1], ld, dx)
    _call("axb_vorticity_from_velocity", ctypes.byref(g), ptr(v), ptr(uz), ptr(ur), None, None, 0, stream_ptr())
    st.finish()


def compute_vorticity_from_velocity_periodic(vort, u_z, u_r, dx, per_communicator):
    """kernels/compute_vorticity_from_velocity.py:16-28"""
    per_communicator(u_r)
    per_communicator(u_z)
    compute_vorticity_from_velocity_unb(vort, u_z, u_r, dx)


def penalise_and_update_vorticity(u_z, u_r, vorticity, u_z_upen, u_r_upen, char_func, lam, dt, U_z, U_r, R, dx,
                                  want_sum=False):
    """Fused G-PEN pass (flow_past_sphere.py:155-175): u = pen(u_upen), vorticity += curl(u - u_upen).
    Returns sum(R * chi * (u_z - U_z)) when ``want_sum`` (the drag / force numerator)."""
    st = Stage()
    (uz, ur, w), (zu, ru, chi), ld = st.fields(outs=[u_z, u_r, vorticity], ins=[u_z_upen, u_r_upen, char_func])
    g = make_grid(w.shape[0], w.shape[1], ld, dx)
    r1 = coord_1d(st, R, 0, w.shape[0])
    acc = torch.zeros(1, dtype=torch.float64, device="cuda") if want_sum else None
    _call("axb_penalise_update_vorticity", ctypes.byref(g), ptr(uz), ptr(ur), ptr(w), ptr(zu), ptr(ru), ptr(chi),
          float(lam), float(dt), None, _scalar(U_z), _scalar(U_r), None, ptr(r1), ptr(acc), stream_ptr())
    st.finish()
    return float(acc) if want_sum else None


# --------------------------------------------------------------------------------------
# kernels/diffusion_RK2.py
# --------------------------------------------------------------------------------------
def diffusion_RK2_unb(vorticity, temp_vorticity, R, nu, dt, dx, _per=None):
    """kernels/diffusion_RK2.py:4-45"""
    if _per is not None:
        _per(vorticity)
    st = Stage()
    (w, tmp), _, ld = st.fields(outs=[vorticity, temp_vorticity])
    g = make_grid(w.shape[0], w.shape[1], ld, dx)
    r1 = coord_1d(st, R, 0, w.shape[0])
    _call("axb_diffusion_rk2_stage1", ctypes.byref(g), ptr(tmp), ptr(w), ptr(r1), float(nu), float(dt), None,
          stream_ptr())
    if _per is not None:
        _per(DeviceField(tmp))
    _call("axb_diffusion_rk2_stage2", ctypes.byref(g), ptr(w), ptr(w), ptr(tmp), ptr(r1), float(nu), float(dt),
          None, stream_ptr())
    st.finish()


def diffusion_RK2_periodic(vorticity, temp_vorticity, R, nu, dt, dx, per_communicator):
    """kernels/diffusion_RK2.py:48-89"""
    diffusion_RK2_unb(vorticity, temp_vorticity, R, nu, dt, dx, _per=per_communicator)


# --------------------------------------------------------------------------------------
# kernels/smooth_Heaviside.py, vortex_stretching.py, compute_forces.py, force_projection.py
# --------------------------------------------------------------------------------------
def smooth_Heaviside(H, phi, blend_w):
    """kernels/smooth_Heaviside.py:5-14"""
    st = Stage()
    (h,), (p,), ld = st.fields(outs=[H], ins=[phi])
    g = make_grid(h.shape[0], h.shape[1], ld, 1.0)
    _call("axb_smooth_heaviside", ctypes.byref(g), ptr(h), ptr(p), float(blend_w), stream_ptr())
    st.finish()


def smooth_Heaviside_sphere(H, Z, R, z_cm, r_cm, radius, blend_w, phi_out=None):
    """phi = radius - sqrt((Z-z_cm)^2 + (R-r_cm)^2) built in-kernel, then smooth_Heaviside
    (flow_past_sphere.py:80-82): 8 B/pt instead of 24."""
    st = Stage()
    outs = [H] + ([phi_out] if phi_out is not None else [])
    touts, _, ld = st.fields(outs=outs)
    h = touts[0]
    po = touts[1] if phi_out is not None else None
    g = make_grid(h.shape[0], h.shape[1], ld, 1.0)
    z1, r1 = coord_1d(st, Z, 1, h.shape[1]), coord_1d(st, R, 0, h.shape[0])
    _call("axb_smooth_heaviside_sphere", ctypes.byref(g), ptr(h), ptr(po), ptr(z1), ptr(r1), float(z_cm),
          float(r_cm), float(radius), float(blend_w), stream_ptr())
    st.finish()


def vortex_stretching(vorticity, u_r, R, dt):
    """kernels/vortex_stretching.py:4-11"""
    st = Stage()
    (w,), (ur,), ld = st.fields(outs=[vorticity], ins=[u_r])
    g = make_grid(w.shape[0], w.shape[1], ld, 1.0)
    r1 = coord_1d(st, R, 0, w.shape[0])
    _call("axb_vortex_stretching", ctypes.byref(g), ptr(w), ptr(ur), ptr(r1), float(dt), stream_ptr())
    st.finish()


# --------------------------------------------------------------------------------------
# kernels/compute_velocity_from_phi.py, kernels/update_baroclinic_vorticity.py  
# --------------------------------------------------------------------------------------
def compute_velocity_from_phi_unb(u_z, u_r, phi, dx):
    """kernels/compute_velocity_from_phi.py:4-17"""
    st = Stage()
    (uz, ur), (p,), ld = st.fields(outs=[u_z, u_r], ins=[phi])
    g = make_grid(p.shape[0], p.shape[1], ld, dx)
    _call("axb_velocity_from_phi", ctypes.byref(g), ptr(uz), ptr(ur), ptr(p), stream_ptr())
    st.finish()


def _baroclinic(mode, vorticity, u_z, u_r, old_u_z, old_u_r, density, penal_z, penal_r, R, nu, dt, dx):
    st = Stage()
    ins = [u_z, u_r, old_u_z, old_u_r, density] + ([penal_z, penal_r] if mode >= 1 else [])
    (w,), tin, ld = st.fields(outs=[vorticity], ins=ins)
    uz, ur, oz, orr, rho = tin[:5]
    pz, pr = (tin[5], tin[6]) if mode >= 1 else (None, None)
    g = make_grid(w.shape[0], w.shape[1], ld, dx)
    r1 = coord_1d(st, R, 0, w.shape[0]) if mode == 2 else None
    _call("axb_baroclinic_vorticity_update", ctypes.byref(g), ptr(w), ptr(uz), ptr(ur), ptr(oz), ptr(orr), ptr(rho),
          ptr(pz), ptr(pr), ptr(r1), float(nu), float(dt), mode, stream_ptr())
    st.finish()


def update_baroclinic_vorticity(vorticity, u_z, u_r, old_u_z, old_u_r, density, dt, dx):
    """kernels/update_baroclinic_vorticity.py:4-35"""
    _baroclinic(0, vorticity, u_z, u_r, old_u_z, old_u_r, density, None, None, None, 0.0, dt, dx)


def update_baroclinic_vorticity_penal(vorticity, u_z, u_r, old_u_z, old_u_r, density, penal_term_z, penal_term_r,
                                      dt, dx):
    """kernels/update_baroclinic_vorticity.py:38-67"""
    _baroclinic(1, vorticity, u_z, u_r, old_u_z, old_u_r, density, penal_term_z, penal_term_r, None, 0.0, dt, dx)


def update_baroclinic_vorticity_diff_penal(vorticity, u_z, u_r, old_u_z, old_u_r, density, penal_term_z,
                                           penal_term_r, R, nu, dt, dx):
    """kernels/update_baroclinic_vorticity.py:70-127"""
    _baroclinic(2, vorticity, u_z, u_r, old_u_z, old_u_r, density, penal_term_z, penal_term_r, R, nu, dt, dx)


def _reduce(name, fields, R=None, off=0.0, init=0.0):
    st = Stage()
    _, tins, ld = st.fields(ins=fields)
    g = make_grid(tins[0].shape[0], tins[0].shape[1], ld, 1.0)
    out = torch.full((1,), init, dtype=torch.float64, device="cuda")
    if name == "axb_reduce_weighted_sum":
        r1 = coord_1d(st, R, 0, tins[0].shape[0])
        _call(name, ctypes.byref(g), ptr(r1), ptr(tins[0]), ptr(tins[1]), float(off), ptr(out), stream_ptr())
    elif name == "axb_reduce_max":
        _call(name, ctypes.byref(g), ptr(tins[0]), ptr(out), stream_ptr())
    else:
        _call(name, ctypes.byref(g), ptr(tins[0]), ptr(tins[1]) if len(tins) > 1 else None, ptr(out), stream_ptr())
    return float(out)


def max_abs_sum(a, b=None):
    """np.amax(np.fabs(a) + np.fabs(b))  -- the CFL reduction of flow_past_sphere.py:152"""
    return _reduce("axb_reduce_max_abs_sum", [a] + ([b] if b is not None else []))


def field_max(a):
    """np.amax(a)  -- flow_past_sphere.py:191"""
    return _reduce("axb_reduce_max", [a], init=-np.inf)


def weighted_sum(R, c, a, off=0.0):
    """np.sum(R * c * (a - off))"""
    return _reduce("axb_reduce_weighted_sum", [c, a], R=R, off=off)


def compute_force_on_body(R, part_char_func, rho_f, brink_lam, u_z, U_z_cm_part, part_vol, dt, diff):
    """kernels/compute_forces.py:4-17"""
    F_pen = rho_f * brink_lam * weighted_sum(R, part_char_func, u_z, U_z_cm_part)
    F_un = (diff * part_vol) / dt
    return F_pen, F_un


def force_projection(rho_s, char_func, u_z, u_r, R):
    """kernels/force_projection.py:5-15 (rho_s scalar)"""
    ones = torch.ones_like(Stage().dev(char_func))
    m = rho_s * weighted_sum(R, char_func, ones)
    return rho_s * weighted_sum(R, char_func, u_z) / m, rho_s * weighted_sum(R, char_func, u_r) / m


# --------------------------------------------------------------------------------------
# pyst_kernels/*  (generator closures)
# --------------------------------------------------------------------------------------
def _check_fixed(fixed_grid_size, shape):
    if fixed_grid_size and tuple(fixed_grid_size) != tuple(shape):
        raise ValueError(f"Wrong shape for fixed-size kernel: expected {tuple(fixed_grid_size)}, got {tuple(shape)}")


def gen_elementwise_sum_pyst_kernel(real_t=np.float64, num_threads=False, fixed_grid_size=False, field_type="scalar"):
    """pyst_kernels/elementwise_ops.py:9-52"""
    assert field_type == "scalar" or field_type == "vector", "Invalid field type"

    def elementwise_sum_pyst_kernel(sum_field, field_1, field_2):
        if field_type == "vector":
            for c in range(unwrap(sum_field).shape[0]):
                elementwise_sum_2d(sum_field[c], field_1[c], field_2[c], fixed_grid_size)
        else:
            elementwise_sum_2d(sum_field, field_1, field_2, fixed_grid_size)

    return elementwise_sum_pyst_kernel


def elementwise_sum_2d(sum_field, field_1, field_2, fixed_grid_size=False):
    st = Stage()
    (s,), (a, b), ld = st.fields(outs=[sum_field], ins=[field_1, field_2])
    _check_fixed(fixed_grid_size, s.shape)
    g = make_grid(s.shape[0], s.shape[1], ld, 1.0)
    _call("axb_elementwise_sum", ctypes.byref(g), ptr(s), ptr(a), ptr(b), stream_ptr())
    st.finish()


def gen_set_fixed_val_pyst_kernel(real_t=np.float64, num_threads=False, fixed_grid_size=False, field_type="scalar"):
    """pyst_kernels/elementwise_ops.py:55-104"""
    assert field_type == "scalar" or field_type == "vector", "Invalid field type"

    def set_fixed_val_pyst_kernel(field, fixed_val):
        st = Stage()
        (f,), _, ld = st.fields(outs=[field])
        _check_fixed(fixed_grid_size, f.shape)
        g = make_grid(f.shape[0], f.shape[1], ld, 1.0)
        _call("axb_set_fixed_val", ctypes.byref(g), ptr(f), float(fixed_val), stream_ptr())
        st.finish()

    if field_type == "scalar":
        return set_fixed_val_pyst_kernel

    def vector_field_set_fixed_val_pyst_kernel(vector_field, fixed_vals):
        set_fixed_val_pyst_kernel(field=vector_field[0], fixed_val=fixed_vals[0])
        set_fixed_val_pyst_kernel(field=vector_field[1], fixed_val=fixed_vals[1])

    return vector_field_set_fixed_val_pyst_kernel


def _gen_flux(conservative, fixed_grid_size):
    def advection_flux_eno3_pyst_kernel(advection_flux, field, velocity, inv_dx):
        st = Stage()
        (fl,), (f, v0, v1), ld = st.fields(outs=[advection_flux], ins=[field, velocity[0], velocity[1]])
        _check_fixed(fixed_grid_size, f.shape)
        g = make_grid(f.shape[0], f.shape[1], ld, 1.0)
        _call("axb_eno3_flux", ctypes.byref(g), ptr(fl), ptr(f), ptr(v0), ptr(v1), float(inv_dx),
              int(conservative), stream_ptr())
        st.finish()

    return advection_flux_eno3_pyst_kernel


def gen_advection_flux_conservative_eno3_pyst_kernel(real_t=np.float64, num_threads=False, fixed_grid_size=False):
    """pyst_kernels/advection_flux.py:9-163"""
    return _gen_flux(True, fixed_grid_size)


def gen_advection_flux_non_conservative_eno3_pyst_kernel(real_t=np.float64, num_threads=False,
                                                         fixed_grid_size=False):
    """pyst_kernels/advection_flux.py:166-332"""
    return _gen_flux(False, fixed_grid_size)


def _gen_timestep(conservative, fixed_grid_size):
    def advection_timestep_euler_forward_eno3_pyst_kernel(field, advection_flux, velocity, dt_by_dx):
        # one fused launch: new = field + flux(field); the flux array receives the step's
        # flux like in the reference (advection_timestep.py:45-54) via a second tiny pass
        st = Stage()
        (f, fl), (v0, v1), ld = st.fields(outs=[field, advection_flux], ins=[velocity[0], velocity[1]])
        _check_fixed(fixed_grid_size, f.shape)
        g = make_grid(f.shape[0], f.shape[1], ld, 1.0)
        _call("axb_set_fixed_val", ctypes.byref(g), ptr(fl), 0.0, stream_ptr())
        _call("axb_eno3_flux", ctypes.byref(g), ptr(fl), ptr(f), ptr(v0), ptr(v1), -float(dt_by_dx),
              int(conservative), stream_ptr())
        _call("axb_elementwise_sum", ctypes.byref(g), ptr(f), ptr(f), ptr(fl), stream_ptr())
        st.finish()

    return advection_timestep_euler_forward_eno3_pyst_kernel


def gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel(real_t=np.float64, num_threads=False,
                                                                       fixed_grid_size=False):
    """pyst_kernels/advection_timestep.py:14-56"""
    return _gen_timestep(True, fixed_grid_size)


def gen_advection_timestep_euler_forward_non_conservative_eno3_pyst_kernel(real_t=np.float64, num_threads=False,
                                                                           fixed_grid_size=False):
    """pyst_kernels/advection_timestep.py:62-104"""
    return _gen_timestep(False, fixed_grid_size)


def eno3_euler_step(field_out, field_in, vel0, vel1, dt_by_dx, conservative=True):
    """Fused out-of-place Euler step on a plain 2-D array (the single-launch form of a5/a6)."""
    st = Stage()
    (fo,), (fi, v0, v1), ld = st.fields(outs=[field_out], ins=[field_in, vel0, vel1])
    g = make_grid(fo.shape[0], fo.shape[1], ld, 1.0)
    _call("axb_eno3_euler_step", ctypes.byref(g), ptr(fo), ptr(fi), ptr(v0), ptr(v1), float(dt_by_dx),
          int(conservative), stream_ptr())
    st.finish()


# --------------------------------------------------------------------------------------
# kernels/advect_vorticity_via_eno3.py, elasto_kernels/advect_refmap_via_eno3.py
# --------------------------------------------------------------------------------------
def gen_advect_vorticity_via_eno3(dx, grid_size_r, grid_size_z, real_t=np.float64, num_threads=False,
                                  _per=None):
    """kernels/advect_vorticity_via_eno3.py:8-44.  The closure owns one scratch field (the
    reference owns four doubled ones) because the fused kernel is out of place."""
    scratch = {}

    def advect_vorticity_via_eno3(vorticity, u_z, u_r, dt):
        if _per is not None:
            _per(u_z)
            _per(u_r)
            _per(vorticity)
        st = Stage()
        (w,), (uz, ur), ld = st.fields(outs=[vorticity], ins=[u_z, u_r])
        if tuple(w.shape) != (grid_size_r, grid_size_z):
            raise ValueError(f"Wrong shape: kernel was generated for {(grid_size_r, grid_size_z)}, got {tuple(w.shape)}")
        if "buf" not in scratch or scratch["buf"].shape != (grid_size_r, ld):
            scratch["buf"] = torch.empty((grid_size_r, ld), dtype=torch.float64, device="cuda")
        buf = scratch["buf"]
        g = make_grid(grid_size_r, grid_size_z, ld, dx)
        _call("axb_advect_vorticity_eno3", ctypes.byref(g), ptr(buf), ptr(w), ptr(uz), ptr(ur), float(dt), None,
              stream_ptr())
        w.copy_(buf[:, :grid_size_z])
        st.finish()

    return advect_vorticity_via_eno3


def gen_advect_vorticity_via_eno3_periodic(dx, grid_size_r, grid_size_z, per_communicator, real_t=np.float64,
                                           num_threads=False):
    """kernels/advect_vorticity_via_eno3.py:47-91"""
    return gen_advect_vorticity_via_eno3(dx, grid_size_r, grid_size_z, real_t, num_threads, _per=per_communicator)


def gen_advect_refmap_via_eno3(dx, grid_size_r, grid_size_z, real_t=np.float64, num_threads=False, _pers=None):
    """elasto_kernels/advect_refmap_via_eno3.py:8-54"""
    scratch = {}

    def advect_refmap_via_eno3(eta1, eta2, u_z, u_r, dt):
        if _pers is not None:
            p1, p2 = _pers
            p1(u_z)
            p1(u_r)
            p2(eta1)
            p1(eta2)
        st = Stage()
        (e1, e2), (uz, ur), ld = st.fields(outs=[eta1, eta2], ins=[u_z, u_r])
        if tuple(e1.shape) != (grid_size_r, grid_size_z):
            raise ValueError(f"Wrong shape: kernel was generated for {(grid_size_r, grid_size_z)}, got {tuple(e1.shape)}")
        if "b" not in scratch or scratch["b"].shape != (2, grid_size_r, ld):
            scratch["b"] = torch.empty((2, grid_size_r, ld), dtype=torch.float64, device="cuda")
        b = scratch["b"]
        g = make_grid(grid_size_r, grid_size_z, ld, dx)
        _call("axb_advect_refmap_eno3", ctypes.byref(g), ptr(b[0]), ptr(b[1]), ptr(e1), ptr(e2), ptr(uz), ptr(ur),
              float(dt), None, stream_ptr())
        e1.copy_(b[0][:, :grid_size_z])
        e2.copy_(b[1][:, :grid_size_z])
        st.finish()

    return advect_refmap_via_eno3


def gen_advect_refmap_via_eno3_periodic(dx, grid_size_r, grid_size_z, per_communicator1, per_communicator2,
                                        real_t=np.float64, num_threads=False):
    """elasto_kernels/advect_refmap_via_eno3.py:57-115"""
    return gen_advect_refmap_via_eno3(dx, grid_size_r, grid_size_z, real_t, num_threads,
                                      _pers=(per_communicator1, per_communicator2))


# --------------------------------------------------------------------------------------
# elasto_kernels/solid_sigma.py, div_tau.py
# --------------------------------------------------------------------------------------
def solid_sigma(sigma_s_11, sigma_s_12, sigma_s_22, G, dx, eta_1, eta_2, eta_1z, eta_1r, eta_2z, eta_2r,
                _chi=None):
    """elasto_kernels/solid_sigma.py:4-29"""
    st = Stage()
    outs, ins, ld = st.fields(outs=[sigma_s_11, sigma_s_12, sigma_s_22, eta_1z, eta_1r, eta_2z, eta_2r],
                              ins=[eta_1, eta_2] + ([_chi] if _chi is not None else []))
    s11, s12, s22, e1z, e1r, e2z, e2r = outs
    g = make_grid(s11.shape[0], s11.shape[1], ld, dx)
    _call("axb_solid_sigma", ctypes.byref(g), ptr(s11), ptr(s12), ptr(s22), float(G), ptr(ins[0]), ptr(ins[1]),
          ptr(e1z), ptr(e1r), ptr(e2z), ptr(e2r), ptr(ins[2]) if _chi is not None else None, stream_ptr())
    st.finish()


def solid_sigma_periodic(sigma_s_11, sigma_s_12, sigma_s_22, G, dx, eta_1, eta_2, eta_1z, eta_1r, eta_2z, eta_2r,
                         per_communicator1, per_communicator2):
    """elasto_kernels/solid_sigma.py:32-60"""
    per_communicator2(eta_1)
    per_communicator1(eta_2)
    solid_sigma(sigma_s_11, sigma_s_12, sigma_s_22, G, dx, eta_1, eta_2, eta_1z, eta_1r, eta_2z, eta_2r)


def update_vorticity_from_solid_stress(vorticity, tau_z, tau_r, tau11, tau12, tau22, R, dt, dx, _per=None):
    """elasto_kernels/div_tau.py:4-34"""
    if _per is not None:
        for f in (tau11, tau12, tau22):
            _per(f)
    st = Stage()
    (w, tz, tr), (a, b, c), ld = st.fields(outs=[vorticity, tau_z, tau_r], ins=[tau11, tau12, tau22])
    g = make_grid(w.shape[0], w.shape[1], ld, dx)
    r1 = coord_1d(st, R, 0, w.shape[0])
    _call("axb_solid_tau", ctypes.byref(g), ptr(tz), ptr(tr), ptr(a), ptr(b), ptr(c), ptr(r1), stream_ptr())
    if _per is not None:
        _per(DeviceField(tr))
        _per(DeviceField(tz))
    _call("axb_solid_vorticity_update", ctypes.byref(g), ptr(w), ptr(tz), ptr(tr), float(dt), None, stream_ptr())
    st.finish()


def update_vorticity_from_solid_stress_periodic(vorticity, tau_z, tau_r, tau11, tau12, tau22, R, dt, dx,
                                                per_communicator):
    """elasto_kernels/div_tau.py:37-72"""
    update_vorticity_from_solid_stress(vorticity, tau_z, tau_r, tau11, tau12, tau22, R, dt, dx, _per=per_communicator)


# --------------------------------------------------------------------------------------
# core/extrapolate_using_least_squares, elasto_kernels/extrapolate_*
# --------------------------------------------------------------------------------------
_ls_work = {}


def _ls_workspace(n0, n1):
    nbytes = int(_call("axb_ls_workspace_bytes", n0, n1))
    w = _ls_work.get("buf")
    if w is None or w.numel() < nbytes:
        w = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        _ls_work["buf"] = w
    return w, nbytes


def extrapolate_using_least_squares_till_first_order(current_flag, target_flag, eta_x, eta_y, grid_x, grid_y):
    """core/src/extrapolate_using_least_squares.hpp:450-467 (pybind signature of
    core/src/extrapolate_using_least_squares_bind.cpp:11-42; flags int16, exact dtypes required)."""
    return _ls_extrapolate("axb_ls_extrapolate_order1", current_flag, target_flag, eta_x, eta_y, grid_x, grid_y)


def extrapolate_using_least_squares_till_second_order(current_flag, target_flag, eta_x, eta_y, grid_x, grid_y):
    """core/src/extrapolate_using_least_squares.hpp:469-486 (binding extrapolate_using_least_squares_bind.cpp:44-75):
    quadratic basis on the same 3 x 3 patch"""
    return _ls_extrapolate("axb_ls_extrapolate_order2", current_flag, target_flag, eta_x, eta_y, grid_x, grid_y)


def _ls_extrapolate(entry, current_flag, target_flag, eta_x, eta_y, grid_x, grid_y):
    st = Stage()
    cur = st.dev(current_flag, out=True, dtype=torch.int16)
    tgt = st.dev(target_flag, dtype=torch.int16)
    ex, ey = st.dev(eta_x, out=True), st.dev(eta_y, out=True)
    for t in (cur, tgt, ex, ey):
        if not t.is_contiguous():
            raise ValueError("core routines read .data() and need C-contiguous arrays, like the reference")
    n0, n1 = cur.shape
    gx, gy = coord_1d(st, grid_x, 1, n1), coord_1d(st, grid_y, 0, n0)
    work, nbytes = _ls_workspace(n0, n1)
    sweeps = ctypes.c_int(0)
    _call(entry, n0, n1, ptr(cur), ptr(tgt), ptr(ex), ptr(ey), ptr(gx), ptr(gy), ptr(work),
          nbytes, 0, ctypes.byref(sweeps), stream_ptr())
    st.finish()
    return sweeps.value


def extrapolate_eta_using_least_squares(inp_phi, phi_thresh_lower_bound, phi_thresh_upper_bound, inp_eta_X,
                                        inp_eta_Y, inp_x, inp_y):
    """elasto_kernels/extrapolate_using_least_squares.py:8-38"""
    st = Stage()
    phi = st.dev(inp_phi)
    cur = (phi < phi_thresh_lower_bound).to(torch.int16)
    tgt = (phi < phi_thresh_upper_bound).to(torch.int16)
    extrapolate_using_least_squares_till_first_order(cur, tgt, inp_eta_X, inp_eta_Y, inp_x, inp_y)


def extrapolate_eta_with_least_squares(inside_solid, ball_phi, eta1, eta2, ball_phi_double, eta1_double,
                                       eta2_double, extrap_zone, grid_size_r, z):
    """elasto_kernels/extrapolate_eta_using_least_squares_unb.py:7-30.  The three ``*_double``
    scratch arrays of the reference signature are accepted and ignored: the doubled staging
    lives in the library workspace."""
    st = Stage()
    (e1, e2), (phi,), ld = st.fields(outs=[eta1, eta2], ins=[ball_phi])
    ins = st.dev(inside_solid, dtype=torch.uint8)
    if not ins.is_contiguous():
        ins = ins.contiguous()
    nr, nz = e1.shape
    g = make_grid(nr, nz, ld, 1.0)
    gx, gy = coord_1d(st, z, 1, nz), coord_1d(st, z, 0, 2 * nr)
    work, nbytes = _ls_workspace(2 * nr, nz)
    sweeps = ctypes.c_int(0)
    _call("axb_ls_extrapolate_eta", ctypes.byref(g), ptr(phi), ptr(ins), ptr(e1), ptr(e2), float(extrap_zone),
          ptr(gx), ptr(gy), ptr(work), nbytes, 0, ctypes.byref(sweeps), stream_ptr())
    st.finish()
    return sweeps.value


# --------------------------------------------------------------------------------------
# core/particles_to_mesh, kernels/advect_particle.py
# --------------------------------------------------------------------------------------
def _p2m(px, py, val, mesh, dx, dy, periodic):
    st = Stage()
    tx, ty, tv = st.dev(px), st.dev(py), st.dev(val)
    tm = st.dev(mesh, out=True)
    for t in (tx, ty, tv, tm):
        if t.ndim != 2 or not t.is_contiguous():
            raise ValueError("core routines read .data() and need C-contiguous 2-D arrays, like the reference")
    n0, n1 = tm.shape
    if tuple(tx.shape) == (n0, n1):      # lattice-shaped particle arrays (the drivers' case): the 2-D launch
        _call("axb_p2m_mp4_2d", n0, n1, ptr(tx), ptr(ty), ptr(tv), ptr(tm), float(dx), float(dy), int(periodic),
              stream_ptr())
    else:                                # any particle-array shape, like the reference (csrc/particles.cu)
        _call("axb_p2m_2d", 1, n0, n1, tx.shape[0], tx.shape[1], ptr(tx), ptr(ty), ptr(tv), ptr(tm), float(dx),
              float(dy), int(periodic), stream_ptr())
    st.finish()


def particles_to_mesh_2D_unbounded_mp4(particle_positions_x, particle_positions_y,
                                       input_field_at_particle_positions, output_field_at_mesh, delta_x, delta_y):
    """core/src/particles_to_mesh.hpp:163-184 (binding core/src/particles_to_mesh_bind.cpp:209-235)"""
    _p2m(particle_positions_x, particle_positions_y, input_field_at_particle_positions, output_field_at_mesh,
         delta_x, delta_y, False)


def particles_to_mesh_2D_mp4(particle_positions_x, particle_positions_y, input_field_at_particle_positions,
                             output_field_at_mesh, delta_x, delta_y):
    """periodic twin (core/src/interpolation/particles_to_mesh_2D.hpp:13-148)"""
    _p2m(particle_positions_x, particle_positions_y, input_field_at_particle_positions, output_field_at_mesh,
         delta_x, delta_y, True)


def _advect_particles(z_particles, r_particles, vort_particles, vorticity, Z_double, R_double, grid_size_r, u_z,
                      u_r, dx, dt, periodic):
    st = Stage()
    nr = grid_size_r
    zp, rp, wp = st.dev(z_particles, out=True), st.dev(r_particles, out=True), st.dev(vort_particles, out=True)
    w, uz, ur = st.dev(vorticity, out=True), st.dev(u_z), st.dev(u_r)
    Zd, Rd = st.dev(Z_double), st.dev(R_double)
    # push + mirror (kernels/advect_particle.py:21-28): tiny glue on the doubled arrays
    zp[nr:] += uz * dt
    zp[:nr] += torch.flip(uz, [0]) * dt
    rp[nr:] += ur * dt
    rp[:nr] += -torch.flip(ur, [0]) * dt
    wp[nr:] = w
    wp[:nr] = -torch.flip(w, [0])
    mesh = torch.empty_like(Zd)
    n0, n1 = mesh.shape
    _call("axb_p2m_mp4_2d", n0, n1, ptr(zp), ptr(rp), ptr(wp), ptr(mesh), float(dx), float(dx), int(periodic),
          stream_ptr())
    zp.copy_(Zd)
    rp.copy_(Rd)
    w.copy_(mesh[nr:])
    st.finish()


def advect_vorticity_via_particles(z_particles, r_particles, vort_particles, vorticity, Z_double, R_double,
                                   grid_size_r, u_z, u_r, dx, dt):
    """kernels/advect_particle.py:5-35"""
    _advect_particles(z_particles, r_particles, vort_particles, vorticity, Z_double, R_double, grid_size_r, u_z,
                      u_r, dx, dt, False)


def advect_vorticity_via_particles_periodic(z_particles, r_particles, vort_particles, vorticity, Z_double,
                                            R_double, grid_size_r, u_z, u_r, dx, dt):
    """kernels/advect_particle.py:38-68"""
    _advect_particles(z_particles, r_particles, vort_particles, vorticity, Z_double, R_double, grid_size_r, u_z,
                      u_r, dx, dt, True)


def advect_vorticity_via_lattice_particles(vorticity_out, vorticity_in, u_z, u_r, z_lattice, r_lattice_double, dx,
                                           dt, periodic=False):
    """Fused G-P2M for lattice particles (what advect_particle.py always starts from): push by
    u*dt, MP4 remesh, physical half only -- no doubled arrays, 40 B/pt."""
    st = Stage()
    (wo,), (wi, uz, ur), ld = st.fields(outs=[vorticity_out], ins=[vorticity_in, u_z, u_r])
    nr, nz = wo.shape
    g = make_grid(nr, nz, ld, dx)
    zl, rl = coord_1d(st, z_lattice, 1, nz), coord_1d(st, r_lattice_double, 0, 2 * nr)
    _call("axb_advect_vorticity_particles", ctypes.byref(g), ptr(wo), ptr(wi), ptr(uz), ptr(ur), ptr(zl), ptr(rl),
          float(dt), None, int(periodic), stream_ptr())
    st.finish()
