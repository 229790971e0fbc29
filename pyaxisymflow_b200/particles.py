"""Particle <-> mesh interpolation family of the reference's C++ core on sm_100a kernels (``csrc/particles.cu``).

Every function of ``pyaxisymflow/core/src/instantiate.yml:1-27`` under its reference name, argument order and
in-place semantics (bindings ``core/src/mesh_to_particles_bind.cpp:394-458``, ``particles_to_mesh_bind.cpp:286-340``):
arrays must be float64 and C-contiguous (the bindings use ``noconvert`` and read ``.data()``), outputs are
overwritten in place, nothing is returned.  NumPy arrays are staged through the GPU (parity mode), CUDA tensors /
``DeviceField`` are used in place.  There is no CPU path.
"""
import numpy as np  # noqa: F401

from . import _lib
from .device import Stage, ptr, stream_ptr

_call = _lib.call
_KERNELS = {"linear_kernel": 0, "mp4": 1, "mp6": 2, "yang_smooth_three_point_kernel": 3}


def _dense(st, a, out=False, ndim=2):
    t = st.dev(a, out=out)
    if t.ndim != ndim or not t.is_contiguous():
        raise ValueError(f"core routines read .data() and need C-contiguous {ndim}-D arrays, like the reference")
    return t


def _m2p_2d(kernel, periodic):
    def mesh_to_particles(input_field_x, input_field_y, particle_positions_x, particle_positions_y, output_field_x,
                          output_field_y, delta_x, delta_y):
        st = Stage()
        fx, fy = _dense(st, input_field_x), _dense(st, input_field_y)
        px, py = _dense(st, particle_positions_x), _dense(st, particle_positions_y)
        ox, oy = _dense(st, output_field_x, out=True), _dense(st, output_field_y, out=True)
        _call("axb_m2p_2d", kernel, fx.shape[0], fx.shape[1], ptr(fx), ptr(fy), px.shape[0], px.shape[1], ptr(px),
              ptr(py), ptr(ox), ptr(oy), float(delta_x), float(delta_y), int(periodic), stream_ptr())
        st.finish()

    return mesh_to_particles


def _p2m_2d(kernel, periodic):
    def particles_to_mesh(particle_positions_x, particle_positions_y, input_field_at_particle_positions,
                          output_field_at_mesh, delta_x, delta_y):
        st = Stage()
        px, py = _dense(st, particle_positions_x), _dense(st, particle_positions_y)
        v, m = _dense(st, input_field_at_particle_positions), _dense(st, output_field_at_mesh, out=True)
        _call("axb_p2m_2d", kernel, m.shape[0], m.shape[1], px.shape[0], px.shape[1], ptr(px), ptr(py), ptr(v), ptr(m),
              float(delta_x), float(delta_y), int(periodic), stream_ptr())
        st.finish()

    return particles_to_mesh


M2P, P2M = {}, {}
for _name, _kid in _KERNELS.items():
    for _per in (True, False):
        _mid = "" if _per else "unbounded_"
        _f = _m2p_2d(_kid, _per)
        _f.__name__ = _f.__qualname__ = f"mesh_to_particles_2D_{_mid}{_name}"
        _f.__doc__ = f"core/src/mesh_to_particles.hpp ({_f.__name__}), binding mesh_to_particles_bind.cpp:416-458"
        M2P[_f.__name__] = _f
        _g = _p2m_2d(_kid, _per)
        _g.__name__ = _g.__qualname__ = f"particles_to_mesh_2D_{_mid}{_name}"
        _g.__doc__ = f"core/src/particles_to_mesh.hpp ({_g.__name__}), binding particles_to_mesh_bind.cpp:291-340"
        P2M[_g.__name__] = _g
globals().update(M2P)
globals().update(P2M)


def mesh_to_particles_1D_mp4(input_field, particle_positions, output_field, delta_x):
    """core/src/mesh_to_particles.hpp:25-37 (periodic, MP4)"""
    st = Stage()
    f, p, o = _dense(st, input_field, ndim=1), _dense(st, particle_positions, ndim=1), _dense(st, output_field, True, 1)
    _call("axb_m2p_1d_mp4", f.shape[0], ptr(f), p.shape[0], ptr(p), ptr(o), float(delta_x), stream_ptr())
    st.finish()


def particles_to_mesh_1D_mp4(particle_positions, input_field_at_particle_positions, output_field, delta_x):
    """core/src/particles_to_mesh.hpp:8-20 (periodic, MP4)"""
    st = Stage()
    p, v = _dense(st, particle_positions, ndim=1), _dense(st, input_field_at_particle_positions, ndim=1)
    m = _dense(st, output_field, True, 1)
    _call("axb_p2m_1d_mp4", m.shape[0], p.shape[0], ptr(p), ptr(v), ptr(m), float(delta_x), stream_ptr())
    st.finish()


def wrap_particles_around_1D_domain(particle_positions, domain_start, domain_end):
    """core/src/mesh_to_particles.hpp:14-22: only the first / last 10 particles are tested (WrappingStrategy FirstN)"""
    st = Stage()
    p = _dense(st, particle_positions, True, 1)
    _call("axb_wrap_particles_2d", 1, p.shape[0], ptr(p), None, float(domain_start), float(domain_end), 0.0, 0.0,
          stream_ptr())
    st.finish()


def wrap_particles_around_2D_domain(particle_positions_x, particle_positions_y, domain_start_x, domain_end_x,
                                    domain_start_y, domain_end_y):
    """core/src/mesh_to_particles.hpp:39-55"""
    st = Stage()
    px, py = _dense(st, particle_positions_x, True), _dense(st, particle_positions_y, True)
    _call("axb_wrap_particles_2d", px.shape[0], px.shape[1], ptr(px), ptr(py), float(domain_start_x),
          float(domain_end_x), float(domain_start_y), float(domain_end_y), stream_ptr())
    st.finish()
