"""Generate tests/golden/core_family.npz and static_pde.npz from the UNMODIFIED reference (build container only):

  * every function of pyaxisymflow/core/src/instantiate.yml -- the reference C++ compiled as-is by
    ``make -C oracle ref`` into oracle/_ref/*.so (mesh_to_particles, particles_to_mesh,
    extrapolate_using_least_squares) -- on seeded inputs;
  * ``StaticPDEExtrapolation`` imported from examples/PeriodicSoftSlab/bounded_static_PDE_extrapolation.py.

    python tests/golden/make_golden_core.py
"""
import importlib.util
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AXB_REFERENCE", "/root/reference")
warnings.filterwarnings("ignore")
sys.path.insert(0, REF)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ref_mod(name):
    refdir = os.path.join(ROOT, "oracle", "_ref")
    path = [f for f in os.listdir(refdir) if f.startswith(name + ".")][0]
    return _load(os.path.join(refdir, path), name)


def main():
    m2p, p2m, els = _ref_mod("mesh_to_particles"), _ref_mod("particles_to_mesh"), _ref_mod("extrapolate_using_least_squares")
    rng = np.random.default_rng(20261019)
    out = {}
    m0, m1 = 26, 34
    dx, dy = 1.0 / 32, 1.0 / 32
    fx, fy = rng.standard_normal((m0, m1)), rng.standard_normal((m0, m1))
    p0, p1 = 19, 23
    # particles anywhere within ~2.4 cells of the mesh box: the clipped ("unbounded") stencils get exercised on
    # every side; the periodic functions take the same particles folded into the box
    px = rng.uniform(-0.4, m1 + 0.4, (p0, p1)) * dx
    py = rng.uniform(-0.4, m0 + 0.4, (p0, p1)) * dy
    px.flat[:8] = np.array([0.0, 0.5, 1.0, 1.5, m1 - 0.5, m1 - 1.0, m1 - 1.5, 2.5]) * dx     # exact half-cell ties
    py.flat[:8] = np.array([0.5, 0.0, 1.5, 1.0, 2.5, m0 - 0.5, m0 - 1.5, m0 - 1.0]) * dy
    pxw, pyw = np.mod(px, m1 * dx), np.mod(py, m0 * dy)
    val = rng.standard_normal((p0, p1))
    out.update(fx=fx, fy=fy, px=px, py=py, pxw=pxw, pyw=pyw, val=val, dx=dx, dy=dy)
    for k in ("linear_kernel", "mp4", "mp6", "yang_smooth_three_point_kernel"):
        for per in (True, False):
            mid = "" if per else "unbounded_"
            qx, qy = (pxw, pyw) if per else (px, py)
            ox, oy = np.full((p0, p1), 7.0), np.full((p0, p1), 7.0)
            getattr(m2p, f"mesh_to_particles_2D_{mid}{k}")(fx, fy, qx, qy, ox, oy, dx, dy)
            mesh = np.full((m0, m1), 7.0)
            getattr(p2m, f"particles_to_mesh_2D_{mid}{k}")(qx, qy, val, mesh, dx, dy)
            out[f"m2p_{mid}{k}_x"], out[f"m2p_{mid}{k}_y"], out[f"p2m_{mid}{k}"] = ox, oy, mesh
    f1 = rng.standard_normal(40)
    q1 = rng.uniform(0, 40, 57) / 32
    o1 = np.zeros(57)
    m2p.mesh_to_particles_1D_mp4(f1, q1, o1, dx)
    v1, mesh1 = rng.standard_normal(57), np.ones(40)
    p2m.particles_to_mesh_1D_mp4(q1, v1, mesh1, dx)
    out.update(f1=f1, q1=q1, m2p_1d=o1, v1=v1, p2m_1d=mesh1)
    # wrap: 2-D (x: first / last 10 entries of a row; y: first / last 10 rows) and 1-D
    wx0 = rng.uniform(-0.3, 1.3, (27, 31))
    wy0 = rng.uniform(-0.2, 0.8, (27, 31))
    wx, wy = wx0.copy(), wy0.copy()
    m2p.wrap_particles_around_2D_domain(wx, wy, 0.0, 1.0, 0.0, 0.5)
    w1 = rng.uniform(-0.3, 1.3, 33)
    w1o = w1.copy()
    m2p.wrap_particles_around_1D_domain(w1o, 0.0, 1.0)
    ws = rng.uniform(-0.3, 1.3, (6, 7))        # fewer than 10 rows / columns
    wsx, wsy = ws.copy(), ws.T.copy().T.copy()
    m2p.wrap_particles_around_2D_domain(wsx, wsy, 0.0, 1.0, 0.0, 1.0)
    out.update(wrap_x0=wx0, wrap_y0=wy0, wrap_x=wx, wrap_y=wy, wrap1_in=w1, wrap1_out=w1o, wrap_small_in=ws,
               wrap_small_x=wsx, wrap_small_y=wsy)

    # ---- second-order least-squares extrapolation (extrapolate_using_least_squares.hpp:469-486) ----
    n0, n1 = 44, 52
    yy, xx = np.meshgrid(np.arange(n0), np.arange(n1), indexing="ij")
    blob = ((xx - 20) ** 2 + (yy - 22) ** 2 < 64) | ((xx - 33) ** 2 + (yy - 18) ** 2 < 30)
    band = ((xx - 20) ** 2 + (yy - 22) ** 2 < 170) | ((xx - 33) ** 2 + (yy - 18) ** 2 < 110)
    cur, tgt = blob.astype(np.int16), band.astype(np.int16)
    gx, gy = np.linspace(0.1, 1.3, n1), np.linspace(-0.4, 0.9, n0)
    X, Y = gx[None, :], gy[:, None]
    ex = np.where(blob, 1.5 * X - 0.7 * Y + 0.4 * X * X - 0.3 * X * Y + 0.1 * np.sin(9 * X), 0.0)
    ey = np.where(blob, np.cos(3 * Y) + X ** 2 - 0.5 * Y * Y, 0.0)
    for order, fn in ((1, els.extrapolate_using_least_squares_till_first_order),
                      (2, els.extrapolate_using_least_squares_till_second_order)):
        c, a, b = cur.copy(), ex.copy(), ey.copy()
        fn(c, tgt, a, b, gx, gy)
        out[f"ls{order}_cur"], out[f"ls{order}_ex"], out[f"ls{order}_ey"] = c, a, b
    out.update(ls_cur=cur, ls_tgt=tgt, ls_ex=ex, ls_ey=ey, ls_gx=gx, ls_gy=gy)
    np.savez_compressed(os.path.join(HERE, "core_family.npz"), **out)
    print("wrote core_family.npz", len(out), "arrays")

    # ---- StaticPDEExtrapolation (examples/PeriodicSoftSlab/bounded_static_PDE_extrapolation.py:5-235) ----
    spe = _load(os.path.join(REF, "examples", "PeriodicSoftSlab", "bounded_static_PDE_extrapolation.py"), "spe")
    from pyaxisymflow.kernels.periodic_boundary_ghost_comm import (
        gen_periodic_boundary_ghost_comm, gen_periodic_boundary_ghost_comm_eta)

    res = {}
    nr, nz = 40, 72
    dxs = 1.0 / nz
    z = np.linspace(dxs / 2, 1 - dxs / 2, nz)
    r = np.linspace(dxs / 2, nr * dxs - dxs / 2, nr)
    Z, R = np.meshgrid(z, r)
    phi = 0.17 - np.sqrt((Z - 0.5) ** 2 + (R - 0.28) ** 2)          # positive inside, away from every wall
    eta0 = (Z + 0.05 * np.sin(9 * Z) * np.cos(7 * R)) * (phi > 0)
    e = eta0.copy()
    s = spe.StaticPDEExtrapolation(dxs, nr, nz, 1e-6, 6 * dxs)
    s.extrapolate(e, phi.copy())
    res.update(box_phi=phi, box_eta0=eta0, box_eta=e, box_tol=1e-6, box_band=6 * dxs, box_dx=dxs,
               box_bounds=np.array([s.r_start, s.r_end, s.z_start, s.z_end]))
    # solid touching the axis and a slab spanning the periodic z direction (the PeriodicSoftSlab use)
    phi2 = 0.12 - np.abs(R - 0.25) + 0.01 * np.sin(2 * np.pi * Z)
    eta2 = (R + 0.03 * np.cos(2 * np.pi * Z)) * (phi2 > 0)
    e2, p2 = eta2.copy(), phi2.copy()
    per = gen_periodic_boundary_ghost_comm(2)
    s2 = spe.StaticPDEExtrapolation(dxs, nr, nz, 1e-7, 5 * dxs, periodic=True, per_communicator_gen=per,
                                    per_communicator_eta=per)
    s2.extrapolate(e2, p2)
    res.update(slab_phi=phi2, slab_eta0=eta2, slab_eta=e2, slab_phi_after=p2, slab_tol=1e-7, slab_band=5 * dxs,
               slab_bounds=np.array([s2.r_start, s2.r_end, s2.z_start, s2.z_end]))
    np.savez_compressed(os.path.join(HERE, "static_pde.npz"), **res)
    print("wrote static_pde.npz")


if __name__ == "__main__":
    main()
