"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run from the repo root:   python tests/golden/make_golden.py
Needs /root/reference (imported, never copied) and the reference C++ core compiled by
``make -C oracle ref`` into oracle/_ref/.  The GPU box has neither; it only reads the
committed .npz files.  Every fixture stores the seeded inputs AND the reference outputs so
tests never have to re-create the inputs bit-for-bit.

pystencils is not installable offline, so no fixture exists for the ENO3 kernels
(pyst_kernels/*): parity unpinned there, see oracle/axisym_oracle.py.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AXB_REFERENCE", "/root/reference")
warnings.filterwarnings("ignore")
sys.path.insert(0, REF)


def _load_ref_core():
    """Expose oracle/_ref/*.so under the names the reference's Python wrappers import."""
    import pyaxisymflow  # noqa: F401  (the reference package)
    import pyaxisymflow.core as core

    refdir = os.path.join(ROOT, "oracle", "_ref")
    for name in ("particles_to_mesh", "extrapolate_using_least_squares"):
        path = [f for f in os.listdir(refdir) if f.startswith(name + ".")][0]
        spec = importlib.util.spec_from_file_location(name, os.path.join(refdir, path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["pyaxisymflow.core." + name] = mod
        setattr(core, name, mod)


def grid(nr, nz, dx=None):
    dx = 1.0 / nz if dx is None else dx
    z = np.linspace(dx / 2, nz * dx - dx / 2, nz)
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    Z, R = np.meshgrid(z, r)
    return dx, z, r, Z, R


def smooth(rng, Z, R, amp=1.0):
    """band-limited blob + a little white noise (exercises every stencil term)."""
    f = amp * np.sin(2 * np.pi * (Z + 0.3 * R)) * np.exp(-((Z - 0.5) ** 2 + R ** 2) / 0.05)
    return f + 0.05 * amp * rng.standard_normal(Z.shape)


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(f"wrote {name}.npz  ({sum(a.nbytes for a in arrays.values() if hasattr(a, 'nbytes'))} B raw)")


def main():
    _load_ref_core()
    from pyaxisymflow.kernels.brinkmann_penalize import brinkmann_penalize
    from pyaxisymflow.kernels.compute_velocity_from_psi import (
        compute_velocity_from_psi_periodic, compute_velocity_from_psi_unb)
    from pyaxisymflow.kernels.compute_vorticity_from_velocity import (
        compute_vorticity_from_velocity_periodic, compute_vorticity_from_velocity_unb)
    from pyaxisymflow.kernels.diffusion_RK2 import diffusion_RK2_periodic, diffusion_RK2_unb
    from pyaxisymflow.kernels.kill_boundary_vorticity_sine import (
        kill_boundary_vorticity_sine_r, kill_boundary_vorticity_sine_z)
    from pyaxisymflow.kernels.periodic_boundary_ghost_comm import (
        gen_periodic_boundary_ghost_comm, gen_periodic_boundary_ghost_comm_eta)
    from pyaxisymflow.kernels.smooth_Heaviside import smooth_Heaviside
    from pyaxisymflow.kernels.vortex_stretching import vortex_stretching
    from pyaxisymflow.kernels.compute_forces import compute_force_on_body
    from pyaxisymflow.kernels.force_projection import force_projection
    from pyaxisymflow.kernels.FastDiagonalisationStokesSolver import FastDiagonalisationStokesSolver
    from pyaxisymflow.kernels.FastDiagonalisationPotentialSolver import FastDiagonalisationPotentialSolver
    from pyaxisymflow.kernels.implicit_diffusion_solver import ImplicitEulerDiffusionStepper
    from pyaxisymflow.kernels.advect_particle import (
        advect_vorticity_via_particles, advect_vorticity_via_particles_periodic)
    from pyaxisymflow.elasto_kernels.solid_sigma import solid_sigma
    from pyaxisymflow.elasto_kernels.div_tau import update_vorticity_from_solid_stress
    from pyaxisymflow.elasto_kernels.extrapolate_eta_using_least_squares_unb import (
        extrapolate_eta_with_least_squares)
    import pyaxisymflow.core.particles_to_mesh as p2m
    import pyaxisymflow.core.extrapolate_using_least_squares as els

    rng = np.random.default_rng(20261017)
    nr, nz = 24, 56
    dx, z, r, Z, R = grid(nr, nz)
    per = gen_periodic_boundary_ghost_comm(2)

    # ---- a9 brinkmann (scalar and field body velocity) ---------------------------------
    chi = np.clip(smooth(rng, Z, R) + 0.5, 0, 1)
    uz, ur = smooth(rng, Z, R), smooth(rng, Z, R)
    Uz_f, Ur_f = smooth(rng, Z, R), smooth(rng, Z, R)
    out = {}
    for tag, (Uz, Ur) in {"scalar": (0.7, -0.2), "field": (Uz_f, Ur_f)}.items():
        pz, pr = np.zeros_like(Z), np.zeros_like(Z)
        brinkmann_penalize(1e4, 3e-3, chi, Uz, Ur, uz, ur, pz, pr)
        out[f"pz_{tag}"], out[f"pr_{tag}"] = pz, pr
    save("brinkmann", lam=1e4, dt=3e-3, chi=chi, uz=uz, ur=ur, Uz_f=Uz_f, Ur_f=Ur_f,
         Uz_s=0.7, Ur_s=-0.2, **out)

    # ---- a8 diffusion -------------------------------------------------------------------
    w0 = smooth(rng, Z, R, 3.0)
    nu, dt = 2e-3, 0.2 * dx * dx / 2e-3
    res = {}
    for tag in ("unb", "periodic"):
        w, tmp = w0.copy(), rng.standard_normal(Z.shape)
        if tag == "unb":
            diffusion_RK2_unb(w, tmp, R, nu, dt, dx)
        else:
            diffusion_RK2_periodic(w, tmp, R, nu, dt, dx, per)
        res[f"w_{tag}"], res[f"tmp_{tag}"] = w, tmp
    save("diffusion", w0=w0, nu=nu, dt=dt, dx=dx, **res)

    # ---- a11 / a10 ----------------------------------------------------------------------
    psi0 = smooth(rng, Z, R)
    res = {}
    for tag in ("unb", "periodic"):
        psi = psi0.copy()
        a, b = rng.standard_normal(Z.shape), rng.standard_normal(Z.shape)
        if tag == "unb":
            compute_velocity_from_psi_unb(a, b, psi, R, dx)
        else:
            compute_velocity_from_psi_periodic(a, b, psi, R, dx, per)
        res[f"uz_{tag}"], res[f"ur_{tag}"], res[f"psi_{tag}"] = a, b, psi
    save("velocity_from_psi", psi0=psi0, dx=dx, **res)

    res = {}
    vort_init = rng.standard_normal(Z.shape)
    for tag in ("unb", "periodic"):
        a, b, v = uz.copy(), ur.copy(), vort_init.copy()
        if tag == "unb":
            compute_vorticity_from_velocity_unb(v, a, b, dx)
        else:
            compute_vorticity_from_velocity_periodic(v, a, b, dx, per)
        res[f"vort_{tag}"], res[f"uz_{tag}"], res[f"ur_{tag}"] = v, a, b
    save("vorticity_from_velocity", uz=uz, ur=ur, vort_init=vort_init, dx=dx, **res)

    # ---- a12 ghost comm (both flavours) ---------------------------------------------------
    f0 = smooth(rng, Z, R)
    fa, fb = f0.copy(), f0.copy()
    per(fa)
    gen_periodic_boundary_ghost_comm_eta(2, 1.0, dx)(fb)
    save("ghost_comm", f0=f0, plain=fa, eta=fb, z_max=1.0, dx=dx)

    # ---- a13 kill boundary ----------------------------------------------------------------
    w = w0.copy()
    kill_boundary_vorticity_sine_z(w, Z, 3, dx)
    wz = w.copy()
    kill_boundary_vorticity_sine_r(w, R, 3, dx)
    save("kill_boundary", w0=w0, after_z=wz, after_zr=w, dx=dx)

    # ---- a14 Heaviside ----------------------------------------------------------------------
    phi = 0.2 - np.sqrt((Z - 0.45) ** 2 + R ** 2)
    H = rng.standard_normal(Z.shape)
    smooth_Heaviside(H, phi, dx * 2 ** 0.5)
    save("heaviside", phi=phi, w=dx * 2 ** 0.5, H=H)

    # ---- vortex stretching + reductions (a15) ----------------------------------------------
    w = w0.copy()
    vortex_stretching(w, ur, R, 1e-3)
    F = compute_force_on_body(R, chi, 1.3, 1e4, uz, 0.25, 0.01, 1e-3, 0.02)
    P = force_projection(2.0, chi, uz, ur, R)
    save("misc", w0=w0, ur=ur, uz=uz, chi=chi, stretched=w, dt=1e-3,
         F_pen=F[0], F_un=F[1], proj_z=P[0], proj_r=P[1])

    # ---- a16 fast diagonalisation ------------------------------------------------------------
    rhs = smooth(rng, Z, R, 5.0)
    res = {}
    for bc in ("homogenous_neumann_along_z_and_r", "homogenous_neumann_along_r_and_periodic_along_z",
               "homogenous_dirichlet_along_r_and_periodic_along_z"):
        s = FastDiagonalisationStokesSolver(nr, nz, dx, bc_type=bc)
        sol = np.zeros_like(Z)
        s.solve(sol, rhs)
        res["stokes_" + bc] = sol
    # strided right-hand side, as examples/PeriodicFlowPastSphere/periodic_flow_past_sphere.py:100-104
    s = FastDiagonalisationStokesSolver(nr, nz - 4, dx, bc_type="homogenous_neumann_along_r_and_periodic_along_z")
    sol = np.zeros((nr, nz - 4))
    s.solve(sol, rhs[:, 2:-2])
    res["stokes_periodic_inner"] = sol
    s = FastDiagonalisationPotentialSolver(nr, nz, dx)
    sol = np.zeros_like(Z)
    s.solve(sol, rhs)
    res["potential"] = sol
    nu_dt = 0.3 * dx * dx
    s = ImplicitEulerDiffusionStepper(nu_dt / 2e-3, 2e-3, nr, nz, dx)
    w = rhs.copy()
    s.step(w, nu_dt / 2e-3)
    res["implicit_diffusion"] = w
    save("fast_diag", rhs=rhs, dx=dx, nu=2e-3, time_step=nu_dt / 2e-3, **res)

    # ---- a18 / a19 solid stress ------------------------------------------------------------------
    e1 = Z + 0.02 * smooth(rng, Z, R)
    e2 = R + 0.02 * smooth(rng, Z, R)
    names = ["s11", "s12", "s22", "e1z", "e1r", "e2z", "e2r"]
    init = {n: rng.standard_normal(Z.shape) for n in names}
    o = {n: init[n].copy() for n in names}
    G = 3.7
    solid_sigma(o["s11"], o["s12"], o["s22"], G, dx, e1, e2, o["e1z"], o["e1r"], o["e2z"], o["e2r"])
    t_init = {n: rng.standard_normal(Z.shape) for n in ("tau_z", "tau_r")}
    tz, tr, w = t_init["tau_z"].copy(), t_init["tau_r"].copy(), w0.copy()
    s11c, s12c, s22c = chi * o["s11"], chi * o["s12"], chi * o["s22"]
    update_vorticity_from_solid_stress(w, tz, tr, s11c, s12c, s22c, R, 2e-3, dx)
    save("solid", eta1=e1, eta2=e2, G=G, dx=dx, dt=2e-3, chi=chi, w0=w0,
         **{"init_" + k: v for k, v in init.items()}, **{"out_" + k: v for k, v in o.items()},
         init_tau_z=t_init["tau_z"], init_tau_r=t_init["tau_r"], out_tau_z=tz, out_tau_r=tr, out_w=w)

    # ---- a20 LS extrapolation (needs 2 Nr == Nz) ---------------------------------------------------
    nr2, nz2 = 32, 64
    dx2, z2, r2, Z2, R2 = grid(nr2, nz2)
    moll = 2 * dx2
    zone = moll + 4 * dx2
    ball_phi = 0.2 - np.sqrt((Z2 - 0.5) ** 2 + R2 ** 2)
    Hs = 0 * Z2
    smooth_Heaviside(Hs, ball_phi, moll)
    inside = Hs > 0.5
    eta1_in = Z2 + 0.03 * np.sin(7 * Z2) * np.cos(5 * R2)
    eta2_in = R2 * (1 + 0.05 * np.cos(6 * Z2))
    eta1, eta2 = eta1_in.copy(), eta2_in.copy()
    pd, e1d, e2d = (np.zeros((2 * nr2, nz2)) for _ in range(3))
    extrapolate_eta_with_least_squares(inside, ball_phi, eta1, eta2, pd, e1d, e2d, zone, nr2, z2)
    # raw core call on irregular flags (two blobs, order-1)
    n0, n1 = 40, 48
    yy, xx = np.meshgrid(np.arange(n0), np.arange(n1), indexing="ij")
    blob = ((xx - 18) ** 2 + (yy - 20) ** 2 < 36) | ((xx - 30) ** 2 + (yy - 17) ** 2 < 20)
    band = ((xx - 18) ** 2 + (yy - 20) ** 2 < 120) | ((xx - 30) ** 2 + (yy - 17) ** 2 < 90)
    cur = blob.astype(np.int16)
    tgt = band.astype(np.int16)
    gx = np.linspace(0.1, 1.3, n1)
    gy = np.linspace(-0.4, 0.9, n0)
    ex = np.where(blob, 1.5 * gx[None, :] - 0.7 * gy[:, None] + 0.1 * np.sin(9 * gx[None, :]), 0.0)
    ey = np.where(blob, np.cos(3 * gy[:, None]) + gx[None, :] ** 2, 0.0)
    cur_o, ex_o, ey_o = cur.copy(), ex.copy(), ey.copy()
    els.extrapolate_using_least_squares_till_first_order(cur_o, tgt, ex_o, ey_o, gx, gy)
    save("ls_extrapolation", ball_phi=ball_phi, inside=inside, eta1_in=eta1_in, eta2_in=eta2_in,
         zone=zone, z=z2, eta1_out=eta1, eta2_out=eta2,
         raw_cur=cur, raw_tgt=tgt, raw_ex=ex, raw_ey=ey, raw_gx=gx, raw_gy=gy,
         raw_cur_out=cur_o, raw_ex_out=ex_o, raw_ey_out=ey_o)

    # ---- a21 MP4 particles-to-mesh ----------------------------------------------------------------
    n0, n1 = 2 * nr, nz
    zd = np.linspace(dx / 2, 1 - dx / 2, nz)
    rd = np.linspace(-nr * dx + dx / 2, nr * dx - dx / 2, 2 * nr)
    Zd, Rd = np.meshgrid(zd, rd)
    # raw call: positions in mesh-index space [0, n) * dx, displaced by up to 1.7 cells
    px = (np.arange(n1)[None, :] + 0.5 + 1.7 * rng.uniform(-1, 1, (n0, n1))) * dx
    py = (np.arange(n0)[:, None] + 0.5 + 1.7 * rng.uniform(-1, 1, (n0, n1))) * dx
    val = rng.standard_normal((n0, n1))
    mesh_unb, mesh_per = np.ones((n0, n1)), np.ones((n0, n1))
    p2m.particles_to_mesh_2D_unbounded_mp4(px, py, val, mesh_unb, dx, dx)
    pxw, pyw = np.mod(px, n1 * dx), np.mod(py, n0 * dx)
    p2m.particles_to_mesh_2D_mp4(pxw, pyw, val, mesh_per, dx, dx)
    # wrapper: one remeshed-particle advection step, as kernels/advect_particle.py:5-35
    # NB the reference passes physical coordinates straight to the mesh routine, so the
    # doubled "r" coordinate must be the non-negative lattice the examples use
    # (examples/ParticleOscillatoryFlowCases/particle_in_bubble_oscillatory_flow.py).
    Zl, Rl = np.meshgrid(zd, np.linspace(dx / 2, 2 * nr * dx - dx / 2, 2 * nr))
    zp, rp, wp = Zl.copy(), Rl.copy(), 0 * Zl
    w = w0.copy()
    dtp = 0.4 * dx / np.amax(np.abs(uz) + np.abs(ur))
    advect_vorticity_via_particles(zp, rp, wp, w, Zl, Rl, nr, uz, ur, dx, dtp)
    save("p2m", dx=dx, px=px, py=py, val=val, mesh_unb=mesh_unb, pxw=pxw, pyw=pyw, mesh_per=mesh_per,
         Zl=Zl, Rl=Rl, w0=w0, uz=uz, ur=ur, dt=dtp, w_adv=w, wp_after=wp)


if __name__ == "__main__":
    main()
