"""Pins the oracle (oracle/axisym_oracle.py + .c) against the reference.

The golden vectors were produced by the UNMODIFIED reference (tests/golden/make_golden.py);
the two known-answer tests the reference itself holds are reproduced literally.
CPU only -- runs in the `-m "not gpu"` suite.
"""
import numpy as np

from conftest import assert_close, golden
from oracle import axisym_oracle as ox

EXACT = 0.0
ULP = 1e-14  # numba's LLVM may fuse/reorder where NumPy does not


def _grid(nr, nz, dx):
    z = np.linspace(dx / 2, nz * dx - dx / 2, nz)
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    return np.meshgrid(z, r)


def test_brinkmann_reference_unit_test():
    """tests/test_kernels/test_brinkmann_penalize.py:6-26 of the reference, verbatim numbers."""
    n = 16
    lam, dt, chi_v, Uz, Ur = 2.0, 3.0, 4.0, 1.0, 2.0
    chi = np.ones((n, n)) * chi_v
    uz, ur = np.zeros((n, n)), np.zeros((n, n))
    pz, pr = np.ones((n, n)), np.ones((n, n))
    ox.brinkmann_penalize(lam, dt, chi, Uz, Ur, uz, ur, pz, pr)
    np.testing.assert_allclose(pz, lam * dt * Uz * chi_v / (1 + lam * dt * chi_v) * np.ones((n, n)))
    np.testing.assert_allclose(pr, lam * dt * Ur * chi_v / (1 + lam * dt * chi_v) * np.ones((n, n)))


def test_gauss_elimination_reference_kat():
    """core/src/lstsq/test_least_squares.cpp:30-55."""
    aug = [[2.56, 0.86, 4.2, 1.964, 1.284], [0.86, 0.32, 1.4, 0.666, 0.45], [4.2, 1.4, 7.0, 3.22, 2.1]]
    sol = ox.gauss_elimination(aug)
    assert np.max(np.abs(sol - np.array([[0.7, 0.2, 0.0], [0.3, 0.6, 0.0]]))) <= 1e-14


def test_brinkmann_golden():
    g = golden("brinkmann")
    for tag, Uz, Ur in (("scalar", float(g["Uz_s"]), float(g["Ur_s"])), ("field", g["Uz_f"], g["Ur_f"])):
        pz, pr = np.zeros_like(g["uz"]), np.zeros_like(g["uz"])
        ox.brinkmann_penalize(float(g["lam"]), float(g["dt"]), g["chi"], Uz, Ur, g["uz"], g["ur"], pz, pr)
        assert_close(pz, g[f"pz_{tag}"], ULP, "pen u_z " + tag)
        assert_close(pr, g[f"pr_{tag}"], ULP, "pen u_r " + tag)


def test_diffusion_golden():
    g = golden("diffusion")
    dx = float(g["dx"])
    Z, R = _grid(*g["w0"].shape, dx)
    for tag, ghost in (("unb", 0), ("periodic", 2)):
        w, tmp = g["w0"].copy(), np.zeros_like(g["w0"])
        ox.diffusion_RK2(w, tmp, R, float(g["nu"]), float(g["dt"]), dx, periodic_ghost=ghost)
        assert_close(w, g[f"w_{tag}"], 1e-13, "diffusion w " + tag)
        assert_close(tmp, g[f"tmp_{tag}"], 1e-13, "diffusion tmp " + tag)


def test_velocity_from_psi_golden():
    g = golden("velocity_from_psi")
    dx = float(g["dx"])
    Z, R = _grid(*g["psi0"].shape, dx)
    for tag, ghost in (("unb", 0), ("periodic", 2)):
        psi = g["psi0"].copy()
        uz, ur = np.zeros_like(psi), np.zeros_like(psi)
        ox.compute_velocity_from_psi(uz, ur, psi, R, dx, periodic_ghost=ghost)
        assert_close(uz, g[f"uz_{tag}"], ULP, "u_z " + tag)
        assert_close(ur, g[f"ur_{tag}"], ULP, "u_r " + tag)
        assert_close(psi, g[f"psi_{tag}"], EXACT, "psi ghosts " + tag)


def test_vorticity_from_velocity_golden():
    g = golden("vorticity_from_velocity")
    for tag, ghost in (("unb", 0), ("periodic", 2)):
        uz, ur, v = g["uz"].copy(), g["ur"].copy(), g["vort_init"].copy()
        ox.compute_vorticity_from_velocity(v, uz, ur, float(g["dx"]), periodic_ghost=ghost)
        assert_close(v, g[f"vort_{tag}"], ULP, "curl " + tag)
        assert_close(uz, g[f"uz_{tag}"], EXACT, "u_z ghosts")


def test_ghost_comm_golden():
    g = golden("ghost_comm")
    a, b = g["f0"].copy(), g["f0"].copy()
    ox.periodic_ghost_comm(a, 2)
    ox.periodic_ghost_comm_eta(b, 2, float(g["z_max"]), float(g["dx"]))
    assert_close(a, g["plain"], EXACT)
    assert_close(b, g["eta"], EXACT)


def test_kill_boundary_golden():
    g = golden("kill_boundary")
    dx = float(g["dx"])
    Z, R = _grid(*g["w0"].shape, dx)
    w = g["w0"].copy()
    ox.kill_boundary_vorticity_sine_z(w, Z, 3, dx)
    assert_close(w, g["after_z"], EXACT, "kill z")
    ox.kill_boundary_vorticity_sine_r(w, R, 3, dx)
    assert_close(w, g["after_zr"], EXACT, "kill r")


def test_heaviside_golden():
    g = golden("heaviside")
    H = np.ones_like(g["phi"])
    ox.smooth_Heaviside(H, g["phi"], float(g["w"]))
    assert_close(H, g["H"], ULP)


def test_misc_golden():
    g = golden("misc")
    nr, nz = g["w0"].shape
    Z, R = _grid(nr, nz, 1.0 / nz)
    w = g["w0"].copy()
    ox.vortex_stretching(w, g["ur"], R, float(g["dt"]))
    assert_close(w, g["stretched"], ULP)
    F = ox.compute_force_on_body(R, g["chi"], 1.3, 1e4, g["uz"], 0.25, 0.01, 1e-3, 0.02)
    assert abs(F[0] - float(g["F_pen"])) <= 1e-12 * abs(float(g["F_pen"]))
    assert F[1] == float(g["F_un"])
    P = ox.force_projection(2.0, g["chi"], g["uz"], g["ur"], R)
    assert abs(P[0] - float(g["proj_z"])) <= 1e-12 * abs(float(g["proj_z"]))
    assert abs(P[1] - float(g["proj_r"])) <= 1e-12 * abs(float(g["proj_r"]))


def test_fast_diagonalisation_golden():
    g = golden("fast_diag")
    rhs, dx = g["rhs"], float(g["dx"])
    nr, nz = rhs.shape
    for bc in ("homogenous_neumann_along_z_and_r", "homogenous_neumann_along_r_and_periodic_along_z",
               "homogenous_dirichlet_along_r_and_periodic_along_z"):
        s = ox.FastDiagonalisationOracle(nr, nz, dx, "stokes", bc)
        sol = np.zeros_like(rhs)
        s.solve(sol, rhs)
        assert_close(sol, g["stokes_" + bc], 1e-11, bc)
    s = ox.FastDiagonalisationOracle(nr, nz - 4, dx, "stokes", "homogenous_neumann_along_r_and_periodic_along_z")
    sol = np.zeros((nr, nz - 4))
    s.solve(sol, rhs[:, 2:-2])
    assert_close(sol, g["stokes_periodic_inner"], 1e-11, "strided rhs")
    s = ox.FastDiagonalisationOracle(nr, nz, dx, "potential")
    s.solve(sol := np.zeros_like(rhs), rhs)
    assert_close(sol, g["potential"], 1e-11, "potential")
    s = ox.FastDiagonalisationOracle(nr, nz, dx, "implicit_diffusion", nu_dt=float(g["nu"]) * float(g["time_step"]))
    s.solve(sol := np.zeros_like(rhs), rhs)
    assert_close(sol, g["implicit_diffusion"], 1e-11, "implicit diffusion")


def test_solid_golden():
    g = golden("solid")
    dx = float(g["dx"])
    nr, nz = g["eta1"].shape
    Z, R = _grid(nr, nz, dx)
    o = {k: g["init_" + k].copy() for k in ("s11", "s12", "s22", "e1z", "e1r", "e2z", "e2r")}
    ox.solid_sigma(o["s11"], o["s12"], o["s22"], float(g["G"]), dx, g["eta1"], g["eta2"],
                   o["e1z"], o["e1r"], o["e2z"], o["e2r"])
    for k, v in o.items():
        assert_close(v, g["out_" + k], 1e-13, "solid_sigma " + k)
    tz, tr, w = g["init_tau_z"].copy(), g["init_tau_r"].copy(), g["w0"].copy()
    chi = g["chi"]
    ox.update_vorticity_from_solid_stress(w, tz, tr, chi * g["out_s11"], chi * g["out_s12"], chi * g["out_s22"],
                                          R, float(g["dt"]), dx)
    assert_close(tz, g["out_tau_z"], 1e-13, "tau_z")
    assert_close(tr, g["out_tau_r"], 1e-13, "tau_r")
    assert_close(w, g["out_w"], 1e-13, "vorticity")


def test_ls_extrapolation_golden_bit_exact():
    g = golden("ls_extrapolation")
    cur, ex, ey = g["raw_cur"].copy(), g["raw_ex"].copy(), g["raw_ey"].copy()
    sweeps = ox.extrapolate_using_least_squares_till_first_order(cur, g["raw_tgt"], ex, ey, g["raw_gx"], g["raw_gy"])
    assert sweeps > 3
    assert np.array_equal(cur, g["raw_cur_out"])
    assert np.array_equal(ex, g["raw_ex_out"]), np.max(np.abs(ex - g["raw_ex_out"]))
    assert np.array_equal(ey, g["raw_ey_out"])
    e1, e2 = g["eta1_in"].copy(), g["eta2_in"].copy()
    ox.extrapolate_eta_with_least_squares(g["inside"], g["ball_phi"], e1, e2, float(g["zone"]), e1.shape[0], g["z"])
    assert np.array_equal(e1, g["eta1_out"])
    assert np.array_equal(e2, g["eta2_out"])
    assert not np.array_equal(e1, g["eta1_in"] * g["inside"])  # something was extrapolated


def test_p2m_golden_bit_exact():
    g = golden("p2m")
    dx = float(g["dx"])
    mesh = np.full_like(g["mesh_unb"], 7.0)
    ox.particles_to_mesh_2D_mp4(g["px"], g["py"], g["val"], mesh, dx, dx, periodic=False)
    assert np.array_equal(mesh, g["mesh_unb"])
    ox.particles_to_mesh_2D_mp4(g["pxw"], g["pyw"], g["val"], mesh, dx, dx, periodic=True)
    assert np.array_equal(mesh, g["mesh_per"])
    nr = g["w0"].shape[0]
    zp, rp, wp, w = g["Zl"].copy(), g["Rl"].copy(), 0 * g["Zl"], g["w0"].copy()
    ox.advect_vorticity_via_particles(zp, rp, wp, w, g["Zl"], g["Rl"], nr, g["uz"], g["ur"], dx, float(g["dt"]))
    assert np.array_equal(w, g["w_adv"])
    assert np.array_equal(wp, g["wp_after"])


def test_eno3_golden_from_reference_kernel_definitions():
    """a1-a7, a17: tests/golden/eno3.npz holds the outputs of the reference's OWN pystencils kernel
    definitions (pyst_kernels/advection_flux.py:9-332, advection_timestep.py:14-104, elementwise_ops.py:9-104,
    kernels/advect_vorticity_via_eno3.py:8-91, elasto_kernels/advect_refmap_via_eno3.py:8-115), imported
    unmodified and executed through tests/pystencils_shim.py (tests/golden/make_golden_eno3.py).  Both
    restatements (NumPy and C) must reproduce them; they follow the source's evaluation order, so bit for bit."""
    g = golden("eno3")
    assert bool(g["generated_with_shim"])
    f0, vel, flux0 = g["raw_field"], g["raw_vel"], g["raw_flux0"]
    rim = np.ones(f0.shape, bool)
    rim[2:-2, 2:-2] = False
    for tag, cons in (("cons", True), ("noncons", False)):
        for step in (ox.eno3_step_numpy, ox.eno3_step):
            f = f0.copy()
            step(f, vel[0].copy(), vel[1].copy(), float(g["step_dt_by_dx"]), cons)
            assert np.array_equal(f, g[f"step_field_{tag}"]), (tag, step.__name__)
            # the raw flux closure (a3 / a4): flux += ENO3(field; inv_dx) on [2:-2, 2:-2]; rim untouched
            f = f0.copy()
            step(f, vel[0].copy(), vel[1].copy(), -float(g["raw_inv_dx"]), cons)
            got = flux0.copy()
            got[2:-2, 2:-2] += (f - f0)[2:-2, 2:-2]
            assert_close(got, g[f"raw_flux_{tag}"], 1e-15, "raw flux " + tag)
        assert np.array_equal(g[f"raw_flux_{tag}"][rim], flux0[rim])
        # a5 / a6: the flux scratch holds the step's flux, its rim the fill value 0 (a1 ran on the whole array)
        assert np.all(g[f"step_flux_{tag}"][rim] == 0)
        assert np.array_equal(f0 + g[f"step_flux_{tag}"], g[f"step_field_{tag}"])
        assert np.array_equal(g[f"step_field_{tag}"][rim], f0[rim])
    # a1 / a2 elementwise closures: whole array (ghost layers 0), vector flavour = both components
    assert np.array_equal(g["ew_sum"], g["ew_a"] + g["ew_b"]) and np.array_equal(g["ew_alias"], g["ew_sum"])
    assert np.array_equal(g["ew_vsum"], g["ew_va"] + g["ew_vb"])
    assert np.all(g["ew_fill"] == -2.5) and np.all(g["ew_vfill"][0] == 1.25) and np.all(g["ew_vfill"][1] == -0.75)
    # a7 wrappers
    dt, dx = float(g["adv_dt"]), float(g["adv_dx"])
    for use_c in (False, True):
        w = g["adv_w0"].copy()
        ox.advect_vorticity_via_eno3(w, g["adv_uz0"].copy(), g["adv_ur0"].copy(), dt, dx, use_c=use_c)
        assert np.array_equal(w, g["adv_w_unb"])
        for _ in range(2):
            ox.advect_vorticity_via_eno3(w, g["adv_uz0"].copy(), g["adv_ur0"].copy(), dt, dx, use_c=use_c)
        assert np.array_equal(w, g["adv_w_unb_3steps"])
        w, uz, ur = g["adv_w0"].copy(), g["adv_uz0"].copy(), g["adv_ur0"].copy()
        ox.advect_vorticity_via_eno3(w, uz, ur, dt, dx, periodic_ghost=2, use_c=use_c)
        assert np.array_equal(w, g["adv_w_per"]) and np.array_equal(uz, g["adv_uz_per"])
        assert np.array_equal(ur, g["adv_ur_per"])
        # a17 wrappers
        e1, e2 = g["ref_e1_0"].copy(), g["ref_e2_0"].copy()
        ox.advect_refmap_via_eno3(e1, e2, g["adv_uz0"].copy(), g["adv_ur0"].copy(), dt, dx, use_c=use_c)
        assert np.array_equal(e1, g["ref_e1_unb"]) and np.array_equal(e2, g["ref_e2_unb"])
        e1, e2, uz, ur = g["ref_e1_0"].copy(), g["ref_e2_0"].copy(), g["adv_uz0"].copy(), g["adv_ur0"].copy()
        ox.advect_refmap_via_eno3_periodic(e1, e2, uz, ur, dt, dx, 2, float(g["ref_z_max"]), use_c=use_c)
        assert np.array_equal(e1, g["ref_e1_per"]) and np.array_equal(e2, g["ref_e2_per"])
        assert np.array_equal(uz, g["ref_uz_per"]) and np.array_equal(ur, g["ref_ur_per"])
    # physical rows advected by the mirrored-domain step: j in [0, Nr-3], k in [2, Nz-3]
    w0, w1 = g["adv_w0"], g["adv_w_unb"]
    assert np.array_equal(w1[-2:], w0[-2:]) and np.array_equal(w1[:, :2], w0[:, :2])
    assert np.array_equal(w1[:, -2:], w0[:, -2:]) and np.all(w1[:-2, 2:-2] != w0[:-2, 2:-2])


def test_pystencils_shim_rules():
    """the stand-in used to generate eno3.npz: ghost layers = max |offset| on every side of every axis
    (pystencils create_domain_kernel with ghost_layers=None), source evaluation order, shape check."""
    import pytest

    import pystencils_shim as ps

    @ps.kernel
    def _k():
        a, b = ps.fields("a, b : float64[2D]")
        a[0, 0] @= (b[0, 1] if b[0, 0] > -b[0, -2] else 2.0 * b[1, 0]) - a[0, 0]

    k = ps.create_kernel(_k, config=ps.CreateKernelConfig(cpu_openmp=False)).compile()
    assert k.ghost == 2
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((9, 11)), rng.standard_normal((9, 11))
    a0 = a.copy()
    k(a=a, b=b)
    exp = a0.copy()
    for j in range(2, 7):
        for i in range(2, 9):
            exp[j, i] = (b[j, i + 1] if b[j, i] > -b[j, i - 2] else 2.0 * b[j + 1, i]) - a0[j, i]
    assert np.array_equal(a, exp)

    @ps.kernel
    def _fixed():
        f = ps.fields("f : float64[4, 6]")
        f[0, 0] @= 1.5

    kf = ps.create_kernel(_fixed).compile()
    with pytest.raises(ValueError):
        kf(f=np.zeros((4, 7)))
    x = np.zeros((4, 6))
    kf(f=x)
    assert np.all(x == 1.5)


def test_eno3_numpy_vs_c_restatement():
    """the two independent restatements must agree on a second seeded case."""
    rng = np.random.default_rng(3)
    n0, n1 = 36, 52
    for cons in (True, False):
        f = rng.standard_normal((n0, n1))
        v0, v1 = rng.standard_normal((n0, n1)), rng.standard_normal((n0, n1))
        a, b = f.copy(), f.copy()
        ox.eno3_step_numpy(a, v0, v1, 0.13, cons)
        ox.eno3_step(b, v0, v1, 0.13, cons)
        assert np.array_equal(a[2:-2, 2:-2], b[2:-2, 2:-2])
        # pystencils ghost-layer rule: the rim of width 2 keeps its old values
        rim = np.ones((n0, n1), bool)
        rim[2:-2, 2:-2] = False
        assert np.array_equal(a[rim], f[rim]) and np.array_equal(b[rim], f[rim])
        assert not np.array_equal(a, f)


def test_eno3_advects_a_gaussian():
    """Physical sanity of axis / sign conventions: one step moves the blob
    by u*dt and the conservative form keeps total mass on a closed interior."""
    nr, nz = 64, 128
    dx = 1.0 / nz
    Z, R = _grid(nr, nz, dx)
    g = lambda z0: np.exp(-((Z - z0) ** 2 + (R - 0.25) ** 2) / 0.003)  # noqa: E731
    w = g(0.5)
    uz, ur = np.full_like(Z, 0.8), np.zeros_like(Z)
    dt = 0.1 * dx / 0.8
    before = w.copy()
    ox.advect_vorticity_via_eno3(w, uz, ur, dt, dx)
    moved = np.max(np.abs(w - g(0.5 + 0.8 * dt)))
    unmoved = np.max(np.abs(before - g(0.5 + 0.8 * dt)))
    assert moved < 0.05 * unmoved
    assert abs(w.sum() - before.sum()) < 1e-12 * before.sum()


def test_c_ports_of_the_numba_kernels_match_the_numpy_forms():
    """bench.py's CPU legs run C/OpenMP ports of the four numba kernels of the rigid-flow loop; they must
    agree with the NumPy restatements (which the reference-generated goldens pin) to rounding."""
    rng = np.random.default_rng(11)
    nr, nz = 37, 61
    dx = 1.0 / nz
    Z, R = _grid(nr, nz, dx)
    w0, psi = rng.standard_normal((nr, nz)), rng.standard_normal((nr, nz))
    a, ta, b, tb = w0.copy(), np.zeros_like(w0), w0.copy(), np.zeros_like(w0)
    ox.diffusion_RK2(a, ta, R, 2e-3, 0.2 * dx * dx / 2e-3, dx)
    ox.c_kernels.diffusion_RK2(b, tb, R, 2e-3, 0.2 * dx * dx / 2e-3, dx)
    assert_close(b, a, 1e-14, "diffusion")
    assert_close(tb, ta, 1e-14, "diffusion tmp")
    uz, ur, vz, vr = (np.zeros_like(w0) for _ in range(4))
    ox.compute_velocity_from_psi(uz, ur, psi, R, dx)
    ox.c_kernels.compute_velocity_from_psi(vz, vr, psi, R, dx)
    assert np.array_equal(uz, vz) and np.array_equal(ur, vr)
    chi = np.clip(rng.standard_normal((nr, nz)), 0, 1)
    pz, pr, qz, qr = (np.zeros_like(w0) for _ in range(4))
    ox.brinkmann_penalize(1e4, 3e-3, chi, 0.7, -0.2, uz, ur, pz, pr)
    ox.c_kernels.brinkmann_penalize(1e4, 3e-3, chi, 0.7, -0.2, uz, ur, qz, qr)
    assert np.array_equal(pz, qz) and np.array_equal(pr, qr)
    v1, v2 = rng.standard_normal((nr, nz)), None
    v2 = v1.copy()
    ox.compute_vorticity_from_velocity(v1, uz, ur, dx)
    ox.c_kernels.compute_vorticity_from_velocity(v2, uz, ur, dx)
    assert np.array_equal(v1, v2)
