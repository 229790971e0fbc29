"""Pins the oracle's restatement of the C++ core family (core/src/instantiate.yml) against outputs of the
UNMODIFIED reference C++ compiled as-is (tests/golden/core_family.npz, made by make_golden_core.py).  CPU only."""
import numpy as np

from conftest import assert_close, golden
from oracle import axisym_oracle as ox

KERNELS = ("linear_kernel", "mp4", "mp6", "yang_smooth_three_point_kernel")


def test_mesh_to_particles_and_particles_to_mesh_family():
    g = golden("core_family")
    dx, dy = float(g["dx"]), float(g["dy"])
    for k in KERNELS:
        for per in (True, False):
            mid = "" if per else "unbounded_"
            px, py = (g["pxw"], g["pyw"]) if per else (g["px"], g["py"])
            ox_, oy_ = np.zeros_like(px), np.zeros_like(px)
            ox.mesh_to_particles_2D(k, per, g["fx"], g["fy"], px, py, ox_, oy_, dx, dy)
            # same operation order as the reference => bit-identical (Yang's kernel: same libm here)
            assert np.array_equal(ox_, g[f"m2p_{mid}{k}_x"]), (k, per)
            assert np.array_equal(oy_, g[f"m2p_{mid}{k}_y"]), (k, per)
            mesh = np.ones_like(g["fx"])
            ox.particles_to_mesh_2D(k, per, px, py, g["val"], mesh, dx, dy)
            assert np.array_equal(mesh, g[f"p2m_{mid}{k}"]), (k, per)
    o1 = np.zeros_like(g["q1"])
    ox.mesh_to_particles_1D_mp4(g["f1"], g["q1"], o1, dx)
    assert np.array_equal(o1, g["m2p_1d"])
    m1 = np.ones_like(g["f1"])
    ox.particles_to_mesh_1D_mp4(g["q1"], g["v1"], m1, dx)
    assert np.array_equal(m1, g["p2m_1d"])


def test_wrap_particles():
    g = golden("core_family")
    wx, wy = g["wrap_x0"].copy(), g["wrap_y0"].copy()
    ox.wrap_particles_around_2D_domain(wx, wy, 0.0, 1.0, 0.0, 0.5)
    assert np.array_equal(wx, g["wrap_x"]) and np.array_equal(wy, g["wrap_y"])
    # only the first / last 10 entries (x) or rows (y) are ever touched
    assert np.array_equal(wx[:, 10:-10], g["wrap_x0"][:, 10:-10]) and np.array_equal(wy[10:-10], g["wrap_y0"][10:-10])
    w1 = g["wrap1_in"].copy()
    ox.wrap_particles_around_1D_domain(w1, 0.0, 1.0)
    assert np.array_equal(w1, g["wrap1_out"])
    sx, sy = g["wrap_small_in"].copy(), g["wrap_small_in"].copy()
    ox.wrap_particles_around_2D_domain(sx, sy, 0.0, 1.0, 0.0, 1.0)
    assert np.array_equal(sx, g["wrap_small_x"]) and np.array_equal(sy, g["wrap_small_y"])


def test_least_squares_extrapolation_orders():
    g = golden("core_family")
    for order in (1, 2):
        c, a, b = g["ls_cur"].copy(), g["ls_ex"].copy(), g["ls_ey"].copy()
        ox.extrapolate_using_least_squares(order, c, g["ls_tgt"], a, b, g["ls_gx"], g["ls_gy"])
        assert np.array_equal(c, g[f"ls{order}_cur"])
        assert np.array_equal(a, g[f"ls{order}_ex"]) and np.array_equal(b, g[f"ls{order}_ey"]), order
    # the quadratic fit reproduces a quadratic field exactly where the patch holds >= 6 known cells in general
    # position; the linear one does not
    X, Y = g["ls_gx"][None, :], g["ls_gy"][:, None]
    quad = 1.5 * X - 0.7 * Y + 0.4 * X * X - 0.3 * X * Y
    new = (g["ls2_cur"] == 1) & (g["ls_cur"] == 0)
    assert new.sum() > 300
    c, a, b = g["ls_cur"].copy(), np.where(g["ls_cur"] == 1, quad, 0.0), np.where(g["ls_cur"] == 1, quad, 0.0)
    ox.extrapolate_using_least_squares(2, c, g["ls_tgt"], a, b, g["ls_gx"], g["ls_gy"])
    good = new & np.isfinite(a)
    err = np.abs(a - quad)[good]
    assert np.median(err) < 1e-9
