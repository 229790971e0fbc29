"""Pins the oracle restatements of the SURVEY.md 8f rows against the unmodified reference
(fixtures written by tests/golden/make_golden_widen.py).  CPU only."""
import numpy as np

from conftest import assert_close, golden
from oracle import axisym_oracle as ox

ULP = 1e-14


def _grid(nr, nz, dx):
    z = np.linspace(dx / 2, nz * dx - dx / 2, nz)
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    return np.meshgrid(z, r)


def test_velocity_from_phi_golden():
    g = golden("velocity_from_phi")
    uz, ur = np.zeros_like(g["phi"]), np.zeros_like(g["phi"])
    ox.compute_velocity_from_phi(uz, ur, g["phi"], float(g["dx"]))
    assert np.array_equal(uz, g["uz"]) and np.array_equal(ur, g["ur"])


def test_baroclinic_golden():
    g = golden("baroclinic")
    dx, dt, nu = float(g["dx"]), float(g["dt"]), float(g["nu"])
    Z, R = _grid(*g["w0"].shape, dx)
    args = (g["u_z"], g["u_r"], g["o_z"], g["o_r"], g["rho"], dt, dx)
    w = g["w0"].copy()
    ox.update_baroclinic_vorticity(w, *args)
    assert_close(w, g["w_plain"], ULP, "baroclinic")
    w = g["w0"].copy()
    ox.update_baroclinic_vorticity(w, *args, penal_term_z=g["p_z"], penal_term_r=g["p_r"])
    assert_close(w, g["w_penal"], ULP, "baroclinic penal")
    w = g["w0"].copy()
    ox.update_baroclinic_vorticity(w, *args, penal_term_z=g["p_z"], penal_term_r=g["p_r"], R=R, nu=nu)
    assert_close(w, g["w_diff_penal"], ULP, "baroclinic diff penal")
    # the source must actually matter in the fixture (guards against a vacuous comparison)
    assert np.max(np.abs(g["w_diff_penal"] - g["w0"])) > 1e-6 * np.max(np.abs(g["w0"]))


# ---------------------------------------------------------------------------------------------
# 8f-3 narrow-band re-initialisation: third party (scikit-fmm), PARITY UNPINNED.  What can be pinned on
# the CPU: the restated marcher behaves like a distance solver, and the fixed-point iteration the GPU
# runs (tests/reinit_model.py) reproduces the marcher bit for bit wherever the field is smooth.
# ---------------------------------------------------------------------------------------------
def _sphere(nr, nz, zc, rc, rad):
    dx = 1.0 / nz
    z = np.linspace(dx / 2, 1 - dx / 2, nz)
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    Z, R = np.meshgrid(z, r)
    return dx, Z, R, rad - np.sqrt((Z - zc) ** 2 + (R - rc) ** 2)


def test_marcher_restatement_is_a_distance_solver():
    import pytest

    dx, Z, R, true = _sphere(40, 128, 0.47, 0.0, 0.15)
    band = 6 * dx
    for order in (1, 2):
        d = ox.fmm_distance(true * (1 + 0.3 * np.sin(9 * Z + 5 * R)), dx, narrow=band, order=order)
        seen = ~d.mask
        assert seen.sum() > 500 and d.mask[0, 0] and d.mask[-1, -1]
        assert np.all(np.sign(d.data[seen]) == np.sign(true[seen]))
        acc = seen & (np.abs(d.data) <= band)
        # the marcher's own accuracy: front cells are first-order (up to ~0.3 dx), it does not grow in the band
        assert np.max(np.abs(d.data - true)[acc]) <= (0.35 if order == 2 else 0.6) * dx
        # accepted values never exceed the band, the tentative ring lies just outside it
        ring = seen & ~acc
        assert ring.any() and np.all(np.abs(d.data[ring]) > band) and np.all(np.abs(d.data[ring]) < band + 1.5 * dx)
    with pytest.raises(ValueError):
        ox.fmm_distance(np.ones((8, 8)), 0.1, narrow=0.3)


def test_gpu_iteration_model_reproduces_the_marcher():
    import reinit_model as rm

    rng = np.random.default_rng(11)
    for case in range(6):
        nz, nr = int(rng.integers(40, 90)), int(rng.integers(20, 50))
        dx, Z, R, true = _sphere(nr, nz, rng.uniform(0.3, 0.7), rng.uniform(0, 0.1), rng.uniform(0.08, 0.25))
        phi = true * (1 + rng.uniform(0, 0.5) * np.sin(rng.uniform(2, 12) * Z + rng.uniform(2, 12) * R))
        for order in (1, 2):
            band = rng.uniform(2, 8) * dx
            ref = ox.fmm_distance(phi, dx, narrow=band, order=order)
            out, seen, sweeps = rm.reinit(phi, dx, band, order)
            assert np.array_equal(seen, ~ref.mask), (case, order)
            assert np.array_equal(out[seen], ref.data[seen]), (case, order)
            assert sweeps <= 8
    # the driver's situation (soft_sphere_streaming.py:190-199): old distances outside, band pinned from the map
    dx, Z, R, phi = _sphere(32, 64, 0.5, 0.0, 0.15)
    for step in range(3):
        pinned = 0.15 - np.sqrt(((Z - 0.5 - 0.3 * dx * (step + 1)) * 1.05) ** 2 + (R / 1.05) ** 2)
        phi = np.where(phi > -3 * dx, pinned, phi)
        ref = ox.fmm_distance(phi, dx, narrow=6 * dx)
        out, seen, _ = rm.reinit(phi, dx, 6 * dx, 2)
        assert np.array_equal(seen, ~ref.mask) and np.array_equal(out[seen], ref.data[seen])
        phi = np.where(ref.mask, phi, ref.data)


def test_gpu_iteration_model_terminates_where_fronts_collide():
    import reinit_model as rm

    dx, Z, R, a = _sphere(18, 36, 0.3, 0.05, 0.1)
    b = 0.12 - np.sqrt((Z - 0.62) ** 2 + (R - 0.1) ** 2)
    phi = np.maximum(a, b) * (1 + 0.3 * np.sin(5 * Z + 7 * R))
    ref = ox.fmm_distance(phi, dx, narrow=5.7 * dx)
    out, seen, sweeps = rm.reinit(phi, dx, 5.7 * dx, 2)
    assert np.array_equal(seen, ~ref.mask)
    assert np.max(np.abs(out[seen] - ref.data[seen])) <= 5e-3 * dx     # only the shock cells differ
