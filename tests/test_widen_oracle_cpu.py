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
