"""GPU parity of the SURVEY.md 8f rows (kernels of csrc/widen.cu) against the oracle and the
reference-generated fixtures.  Bar: <= 1e-10 relative L-infinity; these kernels repeat the reference's
operation order without FMA contraction, so bit equality is asserted where it holds."""
import numpy as np
import pytest

from conftest import assert_close, golden
from oracle import axisym_oracle as ox

pytestmark = pytest.mark.gpu

SHAPES = [(24, 56), (37, 53), (64, 128), (130, 70), (3, 3), (5, 4)]


def _grid(nr, nz, dx=None):
    dx = 1.0 / nz if dx is None else dx
    z = np.linspace(dx / 2, nz * dx - dx / 2, nz)
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    Z, R = np.meshgrid(z, r)
    return dx, z, r, Z, R


def _rand(rng, nr, nz, amp=1.0):
    dx, z, r, Z, R = _grid(nr, nz)
    return amp * (np.sin(2 * np.pi * (Z + 0.3 * R)) * np.exp(-((Z - 0.5) ** 2 + R ** 2) / 0.05)
                  + 0.1 * rng.standard_normal((nr, nz)))


@pytest.fixture(scope="module")
def K():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyaxisymflow_b200.ops as ops

    return ops


def test_velocity_from_phi(K):
    g = golden("velocity_from_phi")
    uz, ur = np.zeros_like(g["phi"]), np.zeros_like(g["phi"])
    K.compute_velocity_from_phi_unb(uz, ur, g["phi"], float(g["dx"]))
    assert np.array_equal(uz, g["uz"]) and np.array_equal(ur, g["ur"])
    rng = np.random.default_rng(3)
    for nr, nz in SHAPES:
        dx = 1.0 / nz
        phi = _rand(rng, nr, nz)
        a, b, c, d = (np.full((nr, nz), 7.0) for _ in range(4))
        K.compute_velocity_from_phi_unb(a, b, phi, dx)
        ox.compute_velocity_from_phi(c, d, phi, dx)
        assert np.array_equal(a, c) and np.array_equal(b, d), (nr, nz)
    # strided views (row pitch != nz), as the periodic drivers pass them
    big = _rand(rng, 40, 100)
    a, b, c, d = (np.zeros((40, 100)) for _ in range(4))
    K.compute_velocity_from_phi_unb(a[:, 2:-3], b[:, 2:-3], big[:, 2:-3], 0.01)
    ox.compute_velocity_from_phi(c[:, 2:-3], d[:, 2:-3], big[:, 2:-3], 0.01)
    assert np.array_equal(a, c) and np.array_equal(b, d)


def test_velocity_from_phi_device_fields(K):
    """zero-copy path: DeviceField in, DeviceField out."""
    import torch
    from pyaxisymflow_b200.device import DeviceField

    rng = np.random.default_rng(4)
    nr, nz = 96, 200
    phi = _rand(rng, nr, nz)
    uz, ur = (DeviceField(torch.zeros(nr, nz, dtype=torch.float64, device="cuda")) for _ in range(2))
    K.compute_velocity_from_phi_unb(uz, ur, DeviceField(torch.from_numpy(phi).cuda()), 1.0 / nz)
    c, d = np.zeros_like(phi), np.zeros_like(phi)
    ox.compute_velocity_from_phi(c, d, phi, 1.0 / nz)
    assert np.array_equal(uz.t.cpu().numpy(), c) and np.array_equal(ur.t.cpu().numpy(), d)


def test_baroclinic(K):
    g = golden("baroclinic")
    dx, dt, nu = float(g["dx"]), float(g["dt"]), float(g["nu"])
    _, _, _, Z, R = _grid(*g["w0"].shape, dx)
    base = (g["u_z"], g["u_r"], g["o_z"], g["o_r"], g["rho"])
    w = g["w0"].copy()
    K.update_baroclinic_vorticity(w, *base, dt, dx)
    assert_close(w, g["w_plain"], 1e-13, "baroclinic")
    w = g["w0"].copy()
    K.update_baroclinic_vorticity_penal(w, *base, g["p_z"], g["p_r"], dt, dx)
    assert_close(w, g["w_penal"], 1e-13, "baroclinic penal")
    w = g["w0"].copy()
    K.update_baroclinic_vorticity_diff_penal(w, *base, g["p_z"], g["p_r"], R, nu, dt, dx)
    assert_close(w, g["w_diff_penal"], 1e-13, "baroclinic diff penal")
    rng = np.random.default_rng(5)
    for nr, nz in SHAPES:
        dx, _, _, Z, R = _grid(nr, nz)
        uz, ur = _rand(rng, nr, nz), _rand(rng, nr, nz)
        oz, orr = uz + 1e-3 * _rand(rng, nr, nz), ur + 1e-3 * _rand(rng, nr, nz)
        rho = 1.0 + 0.5 * np.clip(_rand(rng, nr, nz) + 0.5, 0, 1)
        pz, pr = _rand(rng, nr, nz, 5.0), _rand(rng, nr, nz, 5.0)
        w0 = _rand(rng, nr, nz, 3.0)
        for mode in range(3):
            a, b = w0.copy(), w0.copy()
            if mode == 0:
                K.update_baroclinic_vorticity(a, uz, ur, oz, orr, rho, 2e-3, dx)
                ox.update_baroclinic_vorticity(b, uz, ur, oz, orr, rho, 2e-3, dx)
            elif mode == 1:
                K.update_baroclinic_vorticity_penal(a, uz, ur, oz, orr, rho, pz, pr, 2e-3, dx)
                ox.update_baroclinic_vorticity(b, uz, ur, oz, orr, rho, 2e-3, dx, penal_term_z=pz, penal_term_r=pr)
            else:
                K.update_baroclinic_vorticity_diff_penal(a, uz, ur, oz, orr, rho, pz, pr, R, 1e-2, 2e-3, dx)
                ox.update_baroclinic_vorticity(b, uz, ur, oz, orr, rho, 2e-3, dx, penal_term_z=pz, penal_term_r=pr,
                                               R=R, nu=1e-2)
            assert np.array_equal(a, b), (nr, nz, mode)
            # the rim is not touched
            assert np.array_equal(a[0], w0[0]) and np.array_equal(a[:, -1], w0[:, -1])
