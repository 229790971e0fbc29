"""GPU parity of the SURVEY.md 8f rows (kernels of csrc/phi_baroclinic.cu) against the oracle and the
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


@pytest.mark.parametrize("path", ["march", "tiled"])
def test_velocity_from_phi(K, path):
    from pyaxisymflow_b200 import _lib

    _lib.call("axb_set_stencil_path", 1 if path == "tiled" else 0)
    try:
        _velocity_from_phi_cases(K)
    finally:
        _lib.call("axb_set_stencil_path", 0)


def _velocity_from_phi_cases(K):
    g = golden("velocity_from_phi")
    uz, ur = np.zeros_like(g["phi"]), np.zeros_like(g["phi"])
    K.compute_velocity_from_phi_unb(uz, ur, g["phi"], float(g["dx"]))
    assert np.array_equal(uz, g["uz"]) and np.array_equal(ur, g["ur"])
    rng = np.random.default_rng(3)
    # the larger shapes reach the row-marching interior kernel (>= 258 columns, >= 18 rows, even pitch)
    for nr, nz in SHAPES + [(70, 600), (130, 1030), (40, 777), (18, 258), (35, 2048)]:
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


@pytest.mark.parametrize("path", ["march", "tiled"])
def test_baroclinic(K, path):
    """tiled path: the reference's divisions bit for bit; marching path (interior blocks of grids with >= 258 columns):
    reciprocal multiplications, held to 1e-12 (the bar is 1e-10)"""
    from pyaxisymflow_b200 import _lib

    _lib.call("axb_set_stencil_path", 1 if path == "tiled" else 0)
    try:
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
        for nr, nz in SHAPES + [(70, 600), (40, 777), (18, 258), (35, 1030)]:
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
                if path == "tiled" or nz < 258:
                    assert np.array_equal(a, b), (nr, nz, mode)
                else:
                    assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b - w0)), (nr, nz, mode)
                    assert np.array_equal(a[:, :2], b[:, :2]) and np.array_equal(a[-1], b[-1])   # edge blocks: exact form
                # the rim is not touched
                assert np.array_equal(a[0], w0[0]) and np.array_equal(a[:, -1], w0[:, -1])
    finally:
        _lib.call("axb_set_stencil_path", 0)


# ---------------------------------------------------------------------------------------------
# 8f-3 narrow-band re-initialisation (csrc/reinit.cu) against the restated marcher
# (oracle.fmm_distance; scikit-fmm is third party and not vendored: parity unpinned).
# ---------------------------------------------------------------------------------------------
def _sphere(nr, nz, zc, rc, rad):
    dx = 1.0 / nz
    z = np.linspace(dx / 2, 1 - dx / 2, nz)
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    Z, R = np.meshgrid(z, r)
    return dx, Z, R, rad - np.sqrt((Z - zc) ** 2 + (R - rc) ** 2)


def _check_against_marcher(got, ref, phi, what):
    assert isinstance(got, np.ma.MaskedArray)
    assert np.array_equal(np.ma.getmaskarray(got), np.ma.getmaskarray(ref)), what + ": masks differ"
    seen = ~np.ma.getmaskarray(ref)
    scale = np.max(np.abs(ref.data[seen]))
    err = np.max(np.abs(got.data[seen] - ref.data[seen])) / scale
    assert err <= 1e-12, f"{what}: relative Linf {err:.3e}"
    return int(np.count_nonzero(got.data[seen] != ref.data[seen]))


def test_reinit_matches_marcher(K):
    from pyaxisymflow_b200 import reinit

    rng = np.random.default_rng(21)
    inexact = 0
    for case, (nr, nz) in enumerate([(40, 91), (57, 73), (31, 44), (64, 128), (45, 200), (130, 70)]):
        dx, Z, R, true = _sphere(nr, nz, rng.uniform(0.3, 0.7), rng.uniform(0, 0.1), rng.uniform(0.08, 0.25))
        phi = true * (1 + rng.uniform(0, 0.5) * np.sin(rng.uniform(2, 12) * Z + rng.uniform(2, 12) * R))
        for order in (1, 2):
            band = rng.uniform(2, 8) * dx
            ref = ox.fmm_distance(phi, dx, narrow=band, order=order)
            keep = phi.copy()
            got = reinit.distance(phi, dx=dx, narrow=band, order=order)
            assert np.array_equal(phi, keep)                 # like skfmm.distance: the input is not modified
            inexact += _check_against_marcher(got, ref, phi, f"case {case} order {order}")
    assert inexact == 0, f"{inexact} cells differ from the marcher in the last bits"
    # a band thinner than a cell (front flags decide what is usable) and a strided view
    dx, Z, R, true = _sphere(48, 100, 0.523, 0.033, 0.2)
    phi = true * (1 + 0.2 * np.sin(6 * Z))
    _check_against_marcher(reinit.distance(phi, dx=dx, narrow=0.6 * dx), ox.fmm_distance(phi, dx, narrow=0.6 * dx), phi, "thin")
    view = phi[:, 3:-5]
    _check_against_marcher(reinit.distance(view, dx=dx, narrow=4 * dx), ox.fmm_distance(view, dx, narrow=4 * dx), view, "view")


def test_reinit_errors_and_device_path(K):
    import torch
    from pyaxisymflow_b200 import reinit
    from pyaxisymflow_b200.device import DeviceField

    with pytest.raises(ValueError):
        reinit.distance(np.ones((16, 16)), dx=0.1, narrow=0.3)          # no zero contour (skfmm: ValueError)
    with pytest.raises(ValueError):
        reinit.distance(np.ones((16, 16)), dx=0.1, narrow=0.0)
    dx, Z, R, true = _sphere(64, 160, 0.45, 0.0, 0.2)
    phi = true * (1 + 0.3 * np.sin(8 * Z + 3 * R))
    ref = ox.fmm_distance(phi, dx, narrow=5 * dx)
    d, mask = reinit.distance(DeviceField(torch.from_numpy(phi).cuda()), dx=dx, narrow=5 * dx)
    assert np.array_equal(mask.cpu().numpy(), np.ma.getmaskarray(ref))
    seen = ~np.ma.getmaskarray(ref)
    assert np.max(np.abs(d.t.cpu().numpy()[seen] - ref.data[seen])) <= 1e-12 * 5 * dx
    # in-place form used by the stepper: unreached cells keep their old value
    t = torch.from_numpy(phi).cuda()
    r = reinit.NarrowBandReinit(64, 160)
    r(t, dx, 5 * dx)
    out = t.cpu().numpy()
    assert np.array_equal(out[~seen], phi[~seen]) and np.max(np.abs(out[seen] - ref.data[seen])) <= 1e-12 * 5 * dx
    assert 2 <= r.sweeps <= 12
    with pytest.raises(ValueError):
        r(torch.zeros(8, 8, dtype=torch.float64, device="cuda"), dx, 5 * dx)


def test_reinit_terminates_on_rough_and_colliding_input(K):
    """the marcher is order dependent where fronts collide or the level set is noisy; the GPU iteration must
    still terminate with a defined result close to it (and identical front cells)."""
    from pyaxisymflow_b200 import reinit

    dx, Z, R, a = _sphere(18, 36, 0.3, 0.05, 0.1)
    b = 0.12 - np.sqrt((Z - 0.62) ** 2 + (R - 0.1) ** 2)
    phi = np.maximum(a, b) * (1 + 0.3 * np.sin(5 * Z + 7 * R))
    ref = ox.fmm_distance(phi, dx, narrow=5.7 * dx)
    got = reinit.distance(phi, dx=dx, narrow=5.7 * dx)
    assert np.array_equal(np.ma.getmaskarray(got), np.ma.getmaskarray(ref))
    seen = ~np.ma.getmaskarray(ref)
    assert np.max(np.abs(got.data[seen] - ref.data[seen])) <= 5e-3 * dx
    rng = np.random.default_rng(5)
    dx, Z, R, true = _sphere(30, 48, 0.5, 0.04, 0.2)
    noisy = true + 0.2 * dx * rng.standard_normal(true.shape)
    got = reinit.distance(noisy, dx=dx, narrow=5 * dx)
    front_d, front = ox.fmm_initial_front(noisy, dx)
    assert np.array_equal(got.data[front], front_d[front]) and not np.ma.getmaskarray(got)[front].any()
    seen = ~np.ma.getmaskarray(got)
    assert np.all(np.sign(got.data[seen]) == np.sign(noisy[seen]))


def test_reinit_full_size_against_marcher(K):
    """config-C3 size (2048 x 8192): same cells, same values as the marcher; signed-distance sanity."""
    import torch
    from pyaxisymflow_b200 import reinit

    nr, nz = 2048, 8192
    dx, Z, R, true = _sphere(nr, nz, 0.47, 0.0, 0.15)
    phi = true * (1 + 0.2 * np.sin(9 * Z + 5 * R))
    band = 6 * dx
    ref = ox.fmm_distance(phi, dx, narrow=band)
    t = torch.from_numpy(phi).cuda()
    mask = torch.empty((nr, nz), dtype=torch.uint8, device="cuda")
    r = reinit.NarrowBandReinit(nr, nz)
    r(t, dx, band, mask_out=mask)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t2 = torch.from_numpy(phi).cuda()
    ev0.record()
    r(t2, dx, band)
    ev1.record()
    torch.cuda.synchronize()
    print(f"\nreinit 2048x8192, band 6 dx: {ev0.elapsed_time(ev1):.3f} ms, {r.sweeps} sweep launches")
    got = t.cpu().numpy()
    m = mask.cpu().numpy().astype(bool)
    assert np.array_equal(m, np.ma.getmaskarray(ref))
    seen = ~m
    assert np.array_equal(got[m], phi[m])
    assert np.max(np.abs(got[seen] - ref.data[seen])) <= 1e-12 * band
    acc = seen & (np.abs(got) <= band)
    assert np.max(np.abs(got - true)[acc]) <= 0.35 * dx          # the marcher's own accuracy (first-order front)
    assert np.array_equal(t2.cpu().numpy(), got)                 # deterministic


def test_soft_sphere_stepper_with_reinit(K):
    """config-C3 loop INCLUDING the level-set re-initialisation (soft_sphere_streaming.py:195-199)."""
    from test_cuda_parity import _oracle_soft_sphere_loop
    from pyaxisymflow_b200.timestep import SoftSphereStepper

    nz, steps = 64, 6
    w, eta1, eta2, phi, avg_psi, t = _oracle_soft_sphere_loop(nz, steps, reinit=True, Z_cm=0.47)
    s = SoftSphereStepper(nz, Z_cm=0.47, reinit_levelset=True)
    s.step(steps)
    assert abs(s.t - t) <= 1e-12 * t and s.ls_sweeps >= 4
    assert_close(s.ball_phi.cpu().numpy(), phi, 1e-10, "ball_phi (re-initialised)")
    assert_close(s.eta1.cpu().numpy(), eta1, 1e-10, "eta1")
    assert_close(s.eta2.cpu().numpy(), eta2, 1e-10, "eta2")
    assert_close(s.vorticity.cpu().numpy(), w, 1e-9, "vorticity (soft sphere, reinit)")
    assert_close(s.avg_psi.cpu().numpy(), avg_psi, 1e-9, "avg_psi")


# ---------------------------------------------------------------------------------------------
# 8f-4 output either side of the loop: .vti dumps and restart files taken from device-resident fields
# ---------------------------------------------------------------------------------------------
def test_vti_and_npz_from_device_fields(K, tmp_path):
    import os
    import torch
    from pyaxisymflow_b200 import io
    from pyaxisymflow_b200.device import DeviceField
    from pyaxisymflow_b200.utils.dump_vtk import read_vti, vtk_init, vtk_write

    nr, nz = 96, 256
    rng = np.random.default_rng(8)
    a, b = rng.standard_normal((nr, nz)), rng.standard_normal((nr, nz))
    da, db = DeviceField(torch.from_numpy(a).cuda()), torch.from_numpy(b).cuda()
    image, tmp, writer = vtk_init(nz, nr)
    writer.asynchronous = True
    path = os.path.join(tmp_path, "dev.vti")
    vtk_write(path, image, tmp, writer, ["avg_psi", "u_z"], [da, db], nz, nr)
    da.t.mul_(0.0)                      # the loop goes on: `avg_psi[...] *= 0.0` right after the dump
    db.add_(1.0)
    writer.wait()
    dims, name, fields = read_vti(path)
    assert dims == (nz, nr) and np.array_equal(fields["avg_psi"], a) and np.array_equal(fields["u_z"], b)
    rpath = os.path.join(tmp_path, "restart.npz")
    io.save_npz(rpath, t=0.5, vorticity=db, part_phi=a)
    with np.load(rpath) as f:
        assert float(f["t"]) == 0.5 and np.array_equal(f["vorticity"], b + 1.0) and np.array_equal(f["part_phi"], a)


@pytest.mark.parametrize("kind", ["rigid", "soft", "particle", "soft-device", "particle-device"])
def test_restart_resumes_bit_identically(K, tmp_path, kind):
    """run 3 steps, write restart.npz, run 3 more; a fresh stepper loaded from the file must land on the same bits
    (particle_in_bubble_oscillatory_flow.py:129-147 / :236-257)"""
    import os
    from pyaxisymflow_b200 import io
    from pyaxisymflow_b200.timestep import ParticleFlowStepper, RigidFlowStepper, SoftSphereStepper

    def make():
        if kind == "rigid":
            s = RigidFlowStepper(64)
            s.seed_vorticity()
            return s
        if kind == "soft":
            return SoftSphereStepper(64, Z_cm=0.47, reinit_levelset=True)
        if kind == "soft-device":
            return SoftSphereStepper(64, Z_cm=0.47, device_scalars=True, use_graph=True)
        if kind == "particle-device":
            return ParticleFlowStepper(64, device_scalars=True, use_graph=True)
        return ParticleFlowStepper(64)

    path = os.path.join(tmp_path, "restart.npz")
    s = make()
    s.step(3)
    io.save_restart(s, path, asynchronous=True)
    s.step(3)                           # keeps running while the file is written
    io.wait()
    r = make()
    io.load_restart(r, path)
    r.step(3)
    if kind.endswith("-device"):
        # device-resident loop scalars: save_restart reads them through sync_scalars, load_restart pushes them back
        s.sync_scalars()
        r.sync_scalars()
    if kind.startswith("particle"):
        # the penalisation force is an atomic sum of terms that cancel to ~1e-9 of their size (lambda = 1e12), so two
        # runs of the SAME loop already differ in the rigid-body feedback at ~1e-7: the resumed
        # run is held to that, not to bit equality
        assert_close(r.vorticity.cpu().numpy(), s.vorticity.cpu().numpy(), 1e-5, "resumed vorticity (particle)")
        assert abs(r.t - s.t) <= 1e-12 and r.it == s.it
        assert abs(r.part_Z_cm - s.part_Z_cm) <= 1e-9 and abs(r.U_z_cm_part - s.U_z_cm_part) <= 1e-5 * max(1.0, abs(s.U_z_cm_part))
        return
    assert np.array_equal(r.vorticity.cpu().numpy(), s.vorticity.cpu().numpy())
    if kind == "rigid":
        assert np.array_equal(r.state.cpu().numpy(), s.state.cpu().numpy())
    else:
        assert r.t == s.t and r.it == s.it and r.it == 6
        assert np.array_equal(r.ball_phi.cpu().numpy(), s.ball_phi.cpu().numpy())
        assert np.array_equal(r.eta1.cpu().numpy(), s.eta1.cpu().numpy())


def test_particle_ensemble_interleaved_members(K):
    """config C5: members interleaved on their own streams (one host read per member and step) do exactly what
    a member stepping alone does -- each is compared with the reference loop driven by its own trace."""
    from test_cuda_parity import _oracle_particle_loop
    from pyaxisymflow_b200.timestep import ParticleEnsemble, ParticleFlowStepper

    nz, steps = 80, 6
    params = [(8.0, 0.01), (16.0, 0.02), (12.0, 0.005)]
    first = ParticleFlowStepper(nz, freq=params[0][0], e=params[0][1])
    members = [first] + [ParticleFlowStepper(nz, freq=f, e=e, solver=first.solver) for f, e in params[1:]]
    ens = ParticleEnsemble(members)
    assert members[1].solver is not first.solver and members[1].solver.factors is first.solver.factors
    ens.step(2)
    ens.step(steps - 2)
    for m, (f, e) in zip(members, params):
        assert len(m.trace) == steps and m.it == steps
        w, avg_vort, t, pz, U, forces = _oracle_particle_loop(nz, steps, freq=f, e=e, trace=m.trace)
        assert abs(m.t - t) <= 1e-12 * t
        assert_close(m.vorticity.cpu().numpy(), w, 1e-9, f"ensemble member f={f}")
        assert_close(m.avg_vort.cpu().numpy(), avg_vort, 1e-9, f"ensemble member avg_vort f={f}")
        for got, want in zip(m.trace, forces):
            assert abs(got[4] - want) <= 1e-5 * max(abs(want), 1e-12)


def test_heaviside_shortcuts_are_bit_identical(K):
    """smooth_Heaviside only evaluates the blend expression inside the band, and the analytic-sphere form skips the
    square root for cells safely inside / outside the blend shell when phi is not requested: same bits as the
    reference expression evaluated everywhere (kernels/smooth_Heaviside.py:10-14)."""
    rng = np.random.default_rng(12)
    for nr, nz, zc, rc, rad, wf in [(96, 300, 0.47, 0.0, 0.15, 2.0), (64, 128, 0.3, 0.05, 0.02, 2 ** 0.5),
                                    (50, 200, 0.6, 0.1, 0.01, 3.0), (33, 77, 0.5, 0.0, 0.4, 1.0)]:
        dx, z, r, Z, R = _grid(nr, nz)
        w = wf * dx
        phi = -np.sqrt((Z - zc) ** 2 + (R - rc) ** 2) + rad
        Href = np.zeros_like(phi)
        ox.smooth_Heaviside(Href, phi, w)
        H1, H2, ps = np.zeros_like(phi), np.zeros_like(phi), np.zeros_like(phi)
        K.smooth_Heaviside_sphere(H1, Z, R, zc, rc, rad, w, phi_out=ps)      # exact path
        K.smooth_Heaviside_sphere(H2, Z, R, zc, rc, rad, w)                  # with the shortcuts
        assert np.array_equal(H1, H2), (nr, nz)
        assert np.max(np.abs(H1 - Href)) <= 1e-12
        H3 = np.full_like(phi, 7.0)
        K.smooth_Heaviside(H3, phi, w)
        assert np.max(np.abs(H3 - Href)) <= 1e-15 and np.array_equal(H3[np.abs(phi) >= w], Href[np.abs(phi) >= w])
    # a rough level set with values exactly on the band edges
    phi = rng.standard_normal((40, 90)) * 0.05
    phi[3, 4], phi[5, 6] = 0.02, -0.02
    Href, H = np.zeros_like(phi), np.zeros_like(phi)
    ox.smooth_Heaviside(Href, phi, 0.02)
    K.smooth_Heaviside(H, phi, 0.02)
    assert np.max(np.abs(H - Href)) <= 1e-15 and H[3, 4] == 1.0 and H[5, 6] == 0.0


@pytest.mark.parametrize("nr,nz", [(70, 600), (130, 1030), (18, 258), (40, 777), (35, 2048)])
def test_solid_stress_marching_equals_tiled(K, nr, nz):
    """G-SOL-1/2 (a18, a19): the opt-in row-marching interior kernels (axb_set_solid_march; grids with >= 258 columns
    and >= 18 rows) produce the bits of the default 2-D tiled kernels, which repeat the reference's operation order;
    both are checked against the oracle."""
    from pyaxisymflow_b200 import _lib

    rng = np.random.default_rng(nr * 7 + nz)
    dx, _, _, Z, R = _grid(nr, nz)
    eta1, eta2 = Z + 0.05 * _rand(rng, nr, nz), R + 0.05 * _rand(rng, nr, nz)
    chi = np.clip(_rand(rng, nr, nz) + 0.5, 0, 1)
    names = ("s11", "s12", "s22", "e1z", "e1r", "e2z", "e2r")
    init = {k: rng.standard_normal((nr, nz)) for k in names}          # stale contents the rim keeps
    tau_init = (rng.standard_normal((nr, nz)), rng.standard_normal((nr, nz)))
    w0 = _rand(rng, nr, nz, 3.0)
    results = {}
    for path in ("march", "tiled"):
        _lib.call("axb_set_solid_march", 1 if path == "march" else 0)
        try:
            out = {}
            for tag, c in (("plain", None), ("blend", chi)):
                o = {k: v.copy() for k, v in init.items()}
                K.solid_sigma(o["s11"], o["s12"], o["s22"], 3.7, dx, eta1, eta2, o["e1z"], o["e1r"], o["e2z"], o["e2r"],
                              _chi=c)
                out[tag] = o
            b = out["blend"]
            tz, tr, w = tau_init[0].copy(), tau_init[1].copy(), w0.copy()
            K.update_vorticity_from_solid_stress(w, tz, tr, b["s11"], b["s12"], b["s22"], R, 2e-3, dx)
            out["tau"] = {"tz": tz, "tr": tr, "w": w}
            results[path] = out
        finally:
            _lib.call("axb_set_solid_march", 0)
    for tag in ("plain", "blend", "tau"):
        for k in results["tiled"][tag]:
            assert np.array_equal(results["march"][tag][k], results["tiled"][tag][k]), (tag, k)
    # and against the oracle (the tiled path is the reference's sequence: 1e-13 covers numba-vs-NumPy last bits)
    o = {k: v.copy() for k, v in init.items()}
    ox.solid_sigma(o["s11"], o["s12"], o["s22"], 3.7, dx, eta1, eta2, o["e1z"], o["e1r"], o["e2z"], o["e2r"])
    for k in names:
        assert_close(results["march"]["plain"][k], o[k], 1e-13, "solid_sigma " + k)
    tz, tr, w = tau_init[0].copy(), tau_init[1].copy(), w0.copy()
    b = results["march"]["blend"]
    ox.update_vorticity_from_solid_stress(w, tz, tr, b["s11"], b["s12"], b["s22"], R, 2e-3, dx)
    assert_close(results["march"]["tau"]["tz"], tz, 1e-13, "tau_z")
    assert_close(results["march"]["tau"]["tr"], tr, 1e-13, "tau_r")
    assert_close(results["march"]["tau"]["w"], w, 1e-13, "vorticity after the solid stress")


# ---------------------------------------------------------------------------------------------
# the rest of the C++ core family (core/src/instantiate.yml), goldens from the reference C++ itself
# ---------------------------------------------------------------------------------------------
PARTICLE_KERNELS = ("linear_kernel", "mp4", "mp6", "yang_smooth_three_point_kernel")


def test_core_particle_family_against_reference_cpp():
    """mesh_to_particles_2D_* / particles_to_mesh_2D_* (periodic and "unbounded", four kernels), the 1-D MP4 pair and
    the wrap routines against tests/golden/core_family.npz (reference C++ compiled as-is).  Gathers repeat the
    reference's summation order: bit-identical (Yang's kernel goes through asin / sqrt: <= 4 ulp).  Scatters add
    with atomics: indices and weights exact, sums to 1e-14."""
    import pyaxisymflow_b200.core.mesh_to_particles as m2p
    import pyaxisymflow_b200.core.particles_to_mesh as p2m

    g = golden("core_family")
    dx, dy = float(g["dx"]), float(g["dy"])
    for k in PARTICLE_KERNELS:
        for per in (True, False):
            mid = "" if per else "unbounded_"
            px, py = (g["pxw"], g["pyw"]) if per else (g["px"], g["py"])
            ox_, oy_ = np.full(px.shape, 3.0), np.full(px.shape, 3.0)
            getattr(m2p, f"mesh_to_particles_2D_{mid}{k}")(g["fx"], g["fy"], px, py, ox_, oy_, dx, dy)
            if k.startswith("yang"):
                assert_close(ox_, g[f"m2p_{mid}{k}_x"], 1e-15, k)
                assert_close(oy_, g[f"m2p_{mid}{k}_y"], 1e-15, k)
            else:
                assert np.array_equal(ox_, g[f"m2p_{mid}{k}_x"]), (k, per)
                assert np.array_equal(oy_, g[f"m2p_{mid}{k}_y"]), (k, per)
            mesh = np.full(g["fx"].shape, 3.0)
            getattr(p2m, f"particles_to_mesh_2D_{mid}{k}")(px, py, g["val"], mesh, dx, dy)
            ref = g[f"p2m_{mid}{k}"]
            assert np.array_equal(mesh != 0, ref != 0), (k, per)          # same cells touched: index work exact
            assert_close(mesh, ref, 1e-14, f"p2m {mid}{k}")
    o1 = np.zeros_like(g["q1"])
    m2p.mesh_to_particles_1D_mp4(g["f1"], g["q1"], o1, dx)
    assert np.array_equal(o1, g["m2p_1d"])
    m1 = np.ones_like(g["f1"])
    p2m.particles_to_mesh_1D_mp4(g["q1"], g["v1"], m1, dx)
    assert_close(m1, g["p2m_1d"], 1e-14, "p2m 1-D")
    wx, wy = g["wrap_x0"].copy(), g["wrap_y0"].copy()
    m2p.wrap_particles_around_2D_domain(wx, wy, 0.0, 1.0, 0.0, 0.5)
    assert np.array_equal(wx, g["wrap_x"]) and np.array_equal(wy, g["wrap_y"])
    w1 = g["wrap1_in"].copy()
    m2p.wrap_particles_around_1D_domain(w1, 0.0, 1.0)
    assert np.array_equal(w1, g["wrap1_out"])
    sx, sy = g["wrap_small_in"].copy(), g["wrap_small_in"].copy()
    m2p.wrap_particles_around_2D_domain(sx, sy, 0.0, 1.0, 0.0, 1.0)
    assert np.array_equal(sx, g["wrap_small_x"]) and np.array_equal(sy, g["wrap_small_y"])
    with pytest.raises(TypeError):
        m2p.mesh_to_particles_2D_mp4(g["fx"].astype(np.float32), g["fy"], g["px"], g["py"], ox_, oy_, dx, dy)


@pytest.mark.parametrize("kernel", PARTICLE_KERNELS)
def test_core_particle_family_large_against_oracle(kernel):
    """1024 x 2048 mesh (config C5's grid), 2 M particles displaced by up to 1.7 cells: gather bit-identical to the
    oracle, scatter to 1e-13; interpolating a constant field returns the constant (partition of unity, periodic)."""
    rng = np.random.default_rng(5)
    m0, m1 = 1024, 2048
    dx = 1.0 / m1
    fx, fy = rng.standard_normal((m0, m1)), rng.standard_normal((m0, m1))
    px = (np.arange(m1)[None, :] + 0.5 + 1.7 * rng.uniform(-1, 1, (m0, m1))) * dx
    py = (np.arange(m0)[:, None] + 0.5 + 1.7 * rng.uniform(-1, 1, (m0, m1))) * dx
    val = rng.standard_normal((m0, m1))
    import pyaxisymflow_b200.core.mesh_to_particles as m2p
    import pyaxisymflow_b200.core.particles_to_mesh as p2m

    a, b, c, d = (np.zeros((m0, m1)) for _ in range(4))
    getattr(m2p, f"mesh_to_particles_2D_unbounded_{kernel}")(fx, fy, px, py, a, b, dx, dx)
    ox.mesh_to_particles_2D(kernel, False, fx, fy, px, py, c, d, dx, dx)
    if kernel.startswith("yang"):
        assert_close(a, c, 1e-15, kernel)
        assert_close(b, d, 1e-15, kernel)
    else:
        assert np.array_equal(a, c) and np.array_equal(b, d)
    mesh, ref = np.zeros((m0, m1)), np.zeros((m0, m1))
    getattr(p2m, f"particles_to_mesh_2D_unbounded_{kernel}")(px, py, val, mesh, dx, dx)
    ox.particles_to_mesh_2D(kernel, False, px, py, val, ref, dx, dx)
    assert_close(mesh, ref, 1e-13, "scatter " + kernel)
    pxw, pyw = np.mod(px, m1 * dx), np.mod(py, m0 * dx)
    ones = np.ones((m0, m1))
    getattr(m2p, f"mesh_to_particles_2D_{kernel}")(ones, 2 * ones, pxw, pyw, a, b, dx, dx)
    assert np.max(np.abs(a - 1)) < 1e-13 and np.max(np.abs(b - 2)) < 1e-13


def test_least_squares_extrapolation_second_order():
    """extrapolate_using_least_squares_till_second_order (core/src/extrapolate_using_least_squares.hpp:469-486)
    against the reference C++ output: flags and values bit-exact, like the first-order routine."""
    import pyaxisymflow_b200.core.extrapolate_using_least_squares as els

    g = golden("core_family")
    for order, fn in ((1, els.extrapolate_using_least_squares_till_first_order),
                      (2, els.extrapolate_using_least_squares_till_second_order)):
        c, a, b = g["ls_cur"].copy(), g["ls_ex"].copy(), g["ls_ey"].copy()
        fn(c, g["ls_tgt"], a, b, g["ls_gx"], g["ls_gy"])
        assert np.array_equal(c, g[f"ls{order}_cur"])
        assert np.array_equal(a, g[f"ls{order}_ex"], equal_nan=True), order
        assert np.array_equal(b, g[f"ls{order}_ey"], equal_nan=True), order


def test_static_pde_extrapolation_against_reference():
    """StaticPDEExtrapolation (examples/PeriodicSoftSlab/bounded_static_PDE_extrapolation.py) against outputs of the
    reference class itself (tests/golden/static_pde.npz): a sphere away from the walls and a z-periodic slab.  Same
    Jacobi iterates (every sweep reads the previous one only), termination decided on the device by the same 2-norm."""
    from pyaxisymflow_b200.ops import gen_periodic_boundary_ghost_comm
    from pyaxisymflow_b200.static_pde_extrapolation import StaticPDEExtrapolation

    g = golden("static_pde")
    nr, nz = g["box_phi"].shape
    dx = float(g["box_dx"])
    eta = g["box_eta0"].copy()
    s = StaticPDEExtrapolation(dx, nr, nz, float(g["box_tol"]), float(g["box_band"]))
    s.extrapolate(eta, g["box_phi"].copy())
    assert [s.r_start, s.r_end, s.z_start, s.z_end] == list(g["box_bounds"])
    assert_close(eta, g["box_eta"], 1e-12, "extrapolated eta (sphere)")
    assert s.sweeps[0] > 3 and s.sweeps[1] > 3
    # the extrapolated field is constant along the normals: outside the solid, inside the band, n . grad(eta) ~ grad_n
    assert np.count_nonzero(eta) > np.count_nonzero(g["box_eta0"])
    per = gen_periodic_boundary_ghost_comm(2)
    eta2, phi2 = g["slab_eta0"].copy(), g["slab_phi"].copy()
    s2 = StaticPDEExtrapolation(dx, nr, nz, float(g["slab_tol"]), float(g["slab_band"]), periodic=True,
                                per_communicator_gen=per, per_communicator_eta=per)
    s2.extrapolate(eta2, phi2)
    assert [s2.r_start, s2.r_end, s2.z_start, s2.z_end] == list(g["slab_bounds"])
    assert np.array_equal(phi2, g["slab_phi_after"])
    assert_close(eta2, g["slab_eta"], 1e-12, "extrapolated eta (periodic slab)")
    with pytest.raises(ValueError):
        StaticPDEExtrapolation(dx, nr, nz, 1e-6, 6 * dx, periodic=True)
    # device-resident call: tensors in place
    import torch

    te, tp = torch.from_numpy(g["box_eta0"].copy()).cuda(), torch.from_numpy(g["box_phi"].copy()).cuda()
    StaticPDEExtrapolation(dx, nr, nz, float(g["box_tol"]), float(g["box_band"])).extrapolate(te, tp)
    assert_close(te.cpu().numpy(), g["box_eta"], 1e-12, "device-resident call")


def test_cycle_averages_restart_every_cycle(K):
    """the reference zeroes its running averages when the oscillation-cycle timer wraps
    (particle_in_bubble_oscillatory_flow.py:168-263, soft_sphere_streaming.py:139-165); the steppers hand the
    completed-cycle averages to `on_cycle` and start again from zero."""
    import torch
    from pyaxisymflow_b200.timestep import ParticleFlowStepper, SoftSphereStepper

    p = ParticleFlowStepper(64, freq=16.0, e=0.02)
    seen = []
    p.on_cycle = lambda s: seen.append((s.it, s.avg_vort.clone(), s.avg_time, s.avg_Z_cm))
    p.step(20)
    assert p.cycles == 0 and 0.0 < p.freqTimer < p.freqTimer_limit
    full_avg, avg_time = p.avg_vort.clone(), p.avg_time
    assert full_avg.abs().max().item() > 0 and avg_time > 0
    p.freqTimer = p.freqTimer_limit               # (a cycle is > 100 steps) jump to its end: the next step wraps
    chi_before = p.part_char_func.clone()
    p.step(1)
    assert p.cycles == 1 and len(seen) == 1 and len(p.avg_T) == 1 and len(p.avg_part_trajectory) == 1
    assert seen[0][0] == 20 and torch.equal(seen[0][1], full_avg)          # the hook saw the completed averages
    assert abs(p.avg_T[0] - avg_time * 16.0) <= 1e-15
    # what has accumulated since the wrap is exactly one step's worth, not the whole run
    want = chi_before * (p.dt / p.freqTimer_limit)
    assert float((p.avg_part_char_func - want).abs().max()) <= 1e-15 * float(want.abs().max())
    assert abs(p.freqTimer - p.dt) <= 1e-18 and abs(p.avg_time - (p.t - p.dt) * p.dt) <= 1e-18

    s = SoftSphereStepper(64, Z_cm=0.47)
    s.step(3)
    assert float(s.avg_phi.abs().max()) > 0.0
    s.freqTimer = 0.9999 * s.freqTimer_limit                   # (a cycle is ~600 steps at this size) jump to its end
    calls = []
    s.on_cycle = lambda st: calls.append(float(st.avg_phi.abs().max()))
    s.step(1)                                                  # dt is clipped to end the cycle exactly
    assert s.cycles == 1 and len(calls) == 1 and calls[0] > 0.0
    assert float(s.avg_psi.abs().max()) == 0.0 and float(s.avg_phi.abs().max()) == 0.0 and s.freqTimer == 0.0
    torch.cuda.synchronize()


def _particle_state(m):
    return (m.t, m.dt, m.U_z_cm_part, m.part_Z_cm, m.F_total, m.it, m.freqTimer, m.avg_Z_cm, m.avg_time, m.diff)


def _assert_particle_state(got, want, tag=""):
    """the force is brink_lam (1e12) times a cancelling sum accumulated with atomics (order varies run to run), so it,
    and the particle velocity it integrates to, agree to the sum's conditioning; everything else to rounding"""
    names = "t dt U Z F it timer avgZ avgT diff".split()
    for a, b, name in zip(_particle_state(got), _particle_state(want), names):
        tol = 1e-6 if name in ("U", "F", "diff") else 1e-10
        assert abs(a - b) <= tol * max(abs(b), 1e-300), (tag, name, a, b)


@pytest.mark.parametrize("graph", [False, True])
def test_particle_stepper_device_scalars_match_host_loop(K, graph):
    """SURVEY 8f-1 (config C5): with ``device_scalars=True`` every host decision of
    particle_in_bubble_oscillatory_flow.py:168-170, 255-270, 297-301, 323-355 (dt, the cycle timer, force and the
    rigid-body update) is taken by axb_particle_scalars on a device block and the step can be replayed as a CUDA
    graph; the fields and the scalars follow the host-driven stepper (the only different arithmetic is the device's
    sin(omega t))."""
    from pyaxisymflow_b200.timestep import ParticleFlowStepper

    nz, steps = 80, 9
    h = ParticleFlowStepper(nz, freq=16.0, e=0.02)
    d = ParticleFlowStepper(nz, freq=16.0, e=0.02, solver=h.solver, device_scalars=True, use_graph=graph,
                            trace_capacity=6)
    h.step(steps)
    d.step(4)
    d.step(steps - 4)
    d.sync_scalars()
    _assert_particle_state(d, h)
    for f in ("vorticity", "psi", "avg_vort", "avg_psi", "avg_part_char_func", "part_char_func", "u_z", "u_r"):
        assert_close(getattr(d, f).cpu().numpy(), getattr(h, f).cpu().numpy(), 1e-9, f)
    # the trace ring (capacity 6 < 9 steps) holds the last six rows, oldest first
    assert len(d.trace) == 6
    for got, want in zip(d.trace, h.trace[-6:]):
        for a, b in zip(got, want):
            assert abs(a - b) <= 1e-6 * max(abs(b), 1e-300)
    # cycle wrap decided on the device: averages restart, the completed ones are kept in *_last
    full = d.avg_vort.clone()
    chi_before = d.part_char_func.clone()
    d.state[11] = d.freqTimer_limit
    d.step(1)
    d.sync_scalars()
    assert d.cycles == 1 and len(d.avg_T) == 1 and len(d.avg_part_trajectory) == 1
    import torch
    assert torch.equal(d.avg_vort_last, full)
    want = chi_before * (d.dt / d.freqTimer_limit)
    assert float((d.avg_part_char_func - want).abs().max()) <= 1e-15 * float(want.abs().max())
    assert abs(d.freqTimer - d.dt) <= 1e-18


@pytest.mark.parametrize("fft,launch,nz", [(True, "batched", 512), (False, "batched", 128), (True, "members", 128)])
def test_batched_particle_ensemble(K, fft, launch, nz):
    """SURVEY 8e "Ensemble": members stored as column blocks of shared (nr, batch nz) tensors, one solve for the
    whole ensemble (DCT rows = batch nr, sweep columns = batch nz), per-member scalars on the device, the whole
    ensemble step one replayed graph -- every member equals the same member stepping alone under host control.
    launch="batched": every operation is one launch over all members (axb_grid_t.batch; 512 columns reach the
    interior row-marching kernels); "members": one launch per member and operation on parallel graph branches."""
    from pyaxisymflow_b200.fd import FastDiagonalisationStokesSolver
    from pyaxisymflow_b200.timestep import ParticleEnsemble, ParticleFlowStepper

    nr, steps = 64, 7
    params = [(8.0, 0.01), (16.0, 0.02), (12.0, 0.005), (20.0, 0.01), (24.0, 0.015)]
    kw = dict(basis="analytic", r_method="tridiagonal", z_method="fft") if fft else {}
    solver = FastDiagonalisationStokesSolver(nr, nz, 1.0 / nz, **kw)
    ens = ParticleEnsemble.batched_ensemble(params, nz, nr, use_graph=True, branches=3, solver=solver, launch=launch)
    assert (ens._solver is not None) == fft
    assert ens.members[2].vorticity.stride(0) == len(params) * nz
    hosts = [ParticleFlowStepper(nz, grid_size_r=nr, freq=f, e=e, solver=solver) for f, e in params]
    # the first step has no feedback from the (ill-conditioned) force yet: fields agree to rounding
    ens.step(1)
    for m, h in zip(ens.members, hosts):
        h.step(1)
        for fld in ("vorticity", "psi", "u_z", "u_r", "avg_vort", "part_char_func"):
            assert_close(getattr(m, fld).cpu().numpy(), getattr(h, fld).cpu().numpy(), 1e-12, f"{fld} after one step")
    ens.step(2)
    ens.step(steps - 3)
    ens.sync_scalars()
    # later steps carry U_z_cm_part = integral of F (agrees to the conditioning of the force sum, see
    # _assert_particle_state) into the penalised velocity, so the fields agree to that, not to rounding
    for m, h, (f, e) in zip(ens.members, hosts, params):
        h.step(steps - 1)
        _assert_particle_state(m, h, f)
        for fld in ("vorticity", "psi", "avg_vort", "part_char_func"):
            assert_close(getattr(m, fld).cpu().numpy(), getattr(h, fld).cpu().numpy(), 1e-6, f"{fld} f={f}")


@pytest.mark.parametrize("graph,nz", [(False, 64), (True, 64), (True, 320)])
def test_soft_sphere_stepper_device_scalars_match_host_loop(K, graph, nz):
    """SURVEY 8f-1 (config C3): dt with its cycle / tEnd clamps, the tether's velocity and position and the cycle timer
    on a device block (axb_soft_sphere_scalars, soft_sphere_streaming.py:139-176, 201-203, 262-264), LS extrapolation
    with device-terminated sweeps (axb_ls_extrapolate_eta_device), no buffer swap: the step is a fixed launch sequence
    without a host read, replayed as a CUDA graph.  Same kernels on the same inputs as the host-driven stepper; the
    only different arithmetic is the device's sin / cos of omega t."""
    import torch
    from pyaxisymflow_b200.timestep import SoftSphereStepper

    steps = 7
    h = SoftSphereStepper(nz, Z_cm=0.47)
    d = SoftSphereStepper(nz, Z_cm=0.47, device_scalars=True, use_graph=graph)
    h.step(steps)
    d.step(3)
    d.step(steps - 3)
    d.sync_scalars()
    assert d.it == steps and abs(d.t - h.t) <= 1e-13 * h.t and abs(d.dt - h.dt) <= 1e-13 * h.dt
    assert d.ls_sweeps == h.ls_sweeps >= 4
    for f in ("eta1", "eta2", "ball_phi", "vorticity", "psi", "avg_psi", "avg_phi", "ball_char_func", "tether_char_func",
              "u_z", "u_r"):
        assert_close(getattr(d, f).cpu().numpy(), getattr(h, f).cpu().numpy(), 1e-11, f)
    # the LS sweeps forked onto a side stream (default) against the single-stream launch order: same kernels, same data
    one = SoftSphereStepper(nz, Z_cm=0.47, device_scalars=True, use_graph=graph, overlap_ls=False)
    assert d.overlap_ls and not one.overlap_ls
    one.step(steps)
    one.sync_scalars()
    for f in ("eta1", "eta2", "ball_phi", "vorticity", "psi", "tether_char_func", "u_z", "u_r"):
        assert torch.equal(getattr(d, f), getattr(one, f)), f
    # cycle wrap on the device: the step that completes the cycle is clamped to its end, the next one restarts the
    # averages and keeps the completed ones
    d.state[3] = d.freqTimer_limit - 0.25 * d.dt
    d.step(1)
    d.sync_scalars()
    assert d.cycles == 1 and d.freqTimer == 0.0 and abs(d.dt - 0.25 * h.dt) <= 1e-9 * h.dt
    full = d.avg_psi.clone()
    psi_before = d.psi.clone()
    d.step(1)
    d.sync_scalars()
    assert torch.equal(d.avg_psi_last, full)
    # the restarted average holds this step's psi (solved at the start of the step) times dt
    assert float((d.avg_psi - d.psi * d.dt).abs().max()) <= 1e-15 * float((d.psi * d.dt).abs().max())
    assert not torch.equal(psi_before, d.psi)
    # a sweep budget that is too small is reported, not silently accepted
    few = SoftSphereStepper(nz, Z_cm=0.47, device_scalars=True, ls_sweeps=2)
    few.step(1)
    with pytest.raises(Exception, match="sweep budget"):
        few.sync_scalars()


@pytest.mark.parametrize("nr,nz", [(40, 70), (16, 64), (3, 3), (35, 130), (50, 200), (17, 65), (130, 1030)])
def test_fused_solid_stress_equals_the_three_calls(K, nr, nz):
    """axb_solid_stress_vorticity_update (the soft-sphere stepper's form of a18 + a19): one shared-memory pass
    eta -> sigma -> tau -> w, bit-identical to solid_sigma -> update_vorticity_from_solid_stress on zero-initialised
    work arrays (what the reference driver has), and against the oracle's sequence."""
    import ctypes

    import torch
    from pyaxisymflow_b200 import _lib
    from pyaxisymflow_b200.device import make_grid, ptr, stream_ptr

    rng = np.random.default_rng(nr * 11 + nz)
    dx, _, r, Z, R = _grid(nr, nz)
    eta1, eta2 = Z + 0.05 * _rand(rng, nr, nz), R + 0.05 * _rand(rng, nr, nz)
    chi = np.clip(_rand(rng, nr, nz) + 0.5, 0, 1)
    w0 = _rand(rng, nr, nz, 3.0)
    G, dt = 3.7, 2e-3
    for c in (chi, None):
        z = {k: np.zeros((nr, nz)) for k in ("s11", "s12", "s22", "e1z", "e1r", "e2z", "e2r", "tz", "tr")}
        K.solid_sigma(z["s11"], z["s12"], z["s22"], G, dx, eta1, eta2, z["e1z"], z["e1r"], z["e2z"], z["e2r"], _chi=c)
        want = w0.copy()
        K.update_vorticity_from_solid_stress(want, z["tz"], z["tr"], z["s11"], z["s12"], z["s22"], R, dt, dx)
        g = make_grid(nr, nz, nz, dx)
        dev = [torch.from_numpy(a).cuda() for a in (w0.copy(), eta1, eta2, chi, r)]
        _lib.call("axb_solid_stress_vorticity_update", ctypes.byref(g), ptr(dev[0]), ptr(dev[1]), ptr(dev[2]),
                  ptr(dev[3]) if c is not None else None, ptr(dev[4]), G, dt, None, 1, stream_ptr())
        got = dev[0].cpu().numpy()
        assert np.array_equal(got, want), ("blend" if c is not None else "plain", np.abs(got - want).max())
        # reciprocal multiplications instead of the divisions by 2 dx and r (the stepper's default)
        dev[0].copy_(torch.from_numpy(w0))
        _lib.call("axb_solid_stress_vorticity_update", ctypes.byref(g), ptr(dev[0]), ptr(dev[1]), ptr(dev[2]),
                  ptr(dev[3]) if c is not None else None, ptr(dev[4]), G, dt, None, 0, stream_ptr())
        assert_close(dev[0].cpu().numpy(), want, 1e-12, "fused solid stress, reciprocal form")
        # the oracle's sequence (numba-vs-NumPy last bits)
        o = {k: np.zeros((nr, nz)) for k in z}
        ox.solid_sigma(o["s11"], o["s12"], o["s22"], G, dx, eta1, eta2, o["e1z"], o["e1r"], o["e2z"], o["e2r"])
        if c is not None:
            for k in ("s11", "s12", "s22"):
                o[k] = c * o[k]
        wo = w0.copy()
        ox.update_vorticity_from_solid_stress(wo, o["tz"], o["tr"], o["s11"], o["s12"], o["s22"], R, dt, dx)
        assert_close(got, wo, 1e-12, "fused solid stress vs oracle")


@pytest.mark.parametrize("nr,nz", [(24, 64), (37, 53), (8, 130), (5, 4), (16, 1028)])
def test_heaviside_with_mask(K, nr, nz):
    """axb_smooth_heaviside_mask (soft_sphere_streaming.py:205-206): H = smooth_Heaviside(phi) and the dense uint8
    mask H > 0.5 in one pass -- four columns per thread where the rows allow it, scalar otherwise; same bits as the
    plain Heaviside entry, which is pinned by the golden vectors."""
    import ctypes

    import torch
    from pyaxisymflow_b200 import _lib
    from pyaxisymflow_b200.device import make_grid, ptr, stream_ptr

    rng = np.random.default_rng(nr + nz)
    dx = 1.0 / nz
    phi = (rng.standard_normal((nr, nz)) * 3 * dx)
    want = np.zeros_like(phi)
    ox.smooth_Heaviside(want, phi, 2 * dx)
    g = make_grid(nr, nz, nz, dx)
    d_phi = torch.from_numpy(phi).cuda()
    H = torch.full_like(d_phi, -1.0)
    for ge in (0, 1):
        mask = torch.full((nr, nz), 7, dtype=torch.uint8, device="cuda")
        _lib.call("axb_smooth_heaviside_mask", ctypes.byref(g), ptr(H), ptr(mask), ptr(d_phi), 2 * dx, 0.5, ge, stream_ptr())
        got = H.cpu().numpy()
        assert_close(got, want, 1e-14, "Heaviside with mask")
        m = mask.cpu().numpy().astype(bool)
        assert np.array_equal(m, (got >= 0.5) if ge else (got > 0.5))
