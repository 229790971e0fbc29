"""GPU tests of the r-slab pieces that can be checked on ONE device: the two halves of kill_boundary_vorticity_sine_r,
the owned-row window of the fused reductions (axb_grid_t.ju0 / ju1), the row-halo put, a sequential emulation of the
P-rank r-slab step (every "rank" is a block on the same GPU, halos copied by hand) against the single-GPU stepper,
and the world-size-1 RowSlabRigidFlowStepper.  The real multi-process run is tests/test_multigpu_cuda.py."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.mark.parametrize("tiled", [0, 1])
def test_kill_r_parts_and_owned_row_reductions(T, tiled):
    from pyaxisymflow_b200 import _lib
    from pyaxisymflow_b200.device import make_grid, ptr, stream_ptr

    torch = T
    _lib.call("axb_set_stencil_path", tiled)
    try:
        nr, nz = 70, 300
        dx = 1.0 / nz
        rng = np.random.default_rng(3)
        r1d = torch.from_numpy(np.linspace(dx / 2, nr * dx - dx / 2, nr)).cuda()
        w0 = torch.from_numpy(rng.standard_normal((nr, nz))).cuda()
        g = make_grid(nr, nz, nz, dx)
        both = w0.clone()
        _lib.call("axb_kill_boundary_vorticity_sine_r", ctypes.byref(g), ptr(both), ptr(r1d), 3, stream_ptr())
        a, b = w0.clone(), w0.clone()
        _lib.call("axb_kill_boundary_vorticity_sine_r_parts", ctypes.byref(g), ptr(a), ptr(r1d), 3, 1, stream_ptr())
        _lib.call("axb_kill_boundary_vorticity_sine_r_parts", ctypes.byref(g), ptr(b), ptr(r1d), 3, 2, stream_ptr())
        assert torch.equal(a[-3:], both[-3:]) and torch.equal(a[:-3], w0[:-3])
        assert torch.equal(b[0], both[0]) and torch.equal(b[1:], w0[1:])
        # ---- fused reductions over the owned rows [ju0, ju1) only; the fields themselves do not change
        psi = torch.from_numpy(rng.standard_normal((nr, nz))).cuda()
        for ju0, ju1 in ((0, nr), (2, nr - 2), (5, 37), (33, 34)):
            gw = make_grid(nr, nz, nz, dx, rows=(ju0, ju1))
            uz, ur = torch.zeros_like(psi), torch.zeros_like(psi)
            st = torch.zeros(8, dtype=torch.float64, device="cuda")
            sp = lambda i: ctypes.c_void_p(st.data_ptr() + 8 * i)  # noqa: E731
            _lib.call("axb_velocity_from_psi", ctypes.byref(gw), ptr(uz), ptr(ur), ptr(psi), ptr(r1d), 0.25, 0.0, None,
                      sp(2), stream_ptr())
            uz_all, ur_all = torch.zeros_like(psi), torch.zeros_like(psi)
            _lib.call("axb_velocity_from_psi", ctypes.byref(g), ptr(uz_all), ptr(ur_all), ptr(psi), ptr(r1d), 0.25, 0.0,
                      None, None, stream_ptr())
            assert torch.equal(uz, uz_all) and torch.equal(ur, ur_all)
            want = (uz_all.abs() + ur_all.abs())[ju0:ju1].max().item()
            assert st[2].item() == want, (ju0, ju1)
            chi = torch.rand((nr, nz), dtype=torch.float64, device="cuda")
            w = w0.clone()
            pz, pr = torch.zeros_like(psi), torch.zeros_like(psi)
            _lib.call("axb_penalise_update_vorticity", ctypes.byref(gw), ptr(pz), ptr(pr), ptr(w), ptr(uz_all),
                      ptr(ur_all), ptr(chi), 1e3, 1e-3, None, 0.1, 0.0, None, ptr(r1d), sp(3), stream_ptr())
            want = (r1d[:, None] * chi * (pz - 0.1))[ju0:ju1].sum().item()
            assert abs(st[3].item() - want) <= 1e-11 * max(1.0, abs(want)), (ju0, ju1)
    finally:
        _lib.call("axb_set_stencil_path", 0)


def test_row_halo_put_on_one_device(T):
    """two 'ranks' as two buffers of the same process: the put must land the owned edge rows in the other's halos"""
    from pyaxisymflow_b200 import _lib
    from pyaxisymflow_b200.device import stream_ptr

    torch = T
    nz, nrl, H = 130, 9, 2
    a = torch.arange((nrl + 2 * H) * nz, dtype=torch.float64, device="cuda").reshape(nrl + 2 * H, nz)
    b = -a.clone()
    a0, b0 = a.clone(), b.clone()
    arr = ctypes.c_uint64 * 1
    # rank "a" is below rank "b": a's last owned rows -> b's lower halo, b's first owned rows -> a's upper halo
    for width in (1, 2):
        a.copy_(a0)
        b.copy_(b0)
        for entry in ("axb_row_halo_put", "axb_row_halo_get"):
            a.copy_(a0)
            b.copy_(b0)
            _lib.call(entry, 1, arr(a.data_ptr()), arr(0), arr(b.data_ptr()), nz, nz, nrl, H, width, stream_ptr())
            _lib.call(entry, 1, arr(b.data_ptr()), arr(a.data_ptr()), arr(0), nz, nz, nrl, H, width, stream_ptr())
            assert torch.equal(b[H - width:H], a0[H + nrl - width:H + nrl]), entry
            assert torch.equal(a[H + nrl:H + nrl + width], b0[H:H + width]), entry
            keep_a = torch.ones(nrl + 2 * H, dtype=torch.bool)
            keep_a[H + nrl:H + nrl + width] = False
            assert torch.equal(a[keep_a], a0[keep_a]), entry


@pytest.mark.parametrize("P", [2, 4])
def test_emulated_rowslab_step_matches_single_gpu(T, P):
    """the r-slab schedule with the real CUDA kernels, P blocks processed one after the other on one GPU and the
    halos copied by hand; the solve is done on the gathered field (the partitioned solve has its own test).  Every
    owned value must equal the single-GPU stepper's bit for bit."""
    from pyaxisymflow_b200.rowslab import CudaOps, RowSlabLayout
    from pyaxisymflow_b200.timestep import RigidFlowStepper

    torch = T
    nz, nr, steps = 512, 128, 3
    ref = RigidFlowStepper(nz, grid_size_r=nr)
    ref.seed_vorticity()
    seed = ref.vorticity.clone()
    dx = ref.dx
    r_full = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    Ls = [RowSlabLayout(nr, nz, P, p) for p in range(P)]
    names = ("w", "psi", "uz", "ur", "uzu", "uru", "chi", "tmp", "w2")
    F = [{n: torch.zeros((L.nrs, nz), dtype=torch.float64, device="cuda") for n in names} for L in Ls]
    st = [torch.zeros(8, dtype=torch.float64, device="cuda") for _ in Ls]
    ops = [CudaOps(L, dx, torch.from_numpy(r_full[L.g0:L.g0 + L.nv].copy()).cuda(), ref.z1d, s, ref.nu, ref.brink_lam)
           for L, s in zip(Ls, st)]
    for L, f, o in zip(Ls, F, ops):
        f["w"].copy_(L.scatter_global(seed))
        o.heaviside_sphere(L.block(f["chi"]), 0.25, 0.0, ref.r_sph)

    def exchange(name):
        H = 2
        for p in range(P - 1):
            lo, up = F[p][name], F[p + 1][name]
            n = Ls[p].nrl
            up[0:H].copy_(lo[n:n + H])                 # last owned rows of p -> lower halo of p+1
            lo[H + n:H + n + H].copy_(up[H:2 * H])      # first owned rows of p+1 -> upper halo of p

    sc = (ref.U_0, ref.T_ramp, 0.0, ref.dt_diff_limit, ref.CFL * dx)
    for _ in range(steps):
        for L, f, o in zip(Ls, F, ops):
            B = L.block
            o.scalars(0, sc)
            o.kill_z(B(f["w"]))
            parts = (1 if L.upper is None else 0) | (2 if L.lower is None else 0)
            if parts:
                o.kill_r(B(f["w"]), parts)
        rhs = torch.cat([L.owned(f["w"]) for L, f in zip(Ls, F)], dim=0).contiguous()
        psi = torch.zeros_like(rhs)
        ref.solver.solve(psi, rhs)
        for L, f in zip(Ls, F):
            L.owned(f["psi"]).copy_(psi[L.r_begin:L.r_begin + L.nrl])
        exchange("psi")
        for L, f, o in zip(Ls, F, ops):
            o.velocity(L.block(f["uzu"]), L.block(f["uru"]), L.block(f["psi"]))
        umax = torch.stack([s[2] for s in st]).max()
        for s in st:
            s[2] = umax
        for L, f, o in zip(Ls, F, ops):
            B = L.block
            o.scalars(1, sc)
            o.penalise(B(f["uz"]), B(f["ur"]), B(f["w"]), B(f["uzu"]), B(f["uru"]), B(f["chi"]))
        exchange("w")
        exchange("ur")
        for L, f, o in zip(Ls, F, ops):
            B = L.block
            o.advect(B(f["w2"]), B(f["w"]), B(f["uz"]), B(f["ur"]))
        exchange("w2")
        for L, f, o in zip(Ls, F, ops):
            B = L.block
            o.diffuse(B(f["w"]), B(f["w2"]), B(f["tmp"]))
            o.scalars(2, sc)
    ref.step(steps)
    torch.cuda.synchronize()
    got = torch.cat([L.owned(f["w"]) for L, f in zip(Ls, F)], dim=0)
    assert torch.equal(got, ref.vorticity), (got - ref.vorticity).abs().max().item()
    assert st[0][0].item() == ref.state[0].item() and st[0][1].item() == ref.state[1].item()
    drag = sum(s[7].item() for s in st)
    assert abs(drag - ref.state[7].item()) <= 1e-10 * max(1.0, abs(ref.state[7].item()))


def test_rowslab_stepper_world1_matches_rigid_stepper(T):
    from pyaxisymflow_b200.rowslab import RowSlabRigidFlowStepper
    from pyaxisymflow_b200.timestep import RigidFlowStepper

    torch = T
    nz, nr = 1024, 256
    a = RowSlabRigidFlowStepper(nz, grid_size_r=nr)
    b = RigidFlowStepper(nz, grid_size_r=nr, basis="analytic", r_method="tridiagonal", z_method="fft")   # same solve path
    a.seed_vorticity()
    b.seed_vorticity()
    assert torch.equal(a.gather_vorticity(), b.vorticity)
    # the solve alone: DCT-II + own-block sweeps + (trivial) partition correction + DCT-III vs axb_fd_solve
    pa, pb = torch.zeros_like(b.vorticity), torch.zeros_like(b.vorticity)
    a.solver.solve(pa, b.vorticity.clone())
    b.solver.solve(pb, b.vorticity.clone())
    torch.cuda.synchronize()
    e0 = ((pa - pb).abs().max() / pb.abs().max()).item()
    assert e0 < 1e-13, e0
    a.step(4)
    b.step(4)
    torch.cuda.synchronize()
    err = ((a.gather_vorticity() - b.vorticity).abs().max() / b.vorticity.abs().max()).item()
    assert err < 1e-10, err      # rounding in the solve, amplified by ENO3 stencil switches on the noisy seed
    sa, sb = a.scalars(), b.scalars()
    assert abs(sa["t"] - sb["t"]) <= 1e-14 * sb["t"] and sa["iterations"] == sb["iterations"] == 4
    ph = a.phase_times(2)
    assert set(ph) >= {"solve_dct2", "solve_r_partitioned", "solve_dct3", "halo_psi", "advect", "diffuse"}


def test_rowslab_stepper_world1_periodic_matches_rigid_stepper(T):
    """periodic z (config C2's loop, periodic_flow_past_sphere.py:95-183) on the r-slab stepper: ghost columns live
    inside every row, so the wrap-around is local to a rank; with one rank the stepper must reproduce
    RigidFlowStepper(periodic=True) on the same solve path (real FFT of the inner columns + tridiagonal r)."""
    from pyaxisymflow_b200.rowslab import RowSlabRigidFlowStepper
    from pyaxisymflow_b200.timestep import RigidFlowStepper

    torch = T
    nz, nr = 512 + 4, 128
    kw = dict(periodic=True, r_sph=0.075, Z_cm=0.85)
    a = RowSlabRigidFlowStepper(nz, grid_size_r=nr, **kw)
    b = RigidFlowStepper(nz, grid_size_r=nr, basis="analytic", r_method="tridiagonal", z_method="fft", **kw)
    assert a.solver.periodic and b.solver.plan.z_fft == 2
    a.seed_vorticity()
    b.seed_vorticity()
    assert torch.equal(a.gather_vorticity(), b.vorticity)
    assert torch.equal(a.L.owned(a.char_func), b.char_func)
    a.step(5)
    b.step(5)
    torch.cuda.synchronize()
    err = ((a.gather_vorticity() - b.vorticity).abs().max() / b.vorticity.abs().max()).item()
    assert err < 1e-10, err
    sa, sb = a.scalars(), b.scalars()
    assert abs(sa["t"] - sb["t"]) <= 1e-14 * sb["t"] and sa["iterations"] == sb["iterations"] == 5
