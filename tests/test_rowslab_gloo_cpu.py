"""Host-side logic of the r-slab (row) multi-GPU path on CPU: world_size 2 and 4 over gloo.

The stepper's kernels are injected (``ops``): here they are the ORACLE's restatements of the reference kernels
(oracle/axisym_oracle.py, test infrastructure), each acting on a rank's row block as if it were a whole field --
exactly what the CUDA kernels do on the GPU box.  The same oracle sequence on the undivided field is the
reference.  What this pins: layout arithmetic, the exchange schedule, and the halo-validity argument in
pyaxisymflow_b200/rowslab.py (every owned value must come out as on one domain: the stencil phase bit for bit,
the partitioned r solve to rounding)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from pyaxisymflow_b200 import fd  # noqa: E402
from pyaxisymflow_b200.rowslab import RowSlabComm, RowSlabLayout, RowSlabRigidFlowStepper  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _host_dct(dst, src, tables, inverse):
    n = src.shape[1]
    V = fd.axial_natural_block("neumann", n, n, np.arange(n))                        # orthonormal cosine basis
    V = V if torch.is_tensor(V) else torch.from_numpy(V)
    ck = torch.full((n,), np.sqrt(2.0 / n), dtype=torch.float64)
    ck[0] = np.sqrt(1.0 / n)
    dst.copy_((src / ck) @ V.T if inverse else (src @ V) * ck)


def _host_rfft(dst, src, tables, inverse):
    """periodic z: the real FFT of a row block in the half-complex layout of csrc/pfft.cu (numpy.fft stands in)"""
    if inverse:
        n = dst.shape[1]
        M = n // 2
        spec = src.numpy()
        Y = np.zeros((spec.shape[0], M + 1), complex)
        Y.real, Y.imag[:, 1:M] = spec[:, :M + 1], spec[:, M + 1:n]
        dst.copy_(torch.from_numpy(np.fft.irfft(Y, n=n, axis=1)))
    else:
        n = src.shape[1]
        M = n // 2
        X = np.fft.rfft(src.numpy(), axis=1)
        out = np.zeros(tuple(dst.shape))
        out[:, :M + 1], out[:, M + 1:n] = X.real, X.imag[:, 1:M]
        dst.copy_(torch.from_numpy(out))


class OracleOps:
    """one rigid-flow step's kernels on a (rows, nz) block, from the oracle (NumPy views of the torch storage)"""

    def __init__(self, dx, r1d, z1d, state, nu, lam, ju0, ju1):
        import oracle.axisym_oracle as orc

        self.o, self.dx, self.st, self.nu, self.lam, self.ju0, self.ju1 = orc, dx, state, nu, lam, ju0, ju1
        self.Z, self.R = np.meshgrid(z1d.numpy(), r1d.numpy())

    def scalars(self, phase, sc):
        U0, T_ramp, ur_ramp, dt_lim, cfl_dx = sc
        st = self.st
        if phase == 0:
            t = st[0].item()
            st[4] = U0 * (np.sin(0.5 * np.pi * t / T_ramp) if t < T_ramp else 1.0)
            st[5] = 0.0
            st[2] = 0.0
            st[3] = 0.0
        elif phase == 1:
            st[1] = min(dt_lim, cfl_dx / (st[2].item() + 2.220446049250313e-16))
        else:
            st[0] += st[1]
            st[6] += 1.0
            st[7] = st[3]

    def kill_z(self, w):
        self.o.kill_boundary_vorticity_sine_z(w.numpy(), self.Z, 3, self.dx)

    def kill_r(self, w, parts):
        a = w.numpy()
        full = a.copy()
        self.o.kill_boundary_vorticity_sine_r(full, self.R, 3, self.dx)
        if parts & 1:
            a[-3:] = full[-3:]
        if parts & 2:
            a[0] = 0.0

    def velocity(self, u_z, u_r, psi):
        uz, ur = u_z.numpy(), u_r.numpy()
        self.o.compute_velocity_from_psi(uz, ur, psi.numpy(), self.R, self.dx)
        uz += self.st[4].item()
        ur += self.st[5].item()
        m = (np.abs(uz) + np.abs(ur))[self.ju0:self.ju1].max()
        self.st[2] = max(self.st[2].item(), m)

    def penalise(self, u_z, u_r, w, uzu, uru, chi):
        dt = self.st[1].item()
        self.o.brinkmann_penalize(self.lam, dt, chi.numpy(), 0.0, 0.0, uzu.numpy(), uru.numpy(), u_z.numpy(),
                                  u_r.numpy())
        curl = np.zeros_like(w.numpy())
        self.o.compute_vorticity_from_velocity(curl, u_z.numpy() - uzu.numpy(), u_r.numpy() - uru.numpy(), self.dx)
        w.numpy()[1:-1, 1:-1] += curl[1:-1, 1:-1]
        self.st[3] += (self.R * chi.numpy() * u_z.numpy())[self.ju0:self.ju1].sum()

    def advect(self, w2, w, u_z, u_r):
        a = w.numpy().copy()
        self.o.advect_vorticity_via_eno3(a, u_z.numpy(), u_r.numpy(), self.st[1].item(), self.dx, use_c=False)
        w2.numpy()[...] = a

    def diffuse(self, w, w2, tmp):
        a = w2.numpy().copy()
        self.o.diffusion_RK2(a, tmp.numpy(), self.R, self.nu, self.st[1].item(), self.dx)
        w.numpy()[...] = a

    def heaviside_sphere(self, chi, Z_cm, R_cm, r_sph):
        phi = r_sph - np.sqrt((self.Z - Z_cm) ** 2 + (self.R - R_cm) ** 2)
        self.o.smooth_Heaviside(chi.numpy(), phi, self.dx * 2 ** 0.5)

    # periodic z: the ghost refresh and the two RK2 stages as separate calls (kernels/diffusion_RK2.py:48-89)
    def ghost(self, f, ghost):
        self.o.periodic_ghost_comm(f.numpy(), ghost)

    def diffuse_stage1(self, tmp, w2):
        a, t = w2.numpy(), tmp.numpy()
        t[...] = a
        t[1:-1, 1:-1] += 0.5 * self.nu * self.st[1].item() * self.o._diffusion_operator(a, self.R, self.dx)

    def diffuse_stage2(self, w, w2, tmp):
        a, out = w2.numpy(), w.numpy()
        out[...] = a
        out[1:-1, 1:-1] += self.nu * self.st[1].item() * self.o._diffusion_operator(tmp.numpy(), self.R, self.dx)


def _single_domain(nr, nz, steps, seed_field, kw, periodic=False):
    """the same oracle sequence on the undivided field (what one GPU computes)"""
    if periodic:
        return _single_domain_periodic(nr, nz, steps, seed_field, kw)
    dx = 1.0 / nz
    st = torch.zeros(8, dtype=torch.float64)
    z1d = torch.from_numpy(np.linspace(dx / 2, 1 - dx / 2, nz))
    r1d = torch.from_numpy(np.linspace(dx / 2, nr * dx - dx / 2, nr))
    nu = kw["U_0"] * 2 * kw["r_sph"] / kw["Re"]
    ops = OracleOps(dx, r1d, z1d, st, nu, kw["brink_lam"], 0, nr)
    fac = fd.build_factors("stokes", "homogenous_neumann_along_z_and_r", nr, nz, dx, "analytic",
                           r_method="tridiagonal", z_method="fft")
    f = lambda: torch.zeros((nr, nz), dtype=torch.float64)  # noqa: E731
    w, psi, uz, ur, uzu, uru, chi, tmp, w2 = seed_field.clone(), f(), f(), f(), f(), f(), f(), f(), f()
    ops.heaviside_sphere(chi, kw["Z_cm"], 0.0, kw["r_sph"])
    sc = (kw["U_0"], 20 * kw["r_sph"] / kw["U_0"], 0.0, 0.9 * dx ** 2 / 4 / nu, kw["CFL"] * dx)
    for _ in range(steps):
        ops.scalars(0, sc)
        ops.kill_z(w)
        ops.kill_r(w, 3)
        psi.copy_(torch.from_numpy(fd.apply_factors_host(fac, w.numpy())))
        ops.velocity(uzu, uru, psi)
        ops.scalars(1, sc)
        ops.penalise(uz, ur, w, uzu, uru, chi)
        ops.advect(w2, w, uz, ur)
        ops.diffuse(w, w2, tmp)
        ops.scalars(2, sc)
    return w, st


def _single_domain_periodic(nr, nz, steps, seed_field, kw, gh=2):
    """RigidFlowStepper(periodic=True)'s sequence (periodic_flow_past_sphere.py:95-183) with the oracle's kernels"""
    dx = 1.0 / nz
    st = torch.zeros(8, dtype=torch.float64)
    z1d = torch.from_numpy(np.linspace(dx / 2, 1 - dx / 2, nz))
    r1d = torch.from_numpy(np.linspace(dx / 2, nr * dx - dx / 2, nr))
    nu = kw["U_0"] * 2 * kw["r_sph"] / kw["Re"]
    ops = OracleOps(dx, r1d, z1d, st, nu, kw["brink_lam"], 0, nr)
    fac = fd.build_factors("stokes", "homogenous_neumann_along_r_and_periodic_along_z", nr, nz - 2 * gh, dx, "analytic",
                           r_method="tridiagonal", z_method="fft")
    f = lambda: torch.zeros((nr, nz), dtype=torch.float64)  # noqa: E731
    w, psi, uz, ur, uzu, uru, chi, tmp, w2 = seed_field.clone(), f(), f(), f(), f(), f(), f(), f(), f()
    ops.heaviside_sphere(chi, kw["Z_cm"], 0.0, kw["r_sph"])
    ops.ghost(chi, gh)
    sc = (kw["U_0"], 20 * kw["r_sph"] / kw["U_0"], 5e-2, 0.9 * dx ** 2 / 4 / nu, kw["CFL"] * dx)
    for _ in range(steps):
        ops.scalars(0, sc)
        ops.kill_r(w, 3)
        psi[:, gh:nz - gh] = torch.from_numpy(fd.apply_factors_host(fac, w[:, gh:nz - gh].numpy()))
        ops.ghost(psi, gh)
        ops.velocity(uzu, uru, psi)
        ops.scalars(1, sc)
        ops.ghost(uru, gh)
        ops.ghost(uzu, gh)
        ops.penalise(uz, ur, w, uzu, uru, chi)
        ops.advect(w2, w, uz, ur)
        ops.ghost(w2, gh)
        ops.diffuse_stage1(tmp, w2)
        ops.ghost(tmp, gh)
        ops.diffuse_stage2(w, w2, tmp)
        ops.scalars(2, sc)
    return w, st


def _worker(rank, world, port, nr, nz, steps, q, periodic=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = RowSlabLayout(nr, nz, world, rank)
        comm = RowSlabComm(L)
        rng = np.random.default_rng(11)
        full = torch.from_numpy(rng.standard_normal((nr, nz)))
        # ---- row-halo exchange: owned data only in, halos must come out equal to the global field
        f = torch.zeros((L.nrs, nz), dtype=torch.float64)
        L.owned(f).copy_(full[L.r_begin:L.r_begin + L.nrl])
        g2 = 3.0 * f
        comm.exchange([f, g2], 2)
        want = L.scatter_global(full)
        assert torch.equal(f, want), "row halo exchange (width 2)"
        assert torch.equal(g2, 3.0 * want)
        f1 = torch.zeros((L.nrs, nz), dtype=torch.float64)
        L.owned(f1).copy_(full[L.r_begin:L.r_begin + L.nrl])
        comm.exchange([f1], 1)
        w1 = want.clone()
        w1[0] = 0
        w1[-1] = 0
        assert torch.equal(f1, w1), "row halo exchange (width 1)"
        # ---- the whole step with the oracle's kernels on every rank's block vs the undivided field
        dx = 1.0 / nz
        kw = dict(U_0=1.0, r_sph=0.1, Re=100.0, brink_lam=1e4, Z_cm=0.4, CFL=0.1)
        state = torch.zeros(8, dtype=torch.float64)
        z1d = torch.from_numpy(np.linspace(dx / 2, 1 - dx / 2, nz))
        r_full = np.linspace(dx / 2, nr * dx - dx / 2, nr)
        r_blk = torch.from_numpy(r_full[L.g0:L.g0 + L.nv].copy())
        nu = kw["U_0"] * 2 * kw["r_sph"] / kw["Re"]
        ops = OracleOps(dx, r_blk, z1d, state, nu, kw["brink_lam"], L.ju0, L.ju1)
        if periodic:
            fac = fd.build_factors("stokes", "homogenous_neumann_along_r_and_periodic_along_z", nr, nz - 4, dx, "analytic",
                                   r_method="tridiagonal", z_method="fft")
        else:
            fac = fd.build_factors("stokes", "homogenous_neumann_along_z_and_r", nr, nz, dx, "analytic",
                                   r_method="tridiagonal", z_method="fft")
        s = RowSlabRigidFlowStepper(nz, grid_size_r=nr, device="cpu", ops=ops, factors=fac,
                                    dct=_host_rfft if periodic else _host_dct, host_tridiagonal=True, Z_cm=kw["Z_cm"],
                                    brink_lam=kw["brink_lam"], periodic=periodic)
        s.state = state
        zz, rr = np.meshgrid(z1d.numpy(), r_full)
        seed = torch.from_numpy(rng.standard_normal((nr, nz)) * np.exp(-((zz - 0.5) ** 2 + rr ** 2) / 0.02))
        s.vorticity.copy_(L.scatter_global(seed))
        s.step(steps)
        ref_w, ref_st = _single_domain(nr, nz, steps, seed, kw, periodic)
        got = s.gather_vorticity()
        err = (got - ref_w).abs().max().item() / ref_w.abs().max().item()
        assert err < 1e-10, f"r-slab step differs from the undivided field by {err:.2e}"
        assert abs(state[1].item() - ref_st[1].item()) <= 1e-12 * ref_st[1].item(), "dt (CFL all-reduce)"
        assert abs(state[0].item() - ref_st[0].item()) <= 1e-12 * ref_st[0].item(), "t"
        drag = comm.allreduce(state[7:8].clone(), "sum").item()
        assert abs(drag - ref_st[7].item()) <= 1e-9 * max(1.0, abs(ref_st[7].item())), "drag sum over owned rows"
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic", [(2, False), (4, False), (2, True), (4, True)])
def test_rowslab_step_over_gloo(world, periodic):
    """periodic: config C2's loop -- ghost columns inside every row, real FFT of the inner 64 columns, two RK2 stages with
    a ghost refresh in between; the wrap-around must never need a neighbour"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    nz = 64 + 4 if periodic else 64
    procs = [ctx.Process(target=_worker, args=(r, world, port, 32, nz, 4, q, periodic)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def test_rowslab_layout_arithmetic():
    L = RowSlabLayout(64, 128, 4, 0)
    assert (L.nrl, L.nrs, L.v0, L.v1, L.nv, L.ju0, L.ju1, L.g0) == (16, 20, 2, 20, 18, 0, 16, 0)
    assert L.lower is None and L.upper == 1
    L = RowSlabLayout(64, 128, 4, 2)
    assert (L.v0, L.v1, L.nv, L.ju0, L.ju1, L.g0, L.r_begin) == (0, 20, 20, 2, 18, 30, 32)
    L = RowSlabLayout(64, 128, 4, 3)
    assert (L.v0, L.v1, L.nv, L.ju0, L.ju1, L.g0) == (0, 18, 18, 2, 18, 46) and L.upper is None
    L = RowSlabLayout(64, 128, 1, 0)
    assert (L.v0, L.v1, L.ju0, L.ju1, L.g0) == (2, 66, 0, 64, 0)
    with pytest.raises(ValueError):
        RowSlabLayout(16, 128, 4, 0)
