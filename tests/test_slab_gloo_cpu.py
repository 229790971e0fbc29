"""Host-side logic of the z-slab (N > 1) path on CPU: world_size 2 and 4 over gloo.
Checks the decomposition arithmetic, the halo exchange, the two all-to-all transposes and the
distributed fast-diagonalisation data flow (GEMMs replaced by torch.matmul here -- the point
is the plumbing; the CUDA GEMM has its own parity tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyaxisymflow_b200 import fd
from pyaxisymflow_b200.slab import SlabComm, SlabFdSolver, SlabLayout


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torch_gemm(C, A, B, scale_m=None, scale_n=None, c0=0.0, c1=1.0):
    out = A @ B
    if scale_m is not None:
        out = out * (1.0 / (c0 + c1 * (scale_n[None, :] + scale_m[:, None])))
    C.copy_(out)


def _torch_fold(x, n, inverse):
    x.copy_(torch.from_numpy(fd.fold_host(x.numpy(), n, inverse)))


def _host_dct(dst, src, tables, inverse):
    n = src.shape[1]
    V = fd.axial_natural_block("neumann", n, n, np.arange(n))          # orthonormal cosine basis
    ck = torch.full((n,), np.sqrt(2.0 / n), dtype=torch.float64)
    ck[0] = np.sqrt(1.0 / n)
    dst.copy_((src / ck) @ V.T if inverse else (src @ V) * ck)


class _HostTridiagonal:
    def __init__(self, L, f):
        self.t = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in f["tri"].items()}
        self.lam = f["lam_z"].numpy()[L.z_begin:L.z_begin + L.nzl]
        self.c0, self.c1 = f["c0"], f["c1"]

    def __call__(self, x):
        t = self.t
        x.copy_(torch.from_numpy(fd.thomas_host(x.numpy(), t["sub"], t["diag"], t["sup"], self.lam, t["scale"],
                                                self.c0, self.c1)))


def _worker(rank, world, port, nr, nz, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = SlabLayout(nr, nz, world, rank)
        comm = SlabComm(L)
        rng = np.random.default_rng(7)
        full = torch.from_numpy(rng.standard_normal((nr, nz)))
        # ---- halo exchange: start from owned data only, halos must come out equal to the global field
        f = torch.zeros((nr, L.nzs), dtype=torch.float64)
        L.owned(f).copy_(full[:, L.z_begin:L.z_begin + L.nzl])
        g2 = 3.0 * f
        comm.exchange([f, g2], 2)
        want = L.scatter_global(full)
        assert torch.equal(f, want), "halo exchange (width 2)"
        assert torch.equal(g2, 3.0 * want)
        f1 = torch.zeros((nr, L.nzs), dtype=torch.float64)
        L.owned(f1).copy_(full[:, L.z_begin:L.z_begin + L.nzl])
        comm.exchange([f1], 1)
        w1 = want.clone()
        w1[:, 0] = 0
        w1[:, -1] = 0
        assert torch.equal(f1, w1), "halo exchange (width 1)"
        # ---- transposes
        slab = full[:, L.z_begin:L.z_begin + L.nzl].contiguous()
        rows = torch.empty((L.nrl, nz), dtype=torch.float64)
        comm.slab_to_rows(slab, rows)
        assert torch.equal(rows, full[L.r_begin:L.r_begin + L.nrl]), "slab -> rows"
        back = torch.empty_like(slab)
        comm.rows_to_slab(rows, back)
        assert torch.equal(back, slab), "rows -> slab"
        # ---- distributed solve data flow against the single-process factor application
        dx = 1.0 / nz
        for split in (0, 1):          # dense and parity-split z transforms
            fac = fd.build_factors("stokes", "homogenous_neumann_along_z_and_r", nr, nz, dx, "analytic", split=split)
            if split:
                fac["zsplit"] = fd.axial_split_plan("neumann", 1.0, nz, dx, 1)   # force one level on this small grid
                fac["lam_z"] = torch.from_numpy(fac["zsplit"]["lam_z"])
                fac["Rz"] = fac["Rzb"] = None
            ref = torch.from_numpy(fd.apply_factors_host(fac, full.numpy()))
            solver = SlabFdSolver(L, comm, fac, gemm=_torch_gemm, fold=_torch_fold)
            rhs_slab = L.scatter_global(full)
            psi_slab = torch.zeros_like(rhs_slab)
            solver.solve(psi_slab, rhs_slab)
            got = L.owned(psi_slab)
            want = ref[:, L.z_begin:L.z_begin + L.nzl]
            err = (got - want).abs().max().item() / ref.abs().max().item()
            assert err < 1e-12, f"distributed solve (split={split}) differs by {err:.2e}"
        # ---- the same with the direct r solve: GEMM z transforms (dense / split) and cosine transforms
        for z_method, split in (("gemm", 0), ("gemm", 1), ("fft", 0)):
            fac = fd.build_factors("stokes", "homogenous_neumann_along_z_and_r", nr, nz, dx, "analytic", split=split,
                                   r_method="tridiagonal", z_method=z_method)
            ref = torch.from_numpy(fd.apply_factors_host(fac, full.numpy()))
            solver = SlabFdSolver(L, comm, fac, gemm=_torch_gemm, fold=_torch_fold, dct=_host_dct,
                                  tri=_HostTridiagonal(L, fac))
            psi_slab = torch.zeros_like(rhs_slab)
            solver.solve(psi_slab, rhs_slab)
            err = (L.owned(psi_slab) - ref[:, L.z_begin:L.z_begin + L.nzl]).abs().max().item() / ref.abs().max().item()
            assert err < 1e-12, f"distributed tridiagonal solve ({z_method}, split={split}) differs by {err:.2e}"
        # ---- r solve partitioned over the ranks (r-slab rows all the way, 2 transposes per solve)
        from pyaxisymflow_b200.slab import PartitionedTridiagonal

        fac = fd.build_factors("stokes", "homogenous_neumann_along_z_and_r", nr, nz, dx, "analytic",
                               r_method="tridiagonal", z_method="fft")
        ref = torch.from_numpy(fd.apply_factors_host(fac, full.numpy()))
        part = PartitionedTridiagonal(L, fac, None, host=True)
        solver = SlabFdSolver(L, comm, fac, dct=_host_dct, tri=_HostTridiagonal(L, fac), part=part)
        psi_slab = torch.zeros_like(rhs_slab)
        solver.solve(psi_slab, rhs_slab)
        err = (L.owned(psi_slab) - ref[:, L.z_begin:L.z_begin + L.nzl]).abs().max().item() / ref.abs().max().item()
        assert err < 1e-12, f"partitioned r solve differs by {err:.2e}"
        # ---- reductions
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        assert comm.allreduce(t.clone(), "max").item() == world
        assert comm.allreduce(t.clone(), "sum").item() == world * (world + 1) / 2
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_plumbing_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 32, 64, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def test_layout_arithmetic():
    L = SlabLayout(64, 128, 4, 0)
    assert (L.nzl, L.nrl, L.nzs, L.kz0, L.ku0, L.ku1) == (32, 16, 36, -2, 2, 34)
    assert L.left is None and L.right == 1
    L = SlabLayout(64, 128, 4, 3)
    assert L.kz0 == 94 and L.right is None and L.left == 2
    g = L.grid(1 / 128)
    assert (g.nr, g.nz, g.ld, g.kz0, g.nz_global, g.ku0, g.ku1) == (64, 36, 36, 94, 128, 2, 34)
    assert SlabLayout(64, 128, 4, 3, periodic=True).right == 0
    with pytest.raises(ValueError):
        SlabLayout(64, 130, 4, 0)
    with pytest.raises(ValueError):
        SlabLayout(64, 16, 4, 0)
