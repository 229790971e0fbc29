"""r-slab (row) multi-GPU execution of the rigid-flow timestep -- no transposes.

One process per GPU.  Rank ``p`` of ``P`` owns the ROWS ``[p Nr/P, (p+1) Nr/P)`` of every ``(Nr, Nz)`` field and
stores them with ``H = 2`` halo rows on either side, ``(H + Nr/P + H, Nz)``; rows are contiguous, so a halo is one
``2 Nz``-double run per side and travels as a plain peer-memory read (``axb_row_halo_get``, NVLink) after a
device-side barrier.  Why rows: the z transforms of the fast-diagonalisation solve (DCT-II / DCT-III of whole rows,
``csrc/zfft.cu``) are then local, and the r solve is the partition (SPIKE) method that already lives in this layout
(:class:`pyaxisymflow_b200.slab.PartitionedTridiagonal`: own-block sweeps + a 2P-row interface gather + one
correction pass).  The z-slab flow of ``slab.py`` needs two all-to-all transposes per solve for the same thing;
here the only traffic per step is 3 halo rounds (4 fields x 2 rows), the 2P interface rows and one 8-byte
MAX all-reduce.

The stencil kernels are used UNCHANGED: a rank's block ``[r_begin - H, r_begin + Nr/P + H)`` (clipped to the domain)
is an ordinary field for them, with ``r1d`` = the block's own radii.  Their r-boundary formulas (axis reflection,
one-sided differences, untouched edge rows) then act on the block's first / last rows, which on an interior
interface are halo rows: whatever they compute there is overwritten by the next exchange before anything owned
depends on it --

    psi (owned, from the solve)            -> exchange w=2 -> valid on all block rows
    u = curl(psi)/r (centred, rows +-1)    -> valid on block rows [1, n-2]      (owned rows are [2, n-2))
    w += curl(pen(u) - u)   (rows +-1)     -> valid on owned rows; pen(u) valid on [1, n-2]
    ENO3 advection (rows +-2)              -> needs (w, u_r) on all block rows: exchange w=2 -> w2 valid on owned
    fused RK2 diffusion (rows +-2)         -> needs w2 on all block rows: exchange w=2 -> w valid on owned

-- and the true r boundaries (row 0 := 0 on the axis rank, the sine ramp on the r_max rank;
``kernels/kill_boundary_vorticity_sine.py:17-27``) are applied by the rank that holds them
(``axb_kill_boundary_vorticity_sine_r_parts``).  The fused reductions (CFL maximum in G-VEL, drag sum in G-PEN) count
owned rows only (``axb_grid_t.ju0, ju1``).  Every owned value is computed by the same kernel from the same inputs
as on one GPU, so the stencil phase is bit-identical to the single-GPU stepper; the partitioned r solve differs by
rounding.

The kernels are reached through an ``ops`` object so that the host logic (layout arithmetic, exchange schedule,
halo-validity argument above) runs on CPU in the gloo tests with the oracle's kernels injected
(tests/test_rowslab_gloo_cpu.py); the product default is :class:`CudaOps` -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .device import make_grid, ptr, stream_ptr
from .slab import PartitionedTridiagonal, _cuda_dct

_call = _lib.call
HALO = 2


class RowSlabLayout:
    """Pure index arithmetic of the row decomposition (no device, no communication)."""

    def __init__(self, nr, nz, world, rank, halo=HALO):
        if nr % world:
            raise ValueError(f"{nr} rows are not divisible by {world} ranks")
        if nr // world < 2 * halo + 1:
            raise ValueError("row slabs are thinner than the stencil halo")
        self.nr, self.nz, self.world, self.rank, self.halo = nr, nz, world, rank, halo
        self.nrl = nr // world                # owned rows
        self.nrs = self.nrl + 2 * halo        # stored rows (the same on every rank: peer offsets agree)
        self.r_begin = rank * self.nrl        # global index of the first owned row
        self.lower = rank - 1 if rank > 0 else None
        self.upper = rank + 1 if rank < world - 1 else None
        lo = halo if self.lower is not None else 0
        hi = halo if self.upper is not None else 0
        self.v0, self.v1 = halo - lo, halo + self.nrl + hi      # stored rows the kernels see ("the block")
        self.nv = self.v1 - self.v0
        self.ju0, self.ju1 = lo, lo + self.nrl                  # owned rows inside the block
        self.g0 = self.r_begin - lo                             # global row of block row 0

    def block(self, t):
        """the rows of a stored field that exist in the global domain (what the kernels are given)"""
        return t[self.v0:self.v1]

    def owned(self, t):
        return t[self.halo:self.halo + self.nrl]

    def scatter_global(self, full):
        """stored field (halos filled where they exist) cut out of a global (Nr, Nz) array"""
        out = torch.zeros((self.nrs, full.shape[1]), dtype=full.dtype, device=full.device)
        out[self.v0:self.v1] = full[self.g0:self.g0 + self.nv]
        return out


class RowSlabComm:
    """Row-halo exchange and scalar reductions on a RowSlabLayout."""

    def __init__(self, layout, group=None):
        self.L, self.group = layout, group
        self._peer_fields = {}          # data_ptr -> (symmetric-memory handle, peer base pointers)

    def symmetric_field(self, shape):
        """zero-initialised float64 field whose copies on all ranks are mapped into every process"""
        import torch.distributed._symmetric_memory as symm

        t = symm.empty(tuple(shape), dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
        t.zero_()
        h = symm.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
        if h.buffer_ptrs[self.L.rank] != t.data_ptr():
            raise RuntimeError("symmetric buffer does not start at the tensor's data pointer")
        self._peer_fields[t.data_ptr()] = (h, list(h.buffer_ptrs))
        return t

    def peer_info(self, t):
        return self._peer_fields[t.data_ptr()]

    # -- flag-based synchronisation (plain kernels: a whole step can be captured in a CUDA graph) -----------------
    def enable_flag_sync(self):
        """control block in peer-mapped memory + local epoch counters (include/axisym_b200.h, axb_peer_sync)"""
        import torch.distributed._symmetric_memory as symm

        L = self.L
        dev = torch.device("cuda", torch.cuda.current_device())
        self.ctl = symm.empty((160,), dtype=torch.int64, device=dev)
        self.ctl.zero_()
        h = symm.rendezvous(self.ctl, self.group if self.group is not None else dist.group.WORLD)
        self._ctl_handle = h
        self.ctl_ptrs = (ctypes.c_uint64 * L.world)(*h.buffer_ptrs)
        self.counters = torch.zeros(8, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)        # every rank's control block is zero before anyone signals
        self.flag_sync = True

    def sync(self):
        L = self.L
        _call("axb_peer_sync", L.world, L.rank, self.ctl_ptrs, ptr(self.counters), stream_ptr())

    def check(self):
        """raise if a device-side wait gave up (host synchronisation: diagnostics only)"""
        if getattr(self, "flag_sync", False) and int(self.counters[7].item()) != 0:
            raise _lib.AxbError("a rank-synchronisation wait timed out on the device (peer flags never arrived)")

    def _exchange_peer(self, fields, width):
        L = self.L
        n = len(fields)
        arr = ctypes.c_uint64 * n
        src, lo, up = arr(), arr(), arr()
        h = None
        for i, f in enumerate(fields):
            h, ptrs = self._peer_fields[f.data_ptr()]
            src[i] = f.data_ptr()
            lo[i] = ptrs[L.lower] if L.lower is not None else 0
            up[i] = ptrs[L.upper] if L.upper is not None else 0
        # Pull, not push: a step's kernels also write their own block's halo rows (values nobody uses), so a
        # neighbour's store could be overwritten by a kernel still running here.  With the pull form ONE handshake
        # before it orders everything -- the neighbours have produced their edge rows, this rank's stray halo
        # writes are stream-ordered before its own pull, and a neighbour cannot overwrite the rows being read
        # before it has passed the next exchange's handshake, which this rank only answers after the pull.
        if getattr(self, "flag_sync", False):       # handshake with the two neighbours + pull, one launch
            cp = self.ctl_ptrs
            _call("axb_row_halo_exchange", n, src, lo, up, fields[0].stride(0), L.nz, L.nrl, L.halo, width,
                  ctypes.c_void_p(cp[L.rank]), ctypes.c_void_p(cp[L.lower]) if L.lower is not None else None,
                  ctypes.c_void_p(cp[L.upper]) if L.upper is not None else None, ptr(self.counters), stream_ptr())
            return
        h.barrier(channel=1)
        _call("axb_row_halo_get", n, src, lo, up, fields[0].stride(0), L.nz, L.nrl, L.halo, width, stream_ptr())

    def exchange(self, fields, width):
        """fill `width` halo rows of every stored field in `fields` from the two r-neighbours"""
        L = self.L
        if L.world == 1:
            return
        if fields and all(f.is_cuda and f.data_ptr() in self._peer_fields for f in fields):
            return self._exchange_peer(fields, width)
        H, n = L.halo, L.nrl
        ops = []
        for f in fields:                 # row runs are contiguous: no packing on either side
            if L.lower is not None:
                ops += [dist.P2POp(dist.isend, f[H:H + width], L.lower, self.group),
                        dist.P2POp(dist.irecv, f[H - width:H], L.lower, self.group)]
            if L.upper is not None:
                ops += [dist.P2POp(dist.isend, f[H + n - width:H + n], L.upper, self.group),
                        dist.P2POp(dist.irecv, f[H + n:H + n + width], L.upper, self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def allreduce(self, t, op="max"):
        L = self.L
        if L.world > 1:
            if op == "max" and t.numel() == 1 and t.is_cuda and getattr(self, "flag_sync", False):
                _call("axb_peer_allreduce_max", L.world, L.rank, self.ctl_ptrs, ptr(t), ptr(self.counters), stream_ptr())
                return t
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=self.group)
        return t


def _cuda_rfft(dst, src, tables, inverse):
    """periodic z: real FFT rows (csrc/pfft.cu); dst / src are the half-complex spectral block and the inner columns"""
    if inverse:
        rows, n = dst.shape
        _call("axb_irfft_rows", rows, n, ptr(src), src.stride(0), ptr(dst), dst.stride(0), ptr(tables), 2.0 / n,
              stream_ptr())
    else:
        rows, n = src.shape
        _call("axb_rfft_rows", rows, n, ptr(src), src.stride(0), ptr(dst), dst.stride(0), dst.shape[1], ptr(tables), 1.0,
              stream_ptr())


class RowSlabFdSolver:
    """Fast-diagonalisation solve with the rows split over the ranks: z transform of the owned rows (DCT-II, or the
    real FFT of the inner columns when z is periodic), partitioned r solve, inverse transform -- nothing but the 2P
    interface rows leaves the GPU.  ``dct`` (the z transform, ``(dst, src, tables, inverse)``) is injectable for the
    CPU tests."""

    def __init__(self, layout, factors, part, dct=None, ghost=0):
        if factors.get("zfft") is None or factors.get("tri") is None:
            raise _lib.AxbError("the r-slab solve needs an FFT z path (cosine transforms: power-of-two Nz; periodic z: "
                                "even inner width with small prime factors) and the tridiagonal r path; use "
                                "pyaxisymflow_b200.slab.SlabRigidFlowStepper otherwise")
        self.L, self.f, self.part = layout, factors, part
        self.periodic = factors["zfft"]["family"] == "periodic"
        self.ghost = int(ghost) if self.periodic else 0
        self.dct = dct or (_cuda_rfft if self.periodic else _cuda_dct)
        width = factors["zfft"]["nz_spec"]
        self.spec = torch.empty((layout.nrl, width), dtype=torch.float64, device=factors["lam_z"].device)

    def solve(self, psi_owned, rhs_owned, mark=None):
        tables = self.f["zfft"]["tables"]
        mark = mark or (lambda name: None)
        g = self.ghost
        if g:                                        # periodic z: the solve lives on the inner columns
            nz = psi_owned.shape[1]
            psi_owned, rhs_owned = psi_owned[:, g:nz - g], rhs_owned[:, g:nz - g]
        self.dct(self.spec, rhs_owned, tables, False)
        mark("solve_dct2")
        self.part(self.spec)
        mark("solve_r_partitioned")
        self.dct(psi_owned, self.spec, tables, True)
        mark("solve_dct3")


class CudaOps:
    """The kernels of one rigid-flow step on a row block, through the C ABI (libaxisym_b200)."""

    def __init__(self, layout, dx, r1d_block, z1d, state, nu, brink_lam):
        L = self.L = layout
        self.dx, self.r1d, self.z1d, self.st, self.nu, self.lam = dx, r1d_block, z1d, state, nu, brink_lam
        self.grid = make_grid(L.nv, L.nz, L.nz, dx, rows=(L.ju0, L.ju1))
        self.g = ctypes.byref(self.grid)

    def sp(self, i):
        return ctypes.c_void_p(self.st.data_ptr() + 8 * i)

    def scalars(self, phase, sc):
        _call("axb_rigid_flow_scalars", phase, ptr(self.st), *sc, stream_ptr())

    def kill_z(self, w):
        _call("axb_kill_boundary_vorticity_sine_z", self.g, ptr(w), ptr(self.z1d), 3, stream_ptr())

    def kill_r(self, w, parts):
        _call("axb_kill_boundary_vorticity_sine_r_parts", self.g, ptr(w), ptr(self.r1d), 3, parts, stream_ptr())

    def velocity(self, u_z, u_r, psi):
        _call("axb_velocity_from_psi", self.g, ptr(u_z), ptr(u_r), ptr(psi), ptr(self.r1d), 0.0, 0.0, self.sp(4),
              self.sp(2), stream_ptr())

    def penalise(self, u_z, u_r, w, uzu, uru, chi):
        _call("axb_penalise_update_vorticity", self.g, ptr(u_z), ptr(u_r), ptr(w), ptr(uzu), ptr(uru), ptr(chi),
              self.lam, 0.0, self.sp(1), 0.0, 0.0, None, ptr(self.r1d), self.sp(3), stream_ptr())

    def advect(self, w2, w, u_z, u_r):
        _call("axb_advect_vorticity_eno3", self.g, ptr(w2), ptr(w), ptr(u_z), ptr(u_r), 0.0, self.sp(1), stream_ptr())

    def diffuse(self, w, w2, tmp):
        _call("axb_diffusion_rk2_fused", self.g, ptr(w), ptr(w2), ptr(tmp), ptr(self.r1d), self.nu, 0.0, self.sp(1),
              stream_ptr())

    # periodic z (periodic_flow_past_sphere.py:95-183): ghost columns are refreshed inside every row -- local to a rank
    def ghost(self, f, ghost):
        _call("axb_periodic_ghost_comm", self.g, ptr(f), int(ghost), 0.0, 0.0, stream_ptr())

    def diffuse_stage1(self, tmp, w2):
        _call("axb_diffusion_rk2_stage1", self.g, ptr(tmp), ptr(w2), ptr(self.r1d), self.nu, 0.0, self.sp(1), stream_ptr())

    def diffuse_stage2(self, w, w2, tmp):
        _call("axb_diffusion_rk2_stage2", self.g, ptr(w), ptr(w2), ptr(tmp), ptr(self.r1d), self.nu, 0.0, self.sp(1),
              stream_ptr())

    def heaviside_sphere(self, chi, Z_cm, R_cm, r_sph):
        _call("axb_smooth_heaviside_sphere", self.g, ptr(chi), None, ptr(self.z1d), ptr(self.r1d), float(Z_cm),
              float(R_cm), float(r_sph), float(self.dx * 2 ** 0.5), stream_ptr())


class RowSlabRigidFlowStepper:
    """r-slab version of :class:`pyaxisymflow_b200.timestep.RigidFlowStepper` (same physics, same kernels, same launch
    order -- see the module docstring for the exchanges and why the halos are valid).  ``periodic=True`` is the loop of
    ``periodic_flow_past_sphere.py:95-183`` (config C2): z is periodic with ``ghost_size`` ghost columns, which live
    inside every row -- the wrap-around never leaves a rank, the real FFT of the inner columns is as local as the cosine
    transforms, and the row halos travel exactly as in the non-periodic case."""

    def __init__(self, grid_size_z, grid_size_r=None, domain_AR=0.5, Re=100.0, U_0=1.0, r_sph=0.1, Z_cm=0.25,
                 R_cm=0.0, brink_lam=1e12, CFL=0.1, basis="analytic", group=None, device="cuda", ops=None,
                 factors=None, dct=None, host_tridiagonal=False, use_graph=False, periodic=False, ghost_size=2):
        from .fd import build_factors

        cuda = device != "cpu"
        if cuda and not torch.cuda.is_available():
            raise _lib.AxbError("RowSlabRigidFlowStepper needs CUDA devices (no CPU fallback)")
        if not cuda and ops is None:
            raise _lib.AxbError("a CPU RowSlabRigidFlowStepper only exists for the host-logic tests (inject ops)")
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nz = int(grid_size_z)
        self.nr = int(grid_size_r) if grid_size_r is not None else int(domain_AR * grid_size_z)
        self.dx = dx = 1.0 / self.nz
        self.L = L = RowSlabLayout(self.nr, self.nz, world, rank)
        self.comm = RowSlabComm(L, group)
        self.periodic, self.ghost = bool(periodic), int(ghost_size)
        self.U_0, self.r_sph, self.brink_lam, self.CFL = U_0, r_sph, brink_lam, CFL
        self.Z_cm, self.R_cm = Z_cm, R_cm
        self.nu = U_0 * 2 * r_sph / Re
        self.T_ramp = 20 * r_sph / U_0
        self.dt_diff_limit = 0.9 * dx ** 2 / 4 / self.nu
        self.z1d = torch.from_numpy(np.linspace(0 + dx / 2, 1 - dx / 2, self.nz)).to(device)
        r_full = np.linspace(0 + dx / 2, self.nr * dx - dx / 2, self.nr)
        self.r1d = torch.from_numpy(r_full[L.g0:L.g0 + L.nv].copy()).to(device)     # radii of the block rows

        def field():
            return torch.zeros((L.nrs, self.nz), dtype=torch.float64, device=device)

        # the fields whose halo rows travel live in peer-mapped memory when the box allows it
        xfield, self.peer_halos = field, False
        if cuda and world > 1 and not os.environ.get("AXB_SLAB_NCCL"):
            try:
                probe = self.comm.symmetric_field((L.nrs, self.nz))
                xfield, self.peer_halos = (lambda: self.comm.symmetric_field((L.nrs, self.nz))), True
                self.vorticity = probe
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    print(f"[pyaxisymflow_b200] peer-memory halos unavailable ({e!r}); using NCCL send/recv", flush=True)
        if not self.peer_halos:
            self.vorticity = field()
        if self.peer_halos and not os.environ.get("AXB_SLAB_TORCH_BARRIER"):
            self.comm.enable_flag_sync()
        self.psi, self.u_r, self._w2, self.char_func = xfield(), xfield(), xfield(), xfield()
        self.u_z, self.u_z_upen, self.u_r_upen, self._tmp = field(), field(), field(), field()
        self.state = torch.zeros(8, dtype=torch.float64, device=device)
        self.ops = ops if ops is not None else CudaOps(L, dx, self.r1d, self.z1d, self.state, self.nu, brink_lam)
        self.ops.heaviside_sphere(L.block(self.char_func), Z_cm, R_cm, r_sph)       # analytic: halo rows included
        if self.periodic:
            self.ops.ghost(L.block(self.char_func), self.ghost)
            self.factors = factors if factors is not None else build_factors(
                "stokes", "homogenous_neumann_along_r_and_periodic_along_z", self.nr, self.nz - 2 * self.ghost, dx, basis,
                device=device, r_method="tridiagonal", z_method="fft")
        else:
            self.factors = factors if factors is not None else build_factors(
                "stokes", "homogenous_neumann_along_z_and_r", self.nr, self.nz, dx, basis, device=device,
                r_method="tridiagonal", z_method="auto")
        # the r solve sees the spectral width (= Nz for the cosine transforms, the padded half-complex width when periodic)
        zf = self.factors.get("zfft")
        Ls = L if (zf is None or zf["nz_spec"] == self.nz) else RowSlabLayout(self.nr, zf["nz_spec"], world, rank)
        peer_buf = None
        if self.peer_halos:
            def peer_buf(shape):
                t = self.comm.symmetric_field(shape)
                h, ptrs = self.comm.peer_info(t)
                return t, h, ptrs
        flag_sync = getattr(self.comm, "flag_sync", False)
        self.part = PartitionedTridiagonal(Ls, self.factors, group, peer_ptrs=peer_buf, host=host_tridiagonal,
                                           sync=self.comm.sync if flag_sync else None)
        # a captured step needs every launch to be a plain kernel: the flag-based synchronisation, or one rank
        self._use_graph = bool(use_graph) and cuda and (world == 1 or flag_sync)
        self._graph, self._graphs3, self.graph_launches, self.launches_replayed = None, None, 0, 0
        self.solver = RowSlabFdSolver(L, self.factors, self.part, dct=dct, ghost=self.ghost)
        self._sc = (self.U_0, self.T_ramp, 5e-2 if self.periodic else 0.0, self.dt_diff_limit, self.CFL * self.dx)

    def seed_vorticity(self, seed=0, amplitude=1.0):
        """same global field as RigidFlowStepper.seed_vorticity, cut to this rank's rows"""
        L = self.L
        dev = self.vorticity.device
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        noise = torch.randn((self.nr, self.nz), dtype=torch.float64, device=dev, generator=gen)
        # the same radii as RigidFlowStepper's r1d (np.linspace; torch.linspace rounds differently when dx is no power of 2)
        rf = torch.from_numpy(np.linspace(self.dx / 2, self.nr * self.dx - self.dx / 2, self.nr)).to(dev)
        env = torch.exp(-((self.z1d[None, :] - 0.5) ** 2 + rf[:, None] ** 2) / 0.02)
        self.vorticity.copy_(L.scatter_global(amplitude * noise * env))

    # -- one step, enqueued on the current stream --------------------------------------------------------------
    def _enqueue(self, probe=None, mark=None, phases=(0, 1, 2)):
        """probe: (event, event) recorded around the solve; mark(name): called after every phase (phase_times);
        phases: 0 = everything before the solve, 1 = the solve, 2 = everything after it"""
        L, o, B = self.L, self.ops, self.L.block
        w, psi = self.vorticity, self.psi
        mark = mark or (lambda name: None)
        if 0 in phases:
            o.scalars(0, self._sc)
            if not self.periodic:
                o.kill_z(B(w))
            parts = (1 if L.upper is None else 0) | (2 if L.lower is None else 0)
            if parts:
                o.kill_r(B(w), parts)
            mark("boundaries")
        if probe is not None:
            probe[0].record()
        if 1 in phases:
            self.solver.solve(L.owned(psi), L.owned(w), mark)
        if probe is not None:
            probe[1].record()
        if 2 not in phases:
            return
        per, gh = self.periodic, self.ghost
        if per:
            o.ghost(B(psi), gh)                      # per row; the halo rows arrive from the neighbours with their ghosts
        self.comm.exchange([psi], 2)
        mark("halo_psi")
        o.velocity(B(self.u_z_upen), B(self.u_r_upen), B(psi))
        mark("velocity")
        self.comm.allreduce(self.state[2:3], "max")
        o.scalars(1, self._sc)
        mark("allreduce_cfl")
        if per:
            o.ghost(B(self.u_r_upen), gh)
            o.ghost(B(self.u_z_upen), gh)
        o.penalise(B(self.u_z), B(self.u_r), B(w), B(self.u_z_upen), B(self.u_r_upen), B(self.char_func))
        mark("penalise")
        self.comm.exchange([w, self.u_r], 2)
        mark("halo_w_ur")
        o.advect(B(self._w2), B(w), B(self.u_z), B(self.u_r))
        if per:
            o.ghost(B(self._w2), gh)
        mark("advect")
        self.comm.exchange([self._w2], 2)
        mark("halo_w2")
        if per:
            # two stages with a ghost refresh in between (RigidFlowStepper does the same); the intermediate field is valid
            # on the block rows [1, n-2], which is what the owned rows of stage 2 read
            o.diffuse_stage1(B(self._tmp), B(self._w2))
            o.ghost(B(self._tmp), gh)
            o.diffuse_stage2(B(w), B(self._w2), B(self._tmp))
        else:
            o.diffuse(B(w), B(self._w2), B(self._tmp))
        o.scalars(2, self._sc)
        mark("diffuse")

    def phase_times(self, steps=5):
        """per-phase milliseconds of one step on this rank (CUDA events between the phases, mean over `steps`
        steps; bench.py reports them beside the N > 1 lines)"""
        acc = {}
        for _ in range(steps):
            names, evs = [], [torch.cuda.Event(enable_timing=True)]
            evs[0].record()

            def mark(name):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                names.append(name)
                evs.append(e)

            self._enqueue(mark=mark)
            torch.cuda.synchronize()
            for i, n in enumerate(names):
                acc[n] = acc.get(n, 0.0) + evs[i].elapsed_time(evs[i + 1]) / steps
        return {k: round(v, 4) for k, v in acc.items()}

    def _capture(self, phases):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(graph, stream=side):
            self._enqueue(phases=phases)
        return graph, _lib.launch_count() - before

    def step(self, n=1):
        """advance n timesteps (asynchronous).  With ``use_graph`` the step -- kernels, halo handshakes, the CFL
        all-reduce -- is captured once and replayed; every rank replays the same graph, the device-side flags keep
        the ranks in step."""
        if self._use_graph:
            if self._graph is None:
                self._enqueue()                      # warm-up outside capture (sets kernel attributes)
                torch.cuda.synchronize()
                self._graph, self.graph_launches = self._capture((0, 1, 2))
                n -= 1
            for _ in range(n):
                self._graph.replay()
                self.launches_replayed += self.graph_launches
            return
        for _ in range(n):
            self._enqueue()

    def step_probed(self):
        """one step with CUDA events around the solve (bench.py's roofline leg); with ``use_graph`` three graphs"""
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        if not self._use_graph:
            self._enqueue(probe=ev)
            return ev
        if self._graphs3 is None:
            self._enqueue(probe=ev)                  # the first call is the warm-up: it IS the step, launched eagerly
            torch.cuda.synchronize()
            self._graphs3 = [self._capture((p,)) for p in (0, 1, 2)]
            return ev
        (g0, n0), (g1, n1), (g2, n2) = self._graphs3
        g0.replay()
        ev[0].record()
        g1.replay()
        ev[1].record()
        g2.replay()
        self.launches_replayed += n0 + n1 + n2
        return ev

    def close(self):
        """drop the captured graphs (before the process group goes away)"""
        self._graph, self._graphs3 = None, None

    def step_host(self, vorticity_rows_host, char_func_rows_host, out_rows_host):
        """End-to-end form for host-resident callers: this rank's owned rows (pinned ``(Nr/P, Nz)`` host arrays --
        contiguous slices of the global fields) in, one step, owned vorticity rows out; every rank moves its
        share over its own PCIe link."""
        L = self.L
        L.owned(self.vorticity).copy_(vorticity_rows_host, non_blocking=True)
        if char_func_rows_host is not None:                  # None: fixed body, the resident chi is kept
            L.owned(self.char_func).copy_(char_func_rows_host, non_blocking=True)
            self.comm.exchange([self.char_func], 2)          # the penalisation reads chi on the halo rows
        self.step(1)
        out_rows_host.copy_(L.owned(self.vorticity), non_blocking=True)

    # -- bench.py / diagnostics ----------------------------------------------------------------------------------
    def solve_flops(self):
        from .fd import solve_flops
        return solve_flops(self.nr, self.nz, self.factors) / self.L.world

    def solve_hbm_bytes(self):
        """per-rank algorithmic HBM bytes of the solve (80 B/pt + the 32 B/pt partition correction when P > 1)"""
        from .fd import solve_hbm_bytes
        b = solve_hbm_bytes(self.nr, self.nz, self.factors)
        if b is None:
            return None
        if self.L.world > 1:
            b += 32.0 * self.nr * self.nz
        return b / self.L.world

    def solver_basis(self):
        return self.factors["basis"]

    def solve_kernel_note(self):
        how = ("over NVLink peer memory (k_peer_block_put + device barrier)" if self.peer_halos
               else "by all-gather")
        return ("per rank, r-slabs, no transposes: k_dct_rows (DCT-II) on the owned rows + partitioned r solve "
                "(k_tri_sweep on the rank's own rows + k_tri_partition_correct, 2P interface rows gathered " + how +
                ") + k_dct_rows (DCT-III)")

    def scalars(self):
        self.comm.check()
        st = self.state.clone()
        self.comm.allreduce(st[7:8], "sum")
        st = st.cpu().numpy()
        cd = 2 * 2 * np.pi * self.dx * self.dx * self.brink_lam * st[7] / (np.pi * self.r_sph ** 2)
        return {"t": st[0], "dt": st[1], "umax": st[2], "iterations": int(st[6]), "Cd": cd}

    def gather_vorticity(self):
        """global (Nr, Nz) vorticity on every rank (diagnostics / tests)"""
        mine = self.L.owned(self.vorticity).contiguous()
        if self.L.world == 1:
            return mine
        parts = [torch.empty_like(mine) for _ in range(self.L.world)]
        dist.all_gather(parts, mine, group=self.comm.group)
        return torch.cat(parts, dim=0)
