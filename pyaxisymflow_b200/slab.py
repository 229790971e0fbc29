"""z-slab multi-GPU execution of the rigid-flow timestep.

One process per GPU (``torch.distributed``; NCCL over NVLink on the GPU box, gloo in the CPU
tests of the host logic).  Rank ``p`` of ``P`` owns the columns ``[p Nz/P, (p+1) Nz/P)`` of
every ``(Nr, Nz)`` field, stored as ``(Nr, Nz/P + 2 H)`` with a width-``H = 2`` halo -- the
reference's ``ghost_size`` (periodic_flow_past_sphere.py:49).  The kernels of
``libaxisym_b200`` take the slab placement in ``axb_grid_t`` (``kz0, nz_global, ku0, ku1``),
apply the global-boundary formulas by *global* column index and only write owned columns.

Communication per timestep
  * halo exchanges with the two z-neighbours (2 x Nr doubles per field and side, stored straight
    into the neighbours' halo columns over NVLink peer memory, NCCL send/recv as the fall-back):
    psi (before G-VEL, width 2 so that the velocity is also valid on one halo column),
    (vorticity, u_z) width 2 before G-ADV, the advected vorticity width 2 before the fused RK2
    diffusion pass (which evaluates its intermediate field on one halo column itself);
  * one MAX all-reduce of a double for the CFL time step (and a SUM when the drag is read);
  * the fast-diagonalisation solve: transpose to r-slabs ``(Nr/P, Nz)``, cosine transform of whole
    rows, r solve partitioned over the ranks (own-block sweeps + a 2P-row interface gather + one
    correction pass, :class:`PartitionedTridiagonal`), inverse transform, transpose back.  (Other
    flows: transposes to z-mode slabs around plain r sweeps, ``AXB_SLAB_TRANSPOSE4``; the
    eigen-decomposition path: local r GEMM, transpose, z GEMMs, transpose, local r GEMM.)  Each transpose moves ``Nr Nz 8 (P-1)/P^2`` bytes per rank,
    written directly into the peers' buffers (``axb_peer_block_put``) or through NCCL all-to-all.
The r-direction (axis reflection, 1/r terms) is never split.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .device import make_grid, ptr, stream_ptr

_call = _lib.call
HALO = 2


class SlabLayout:
    """Pure index arithmetic of the decomposition (no device, no communication)."""

    def __init__(self, nr, nz, world, rank, halo=HALO, periodic=False):
        if nz % world or nr % world:
            raise ValueError(f"grid {nr}x{nz} is not divisible by {world} ranks in both directions")
        if nz // world < 2 * halo + 1:
            raise ValueError("slabs are thinner than the stencil halo")
        self.nr, self.nz, self.world, self.rank, self.halo, self.periodic = nr, nz, world, rank, halo, periodic
        self.nzl = nz // world            # owned columns
        self.nrl = nr // world            # owned rows in the transposed (r-slab) layout
        self.nzs = self.nzl + 2 * halo    # stored columns
        self.z_begin = rank * self.nzl    # global index of the first owned column
        self.kz0 = self.z_begin - halo    # global index of local column 0
        self.ku0, self.ku1 = halo, halo + self.nzl
        self.r_begin = rank * self.nrl
        left, right = rank - 1, rank + 1
        if periodic:
            left, right = left % world, right % world
        self.left = left if left >= 0 else None
        self.right = right if right < world else None

    def grid(self, dx, ku0=None, ku1=None):
        return make_grid(self.nr, self.nzs, self.nzs, dx,
                         slab=(self.kz0, self.nz, self.ku0 if ku0 is None else ku0, self.ku1 if ku1 is None else ku1))

    def owned(self, t):
        """owned-column view of a stored slab field"""
        return t[:, self.ku0:self.ku1]

    def scatter_global(self, full):
        """stored slab (with halos filled where they exist) cut out of a global array -- test helper"""
        out = torch.zeros((self.nr, self.nzs), dtype=full.dtype, device=full.device)
        lo, hi = max(0, self.kz0), min(self.nz, self.kz0 + self.nzs)
        out[:, lo - self.kz0:hi - self.kz0] = full[:, lo:hi]
        return out


class SlabComm:
    """Halo exchange, slab<->row transposes and scalar reductions on a SlabLayout."""

    def __init__(self, layout, group=None):
        self.L, self.group = layout, group
        self._bufs = {}
        self._peer_fields = {}          # data_ptr -> (symmetric-memory handle, peer base pointers)

    # -- fields in peer-mapped memory: their halos are exchanged by direct NVLink stores -----------
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

    def _exchange_peer(self, fields, width, dx):
        L = self.L
        g = L.grid(dx)
        h = None
        for f in fields:
            h, ptrs = self._peer_fields[f.data_ptr()]
            _call("axb_halo_put", ctypes.byref(g), ptr(f),
                  ctypes.c_void_p(ptrs[L.left]) if L.left is not None else None,
                  ctypes.c_void_p(ptrs[L.right]) if L.right is not None else None, width, 0.0, stream_ptr())
        h.barrier(channel=1)             # one device-side barrier closes the whole batch

    # -- halos ----------------------------------------------------------------------------------
    def _buffers(self, key, n, like):
        b = self._bufs.get(key)
        if b is None or b[0].numel() != n or b[0].device != like.device:
            b = tuple(torch.empty(n, dtype=torch.float64, device=like.device) for _ in range(4))
            self._bufs[key] = b
        return b

    def exchange(self, fields, width, dx=1.0):
        """fill `width` halo columns of every field in `fields` from the z-neighbours"""
        L = self.L
        if L.world == 1 and not L.periodic:
            return
        if L.world > 1 and fields and all(f.is_cuda and f.data_ptr() in self._peer_fields for f in fields):
            return self._exchange_peer(fields, width, dx)
        n = L.nr * width
        ops, pending = [], []
        for i, f in enumerate(fields):
            sl, sr, rl, rr = self._buffers((i, width), n, f)
            if f.is_cuda:
                g = L.grid(dx)
                _call("axb_halo_pack", ctypes.byref(g), ptr(f), ptr(sl) if L.left is not None else None,
                      ptr(sr) if L.right is not None else None, width, stream_ptr())
            else:
                sl.copy_(f[:, L.ku0:L.ku0 + width].reshape(-1))
                sr.copy_(f[:, L.ku1 - width:L.ku1].reshape(-1))
            if L.left is not None:
                ops += [dist.P2POp(dist.isend, sl, L.left, self.group), dist.P2POp(dist.irecv, rl, L.left, self.group)]
            if L.right is not None:
                ops += [dist.P2POp(dist.isend, sr, L.right, self.group), dist.P2POp(dist.irecv, rr, L.right, self.group)]
            pending.append((f, rl, rr))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for f, rl, rr in pending:
            if f.is_cuda:
                g = L.grid(dx)
                _call("axb_halo_unpack", ctypes.byref(g), ptr(f), ptr(rl) if L.left is not None else None,
                      ptr(rr) if L.right is not None else None, width, 0.0, stream_ptr())
            else:
                if L.left is not None:
                    f[:, L.ku0 - width:L.ku0] = rl.view(L.nr, width)
                if L.right is not None:
                    f[:, L.ku1:L.ku1 + width] = rr.view(L.nr, width)

    # -- transposes around the z-direction GEMMs ---------------------------------------------------
    def slab_to_rows(self, t_slab, rows):
        """(nr x nzl) contiguous slab  ->  (nrl x nz) rows.  Block q of the slab (rows of rank q) is
        already contiguous, so the send side needs no packing."""
        L = self.L
        recv = torch.empty((L.world, L.nrl, L.nzl), dtype=torch.float64, device=t_slab.device)
        dist.all_to_all_single(recv.view(-1), t_slab.reshape(-1), group=self.group)
        if recv.is_cuda:
            _call("axb_blocks_to_rows", L.nrl, L.nzl, L.world, ptr(recv), ptr(rows), rows.stride(0), stream_ptr())
        else:
            rows.copy_(recv.permute(1, 0, 2).reshape(L.nrl, L.nz))
        return rows

    def rows_to_slab(self, rows, t_slab):
        """(nrl x nz) rows  ->  (nr x nzl) contiguous slab"""
        L = self.L
        send = torch.empty((L.world, L.nrl, L.nzl), dtype=torch.float64, device=rows.device)
        if rows.is_cuda:
            _call("axb_rows_to_blocks", L.nrl, L.nzl, L.world, ptr(rows), rows.stride(0), ptr(send), stream_ptr())
        else:
            send.copy_(rows.reshape(L.nrl, L.world, L.nzl).permute(1, 0, 2))
        dist.all_to_all_single(t_slab.view(-1), send.view(-1), group=self.group)
        return t_slab

    def allreduce(self, t, op="max"):
        if self.L.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=self.group)
        return t


class PeerTranspose:
    """slab <-> rows transposes through peer-mapped (symmetric) buffers: every rank stores its blocks
    straight into the destination buffers of its peers over NVLink (``axb_peer_block_put``) and a
    device-side barrier closes the exchange -- no pack / unpack passes, no collective library call."""

    def __init__(self, layout, group=None):
        import torch.distributed._symmetric_memory as symm

        L = self.L = layout
        g = group if group is not None else dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self.slab = symm.empty((L.nr, L.nzl), dtype=torch.float64, device=dev)
        self.rows = symm.empty((L.nrl, L.nz), dtype=torch.float64, device=dev)
        self.h_slab = symm.rendezvous(self.slab, g)
        self.h_rows = symm.rendezvous(self.rows, g)
        if self.h_slab.buffer_ptrs[L.rank] != self.slab.data_ptr() or self.h_rows.buffer_ptrs[L.rank] != self.rows.data_ptr():
            raise RuntimeError("symmetric buffers do not start at the tensors' data pointers")
        self.slab_ptrs = (ctypes.c_uint64 * L.world)(*self.h_slab.buffer_ptrs)
        self.rows_ptrs = (ctypes.c_uint64 * L.world)(*self.h_rows.buffer_ptrs)

    def slab_to_rows(self, src):
        """(nr x nzl) local slab (any pitch) -> the (nrl x nz) row buffers of all ranks; returns this rank's"""
        L = self.L
        _call("axb_peer_block_put", L.world, L.rank, self.rows_ptrs, L.rank * L.nzl, L.nz, ptr(src),
              L.nrl * src.stride(0), src.stride(0), L.nrl, L.nzl, stream_ptr())
        self.h_rows.barrier(channel=0)
        return self.rows

    def rows_to_slab(self, src):
        """(nrl x nz) local rows -> the (nr x nzl) slab buffers of all ranks; returns this rank's"""
        L = self.L
        _call("axb_peer_block_put", L.world, L.rank, self.slab_ptrs, L.rank * L.nrl * L.nzl, L.nzl, ptr(src),
              L.nzl, src.stride(0), L.nrl, L.nzl, stream_ptr())
        self.h_slab.barrier(channel=0)
        return self.slab


def _cuda_gemm(C, A, B, scale_m=None, scale_n=None, c0=0.0, c1=1.0):
    """C = A @ B (+ fused spectral scaling) through axb_dgemm; operands may be strided row views."""
    M, K = A.shape
    N = B.shape[1]
    _call("axb_dgemm", M, N, K, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(C), C.stride(0),
          ptr(scale_m), ptr(scale_n), float(c0), float(c1), stream_ptr())


def _cuda_fold(x, n, inverse):
    _call("axb_fd_fold", x.shape[0], n, ptr(x), x.stride(0), int(inverse), stream_ptr())


def _cuda_dct(dst, src, tables, inverse):
    rows, n = src.shape
    if inverse:
        _call("axb_dct3_rows", rows, n, ptr(src), src.stride(0), ptr(dst), dst.stride(0), ptr(tables), stream_ptr())
    else:
        _call("axb_dct2_rows", rows, n, ptr(src), src.stride(0), ptr(dst), dst.stride(0), ptr(tables), 1.0 / n, 2.0 / n,
              stream_ptr())


class _CudaTridiagonal:
    """r solve of this rank's z-modes (columns [z_begin, z_begin + nzl) of the spectral field)"""

    def __init__(self, L, f):
        tri = f["tri"]
        self.tri, self.c0, self.c1 = tri, float(f["c0"]), float(f["c1"])
        self.lam = f["lam_z"][L.z_begin:L.z_begin + L.nzl].contiguous()
        self.inv = None
        if L.nzl % 16 == 0:
            self.inv = torch.empty((L.nr, L.nzl), dtype=torch.float64, device=self.lam.device)
            self.rc = torch.empty((L.nr, 4), dtype=torch.float64, device=self.lam.device)
            _call("axb_tridiag_factor_columns", L.nr, L.nzl, ptr(tri["sub"]), ptr(tri["diag"]), ptr(tri["sup"]),
                  ptr(self.lam), ptr(tri["scale"]), self.c0, self.c1, ptr(self.inv), ptr(self.rc), stream_ptr())
        else:
            self.scratch = torch.empty((L.nr, L.nzl), dtype=torch.float64, device=self.lam.device)

    def __call__(self, x):
        nr, nzl = x.shape
        tri = self.tri
        if self.inv is not None:
            _call("axb_tridiag_solve_factored", nr, nzl, ptr(x), x.stride(0), ptr(self.inv), ptr(self.rc), stream_ptr())
        else:
            _call("axb_tridiag_solve_columns", nr, nzl, ptr(x), x.stride(0), ptr(tri["sub"]), ptr(tri["diag"]),
                  ptr(tri["sup"]), ptr(self.lam), ptr(tri["scale"]), self.c0, self.c1, ptr(self.scratch), stream_ptr())


class PartitionedTridiagonal:
    """r solve with the ROWS split over the ranks (r-slab layout, every rank holds all z-modes): the
    partition / SPIKE method.  Each rank solves its own diagonal block T_p g = rhs (the factored sweeps on
    its rows), the first and last row of every rank's g are gathered (2 P rows of Nz doubles), and
    x = g - V xl - W xr with the spikes V = T_p^-1 (a e_first), W = T_p^-1 (u e_last) and the two neighbouring
    interface unknowns xl, xr taken from the inverted 2P x 2P reduced system.  Everything but g is operator
    only and prepared here once.  Compared with transposing to z-mode slabs and back this trades two
    all-to-all transposes for a 2 P Nz gather and one 32 B/pt correction pass."""

    def __init__(self, layout, factors, group=None, peer_ptrs=None, host=False, sync=None):
        L, f = self.L, self.f = layout, factors
        tri = f["tri"]
        self.group, self.host = group, host
        self.sync = sync              # rank barrier after the interface put (default: the symmetric-memory barrier)
        c0, c1 = float(f["c0"]), float(f["c1"])
        r0, r1, n = L.r_begin, L.r_begin + L.nrl, L.nrl
        sub, diag, sup, scale, lam = tri["sub"], tri["diag"], tri["sup"], tri["scale"], f["lam_z"]
        self.sub_l, self.diag_l, self.sup_l = sub[r0:r1 - 1].contiguous(), diag[r0:r1].contiguous(), sup[r0:r1 - 1].contiguous()
        self.scale_l = None if scale is None else scale[r0:r1].contiguous()
        self.lam, self.c0, self.c1 = lam, c0, c1
        dev = lam.device
        if not host:
            self.inv = torch.empty((n, L.nz), dtype=torch.float64, device=dev)
            self.rc = torch.empty((n, 4), dtype=torch.float64, device=dev)
            _call("axb_tridiag_factor_columns", n, L.nz, ptr(self.sub_l), ptr(self.diag_l), ptr(self.sup_l), ptr(lam),
                  ptr(self.scale_l), c0, c1, ptr(self.inv), ptr(self.rc), stream_ptr())
            self.rc_unit = self.rc.clone()
            self.rc_unit[:, 1] = 1.0
        # spikes: unit-scaled local solves of a e_first and u e_last
        self.V = torch.zeros((n, L.nz), dtype=torch.float64, device=dev)
        self.W = torch.zeros((n, L.nz), dtype=torch.float64, device=dev)
        if L.rank > 0:
            self.V[0] = c1 * float(sub[r0 - 1])
            self._local_solve(self.V, unit=True)
        if L.rank < L.world - 1:
            self.W[-1] = c1 * float(sup[r1 - 1])
            self._local_solve(self.W, unit=True)
        # where the spikes matter: V decays from the first row down, W from the last row up (mode k like rho_k^m), so for
        # all but the lowest z modes the correction x = g - V xl - W xr touches a few dozen rows next to the interfaces.
        # Per 256-column group of the correction kernel: rows [0, vcut) need V, rows [wcut, n) need W.
        self.vcut = self.wcut = None
        if not host and not os.environ.get("AXB_PARTITION_FULL"):
            self.vcut, self.wcut = self._spike_cutoffs(self.V, self.W)
        # reduced interface system, inverted once on the host: unknowns (first, last) of every rank
        mine = torch.stack([self.V[0], self.V[-1], self.W[0], self.W[-1]]).contiguous()
        parts = [torch.empty_like(mine) for _ in range(L.world)]
        if L.world > 1:
            dist.all_gather(parts, mine, group=group)
        else:
            parts = [mine]
        co = torch.stack(parts).cpu().numpy()                      # (P, 4, nz)
        P2 = 2 * L.world
        R = np.zeros((L.nz, P2, P2))
        for p in range(L.world):
            for a in (0, 1):
                i = 2 * p + a
                R[:, i, i] = 1.0
                if p > 0:
                    R[:, i, 2 * (p - 1) + 1] = co[p, a]
                if p < L.world - 1:
                    R[:, i, 2 * (p + 1)] = co[p, 2 + a]
        Rinv = np.linalg.inv(R)
        CL = Rinv[:, 2 * (L.rank - 1) + 1, :].T if L.rank > 0 else np.zeros((P2, L.nz))
        CR = Rinv[:, 2 * (L.rank + 1), :].T if L.rank < L.world - 1 else np.zeros((P2, L.nz))
        self.CL = torch.from_numpy(np.ascontiguousarray(CL)).to(dev)
        self.CR = torch.from_numpy(np.ascontiguousarray(CR)).to(dev)
        self.n_iface = P2
        # gathered interface rows: peer-mapped buffer filled by direct stores, or a plain all-gather
        self.G, self.G_ptrs, self.G_handle = None, None, None
        if peer_ptrs is not None:
            self.G, self.G_handle, ptrs = peer_ptrs((P2, L.nz))
            self.G_ptrs = (ctypes.c_uint64 * L.world)(*ptrs)
        else:
            self.G = torch.zeros((P2, L.nz), dtype=torch.float64, device=dev)

    @staticmethod
    def _spike_cutoffs(V, W, tol=1e-20, group=256):
        """int32 device arrays, one entry per `group` columns: first row from which |V| <= tol max|V| in every column
        of the group, first row in which |W| > tol max|W| in some column of the group"""
        n, nz = V.shape
        rows = torch.arange(n, device=V.device)[:, None]
        mv = V.abs() > tol * max(float(V.abs().max()), 1e-300)
        mw = W.abs() > tol * max(float(W.abs().max()), 1e-300)
        vcol = torch.where(mv, rows + 1, torch.zeros_like(rows)).amax(0)          # rows [0, vcol) matter
        wcol = torch.where(mw, rows, torch.full_like(rows, n)).amin(0)            # rows [wcol, n) matter
        ng = (nz // 2 + 128) // 128                                               # the kernel's column groups
        pad = ng * group - nz
        if pad > 0:
            vcol = torch.cat([vcol, vcol.new_zeros(pad)])
            wcol = torch.cat([wcol, wcol.new_full((pad,), n)])
        vcut = vcol[:ng * group].view(ng, group).amax(1).to(torch.int32).contiguous()
        wcut = wcol[:ng * group].view(ng, group).amin(1).to(torch.int32).contiguous()
        return vcut, wcut

    def _local_solve(self, x, unit=False):
        from . import fd

        n, nz = x.shape
        if self.host:
            x.copy_(torch.from_numpy(fd.thomas_host(
                x.numpy(), self.sub_l.numpy(), self.diag_l.numpy(), self.sup_l.numpy(), self.lam.numpy(),
                None if (unit or self.scale_l is None) else self.scale_l.numpy(), self.c0, self.c1)))
            return
        _call("axb_tridiag_solve_factored", n, nz, ptr(x), x.stride(0), ptr(self.inv),
              ptr(self.rc_unit if unit else self.rc), stream_ptr())

    def __call__(self, x):
        """x: this rank's rows of the spectral right-hand side (nrl x nz), solved in place"""
        L = self.L
        n, nz = x.shape
        self._local_solve(x)
        if self.G_ptrs is not None:                   # rows 0 and n-1 straight into every rank's G over NVLink
            _call("axb_peer_block_put", L.world, L.rank, self.G_ptrs, L.rank * 2 * nz, nz, ptr(x), 0,
                  (n - 1) * x.stride(0), 2, nz, stream_ptr())
            if self.sync is not None:
                self.sync()
            else:
                self.G_handle.barrier(channel=0)
        else:
            mine = torch.stack([x[0], x[-1]]).contiguous()
            if L.world > 1:
                dist.all_gather_into_tensor(self.G.view(-1), mine.view(-1), group=self.group)
            else:
                self.G.copy_(mine)
        if self.host:
            xl = (self.CL * self.G).sum(0)
            xr = (self.CR * self.G).sum(0)
            x.copy_((x - self.V * xl) - self.W * xr)
            return
        _call("axb_tridiag_partition_correct_banded", n, nz, ptr(x), x.stride(0), ptr(self.V), ptr(self.W), ptr(self.G),
              ptr(self.CL), ptr(self.CR), self.n_iface, ptr(self.vcut) if self.vcut is not None else None,
              ptr(self.wcut) if self.wcut is not None else None, stream_ptr())


class SlabFdSolver:
    """Distributed fast-diagonalisation solve on z-slabs (factors replicated on every rank).

    eigen r method:       r GEMM on the slab -> all-to-all -> z transforms on rows -> all-to-all -> r GEMM
    tridiagonal r method: all-to-all -> forward z transform on rows (GEMM leaves or DCT-II) -> all-to-all ->
                          r solves of this rank's z-modes -> all-to-all -> backward z transform -> all-to-all
    ``gemm`` / ``fold`` / ``dct`` / ``tri`` are injectable so the plumbing runs on CPU in the gloo tests."""

    def __init__(self, layout, comm, factors, gemm=None, fold=None, dct=None, tri=None, peer=None, part=None):
        self.L, self.comm, self.f = layout, comm, factors
        self.peer = peer
        self.part = part          # PartitionedTridiagonal: r solve in the r-slab layout (2 transposes per solve)
        self.gemm = gemm or _cuda_gemm
        self.fold = fold or _cuda_fold
        self.dct = dct or _cuda_dct
        dev = factors["lam_z"].device
        L = layout
        self.t_slab = torch.empty((L.nr, L.nzl), dtype=torch.float64, device=dev)
        self.rows_a = torch.empty((L.nrl, L.nz), dtype=torch.float64, device=dev)
        self.rows_b = torch.empty((L.nrl, L.nz), dtype=torch.float64, device=dev)
        self.tri = None
        if factors.get("tri") is not None:
            self.tri = tri or _CudaTridiagonal(L, factors)
        else:
            self.lam_r_local = factors["lam_r"][L.r_begin:L.r_begin + L.nrl].contiguous()

    def _z_transform(self, inverse, rows_in=None, rows_out=None):
        """z transform of whole rows: rows_in (default rows_a; folded in place on the GEMM-leaf path) ->
        rows_out (default rows_b; DCT path only), which is returned"""
        f = self.f
        rows_a = self.rows_a if rows_in is None else rows_in
        if f.get("zfft") is not None:
            out = self.rows_b if rows_out is None else rows_out
            self.dct(out, rows_a, f["zfft"]["tables"], inverse)
            return out
        zs = f.get("zsplit")
        if zs is None:
            self.gemm(self.rows_b, rows_a, f["Rzb"] if inverse else f["Rz"])
            return self.rows_b
        if not inverse:
            for n in zs["fold_len"]:
                self.fold(rows_a, n, False)
        for n, off, F in zip(zs["leaf_n"], zs["leaf_off"], zs["bwd"] if inverse else zs["fwd"]):
            self.gemm(self.rows_b[:, off:off + n], rows_a[:, off:off + n], F)
        if inverse:
            for n in reversed(zs["fold_len"]):
                self.fold(self.rows_b, n, True)
        return self.rows_b

    def _solve_tridiagonal(self, psi_slab, rhs_slab):
        L = self.L
        if self.part is not None and self.f.get("zfft") is not None:
            # r-slab rows all the way: transpose, DCT-II, partitioned r solve, DCT-III, transpose
            pt = self.peer
            if pt is not None:
                rows = pt.slab_to_rows(L.owned(rhs_slab))
            else:
                self.t_slab.copy_(L.owned(rhs_slab))
                rows = self.comm.slab_to_rows(self.t_slab, self.rows_a)
            spec = self._z_transform(False, rows)                                 # -> rows_b
            self.part(spec)
            out = self._z_transform(True, spec, rows)                             # back into the row buffer
            if pt is not None:
                L.owned(psi_slab).copy_(pt.rows_to_slab(out))
            else:
                self.comm.rows_to_slab(out, self.t_slab)
                L.owned(psi_slab).copy_(self.t_slab)
            return
        if self.peer is not None:                                                 # transposes over peer memory
            pt = self.peer
            spec = self._z_transform(False, pt.slab_to_rows(L.owned(rhs_slab)))
            modes = pt.rows_to_slab(spec)
            self.tri(modes)
            out = self._z_transform(True, pt.slab_to_rows(modes))
            L.owned(psi_slab).copy_(pt.rows_to_slab(out))
            return
        self.t_slab.copy_(L.owned(rhs_slab))
        self.comm.slab_to_rows(self.t_slab, self.rows_a)                          # all-to-all #1
        spec = self._z_transform(False)
        self.comm.rows_to_slab(spec, self.t_slab)                                 # all-to-all #2: modes of this rank
        self.tri(self.t_slab)
        self.comm.slab_to_rows(self.t_slab, self.rows_a)                          # all-to-all #3
        out = self._z_transform(True)
        self.comm.rows_to_slab(out, self.t_slab)                                  # all-to-all #4
        L.owned(psi_slab).copy_(self.t_slab)

    def solve(self, psi_slab, rhs_slab):
        L, f = self.L, self.f
        if self.tri is not None:
            return self._solve_tridiagonal(psi_slab, rhs_slab)
        self.gemm(self.t_slab, f["Lr"], L.owned(rhs_slab))                       # r-transform, local columns
        self.comm.slab_to_rows(self.t_slab, self.rows_a)                          # all-to-all #1
        zs = f.get("zsplit")
        if zs is None:
            self.gemm(self.rows_b, self.rows_a, f["Rz"], self.lam_r_local, f["lam_z"], f["c0"], f["c1"])
            self.gemm(self.rows_a, self.rows_b, f["Rzb"])
        else:                                                                     # parity-split z transforms
            for n in zs["fold_len"]:
                self.fold(self.rows_a, n, False)
            for n, off, F in zip(zs["leaf_n"], zs["leaf_off"], zs["fwd"]):
                self.gemm(self.rows_b[:, off:off + n], self.rows_a[:, off:off + n], F, self.lam_r_local,
                          f["lam_z"][off:off + n], f["c0"], f["c1"])
            for n, off, B in zip(zs["leaf_n"], zs["leaf_off"], zs["bwd"]):
                self.gemm(self.rows_a[:, off:off + n], self.rows_b[:, off:off + n], B)
            for n in reversed(zs["fold_len"]):
                self.fold(self.rows_a, n, True)
        self.comm.rows_to_slab(self.rows_a, self.t_slab)                          # all-to-all #2
        self.gemm(L.owned(psi_slab), f["Lrb"], self.t_slab)                       # r back-transform

    def flops_per_rank(self):
        L = self.L
        if self.tri is not None:
            from .fd import solve_flops
            return solve_flops(L.nr, L.nz, self.f) / L.world
        zs = self.f.get("zsplit")
        z = 2.0 * L.nz * L.nz if zs is None else sum(2.0 * n * n for n in zs["leaf_n"])
        return 4.0 * L.nr * L.nr * L.nzl + 2.0 * L.nrl * z


class SlabRigidFlowStepper:
    """z-slab version of :class:`pyaxisymflow_b200.timestep.RigidFlowStepper` (same physics, same
    kernels; see the module docstring for the exchanges)."""

    def __init__(self, grid_size_z, grid_size_r=None, domain_AR=0.5, Re=100.0, U_0=1.0, r_sph=0.1, Z_cm=0.25,
                 R_cm=0.0, brink_lam=1e12, CFL=0.1, basis="analytic", group=None, r_method="auto",
                 z_method="auto"):
        from .fd import build_factors

        if not torch.cuda.is_available():
            raise _lib.AxbError("SlabRigidFlowStepper needs CUDA devices (no CPU fallback)")
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.nz = int(grid_size_z)
        self.nr = int(grid_size_r) if grid_size_r is not None else int(domain_AR * grid_size_z)
        self.dx = 1.0 / self.nz
        self.L = L = SlabLayout(self.nr, self.nz, world, rank)
        self.comm = SlabComm(L, group)
        self.U_0, self.r_sph, self.brink_lam, self.CFL = U_0, r_sph, brink_lam, CFL
        self.nu = U_0 * 2 * r_sph / Re
        self.T_ramp = 20 * r_sph / U_0
        self.dt_diff_limit = 0.9 * self.dx ** 2 / 4 / self.nu
        dx, nr = self.dx, self.nr
        z_full = np.linspace(0 + dx / 2, 1 - dx / 2, self.nz)
        z_loc = np.zeros(L.nzs)
        lo, hi = max(0, L.kz0), min(self.nz, L.kz0 + L.nzs)
        z_loc[lo - L.kz0:hi - L.kz0] = z_full[lo:hi]
        self.z1d = torch.from_numpy(z_loc).cuda()
        self.r1d = torch.from_numpy(np.linspace(0 + dx / 2, nr * dx - dx / 2, nr)).cuda()

        def field():
            return torch.zeros((nr, L.nzs), dtype=torch.float64, device="cuda")

        # the five fields whose halos travel live in peer-mapped memory when the box allows it
        xfield, self.peer_halos = field, False
        if world > 1 and not os.environ.get("AXB_SLAB_NCCL"):
            try:
                probe = self.comm.symmetric_field((nr, L.nzs))
                xfield, self.peer_halos = (lambda: self.comm.symmetric_field((nr, L.nzs))), True
                self.vorticity = probe
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    print(f"[pyaxisymflow_b200] peer-memory halos unavailable ({e!r}); using NCCL send/recv", flush=True)
        if not self.peer_halos:
            self.vorticity = field()
        self.psi = xfield()
        self.u_z, self.u_r, self.u_z_upen, self.u_r_upen = xfield(), field(), field(), field()
        self.char_func, self._tmp, self._w2 = field(), xfield(), xfield()
        self.state = torch.zeros(8, dtype=torch.float64, device="cuda")
        self.grid = L.grid(dx)
        # velocity is also evaluated on one halo column each side (needs the width-2 psi halo)
        ext0 = L.ku0 - 1 if L.left is not None else L.ku0
        ext1 = L.ku1 + 1 if L.right is not None else L.ku1
        self.grid_ext = L.grid(dx, ext0, ext1)
        # char_func on owned + halo columns directly (it is analytic)
        gall = L.grid(dx, max(0, -L.kz0), min(L.nzs, self.nz - L.kz0))
        _call("axb_smooth_heaviside_sphere", ctypes.byref(gall), ptr(self.char_func), None, ptr(self.z1d),
              ptr(self.r1d), float(Z_cm), float(R_cm), float(r_sph), float(dx * 2 ** 0.5), stream_ptr())
        if r_method == "auto":          # same rule as the single-GPU solver classes
            r_method = "tridiagonal" if max(nr, self.nz) >= 1536 else "eigen"
        if r_method != "tridiagonal" and z_method == "auto":
            z_method = "gemm"
        self.factors = build_factors("stokes", "homogenous_neumann_along_z_and_r", nr, self.nz, dx, basis,
                                     device="cuda", r_method=r_method, z_method=z_method)
        self.peer = None
        if self.factors.get("tri") is not None and world > 1 and not os.environ.get("AXB_SLAB_NCCL"):
            try:
                self.peer = PeerTranspose(L, group)
            except Exception as e:  # noqa: BLE001  (no peer mapping on this box: NCCL all-to-all instead)
                if rank == 0:
                    print(f"[pyaxisymflow_b200] peer-memory transposes unavailable ({e!r}); using NCCL all-to-all",
                          flush=True)
        self.part = None
        if self.factors.get("zfft") is not None and world > 1 and not os.environ.get("AXB_SLAB_TRANSPOSE4"):
            def peer_buf(shape):
                t = self.comm.symmetric_field(shape)
                h, ptrs = self.comm._peer_fields[t.data_ptr()]
                return t, h, ptrs
            self.part = PartitionedTridiagonal(L, self.factors, group,
                                               peer_ptrs=peer_buf if self.peer is not None else None)
        self.solver = SlabFdSolver(L, self.comm, self.factors, peer=self.peer, part=self.part)

    def seed_vorticity(self, seed=0, amplitude=1.0):
        """same global field as RigidFlowStepper.seed_vorticity, cut to this rank's slab"""
        L = self.L
        gen = torch.Generator(device="cuda")
        gen.manual_seed(seed)
        noise = torch.randn((self.nr, self.nz), dtype=torch.float64, device="cuda", generator=gen)
        zf = torch.linspace(self.dx / 2, 1 - self.dx / 2, self.nz, dtype=torch.float64, device="cuda")
        env = torch.exp(-((zf[None, :] - 0.5) ** 2 + self.r1d[:, None] ** 2) / 0.02)
        self.vorticity.copy_(L.scatter_global(amplitude * noise * env))

    def _enqueue(self, probe=None):
        s = stream_ptr()
        g, ge = ctypes.byref(self.grid), ctypes.byref(self.grid_ext)
        st = self.state
        sp = lambda i: ctypes.c_void_p(st.data_ptr() + 8 * i)  # noqa: E731
        w, psi = self.vorticity, self.psi
        sc = (self.U_0, self.T_ramp, 0.0, self.dt_diff_limit, self.CFL * self.dx)
        _call("axb_rigid_flow_scalars", 0, ptr(st), *sc, s)
        _call("axb_kill_boundary_vorticity_sine_z", g, ptr(w), ptr(self.z1d), 3, s)
        _call("axb_kill_boundary_vorticity_sine_r", g, ptr(w), ptr(self.r1d), 3, s)
        if probe is not None:
            probe[0].record()
        self.solver.solve(psi, w)
        if probe is not None:
            probe[1].record()
        self.comm.exchange([psi], 2, self.dx)
        _call("axb_velocity_from_psi", ge, ptr(self.u_z_upen), ptr(self.u_r_upen), ptr(psi), ptr(self.r1d), 0.0, 0.0,
              sp(4), sp(2), s)
        self.comm.allreduce(st[2:3], "max")
        _call("axb_rigid_flow_scalars", 1, ptr(st), *sc, s)
        _call("axb_penalise_update_vorticity", g, ptr(self.u_z), ptr(self.u_r), ptr(w), ptr(self.u_z_upen),
              ptr(self.u_r_upen), ptr(self.char_func), self.brink_lam, 0.0, sp(1), 0.0, 0.0, None, ptr(self.r1d),
              sp(3), s)
        self.comm.exchange([w, self.u_z], 2, self.dx)
        _call("axb_advect_vorticity_eno3", g, ptr(self._w2), ptr(w), ptr(self.u_z), ptr(self.u_r), 0.0, sp(1), s)
        self.comm.exchange([self._w2], 2, self.dx)            # width 2: the fused RK2 evaluates tmp on one halo column
        _call("axb_diffusion_rk2_fused", g, ptr(w), ptr(self._w2), ptr(self._tmp), ptr(self.r1d), self.nu, 0.0,
              sp(1), s)
        _call("axb_rigid_flow_scalars", 2, ptr(st), *sc, s)

    def step(self, n=1):
        for _ in range(n):
            self._enqueue()

    def step_probed(self):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._enqueue(probe=ev)
        return ev

    def step_host(self, vorticity_slab_host, char_func_slab_host, out_slab_host):
        """End-to-end form for host-resident callers: this rank's owned columns (pinned ``(Nr, Nz/P)`` host
        arrays) in, one step, owned vorticity columns out; every rank moves its share over its own PCIe link."""
        L = self.L
        L.owned(self.vorticity).copy_(vorticity_slab_host, non_blocking=True)
        if char_func_slab_host is not None:                  # None: fixed body, the resident chi is kept
            L.owned(self.char_func).copy_(char_func_slab_host, non_blocking=True)
            self._refresh_char_halo()                        # the penalisation reads chi one column into the halo
        self.step(1)
        out_slab_host.copy_(L.owned(self.vorticity), non_blocking=True)

    def _refresh_char_halo(self):
        """halo columns of the characteristic function after a host upload (plain NCCL send/recv: it is not one
        of the peer-mapped fields)"""
        L = self.L
        if L.world == 1:
            return
        f = self.char_func
        n = L.nr * 2
        sl, sr, rl, rr = (torch.empty(n, dtype=torch.float64, device=f.device) for _ in range(4))
        g = L.grid(self.dx)
        _call("axb_halo_pack", ctypes.byref(g), ptr(f), ptr(sl) if L.left is not None else None,
              ptr(sr) if L.right is not None else None, 2, stream_ptr())
        ops = []
        if L.left is not None:
            ops += [dist.P2POp(dist.isend, sl, L.left, self.comm.group), dist.P2POp(dist.irecv, rl, L.left, self.comm.group)]
        if L.right is not None:
            ops += [dist.P2POp(dist.isend, sr, L.right, self.comm.group), dist.P2POp(dist.irecv, rr, L.right, self.comm.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        _call("axb_halo_unpack", ctypes.byref(g), ptr(f), ptr(rl) if L.left is not None else None,
              ptr(rr) if L.right is not None else None, 2, 0.0, stream_ptr())

    def solve_flops(self):
        return self.solver.flops_per_rank()

    def solve_hbm_bytes(self):
        """per-rank algorithmic HBM bytes of the solve on the DCT path (the all-to-all traffic is NVLink's)"""
        from .fd import solve_hbm_bytes
        b = solve_hbm_bytes(self.nr, self.nz, self.factors)
        if b is not None and self.part is not None:
            b += 32.0 * self.nr * self.nz                      # the partition correction pass
        return None if b is None else b / self.L.world

    def solver_basis(self):
        return self.factors["basis"]

    def solve_kernel_note(self):
        zs = self.factors.get("zsplit")
        n = 0 if zs is None else len(zs["leaf_n"])
        if self.factors.get("zfft") is not None and self.part is not None:
            how = ("over NVLink peer memory (k_peer_block_put + device barrier)" if self.peer is not None
                   else "by NCCL all-to-all / all-gather")
            return ("per rank: k_dct_rows and the partitioned r solve (k_tri_sweep on the rank's own rows + "
                    "k_tri_partition_correct) on r-slabs; 2 transposes + one 2P-row interface gather " + how)
        if self.factors.get("zfft") is not None:
            how = ("4 transposes over NVLink peer memory (k_peer_block_put + device barrier)" if self.peer is not None
                   else "4 NCCL all-to-all")
            return ("per rank: k_dct_rows on r-slabs, k_tri_sweep on this rank's z-modes, " + how + " in between")
        if self.factors.get("tri") is not None:
            return (f"per rank: {'2 dense' if zs is None else '2x%d parity-split' % n} k_dgemm_tma z-transforms on "
                    "r-slabs, k_tri_sweep on this rank's z-modes, 4 NCCL all-to-all in between")
        return (f"k_dgemm_tma per rank: 2 slab r-transforms + {'2 dense' if zs is None else '2x%d parity-split' % n} "
                "z-transforms on r-slabs, 2 NCCL all-to-all in between")

    def scalars(self):
        st = self.state.clone()
        self.comm.allreduce(st[7:8], "sum")
        st = st.cpu().numpy()
        cd = 2 * 2 * np.pi * self.dx * self.dx * self.brink_lam * st[7] / (np.pi * self.r_sph ** 2)
        return {"t": st[0], "dt": st[1], "umax": st[2], "iterations": int(st[6]), "Cd": cd}

    def gather_vorticity(self):
        """global (Nr, Nz) vorticity on every rank (diagnostics / tests)"""
        L = self.L
        mine = L.owned(self.vorticity).contiguous()
        parts = [torch.empty_like(mine) for _ in range(L.world)]
        dist.all_gather(parts, mine, group=self.comm.group)
        return torch.cat(parts, dim=1)
