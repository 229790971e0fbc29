"""Device-resident timestep drivers: the loop bodies of the reference's example scripts with
every kernel call replaced by a fused ``libaxisym_b200`` launch and every piece of NumPy glue
between them (free-stream ramp, CFL reduction, drag sum, ``t += dt``) moved onto the GPU, so
that one timestep is a fixed sequence of launches with no host round trip -- and can therefore
be captured once in a CUDA graph and replayed.

:class:`RigidFlowStepper` is the body of ``examples/FlowPastSphere/flow_past_sphere.py:107-207``
(configs C1 / C4 of BASELINE.json) and, with ``periodic=True``, of
``examples/PeriodicFlowPastSphere/periodic_flow_past_sphere.py:95-183`` (config C2).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .device import make_grid, ptr, stream_ptr
from .fd import FastDiagonalisationStokesSolver

_call = _lib.call

# layout of the device scalar block (include/axisym_b200.h, axb_rigid_flow_scalars)
S_T, S_DT, S_UMAX, S_SUM, S_UZADD, S_URADD, S_IT, S_SUMCOPY = range(8)


class RigidFlowStepper:
    """Flow past a Brinkman-penalised rigid sphere, one fused step per call.

    Fields (CUDA float64 tensors, shape (Nr, Nz)): ``vorticity, psi, u_z, u_r, char_func``.
    Launch sequence of one step:
      scalars(0) -> kill_z -> kill_r -> solve (psi; DCT-II / rfft + two tridiagonal sweeps + DCT-III / irfft on large
      grids, the four DMMA GEMMs of the eigen-decomposition on small ones) -> velocity(+U, max) -> scalars(1: dt)
      -> penalise+curl+drag -> ENO3 advect (w -> w2) -> fused RK2 diffusion (w2 -> w; two stages with ghost
      refreshes in between when z is periodic) -> scalars(2: t += dt)
    """

    def __init__(self, grid_size_z, domain_AR=0.5, Re=100.0, U_0=1.0, r_sph=0.1, Z_cm=0.25, R_cm=0.0,
                 brink_lam=1e12, CFL=0.1, T_ramp=None, periodic=False, ghost_size=2, basis="auto",
                 grid_size_r=None, use_graph=False, r_method="auto", z_method="auto"):
        if not torch.cuda.is_available():
            raise _lib.AxbError("RigidFlowStepper needs a CUDA device (no CPU fallback)")
        self.nz = int(grid_size_z)
        self.nr = int(grid_size_r) if grid_size_r is not None else int(domain_AR * grid_size_z)
        self.dx = 1.0 / self.nz
        self.periodic, self.ghost = bool(periodic), int(ghost_size)
        self.U_0, self.r_sph, self.brink_lam, self.CFL = U_0, r_sph, brink_lam, CFL
        self.nu = U_0 * 2 * r_sph / Re
        self.T_ramp = 20 * r_sph / U_0 if T_ramp is None else T_ramp
        self.dt_diff_limit = 0.9 * self.dx ** 2 / 4 / self.nu
        self.ur_ramp = 5e-2 if periodic else 0.0   # periodic_flow_past_sphere.py:113
        dx, nr, nz = self.dx, self.nr, self.nz
        dev = "cuda"
        # z = linspace(dx/2, 1 - dx/2, nz), r = linspace(dx/2, AR - dx/2, nr)   (flow_past_sphere.py:58-59)
        self.z1d = torch.from_numpy(np.linspace(0 + dx / 2, 1 - dx / 2, nz)).to(dev)
        self.r1d = torch.from_numpy(np.linspace(0 + dx / 2, nr * dx - dx / 2, nr)).to(dev)

        def field():
            return torch.zeros((nr, nz), dtype=torch.float64, device=dev)

        self.vorticity, self.psi = field(), field()
        self.u_z, self.u_r, self.u_z_upen, self.u_r_upen = field(), field(), field(), field()
        self.char_func, self._tmp, self._w2 = field(), field(), field()
        self.state = torch.zeros(8, dtype=torch.float64, device=dev)
        self.grid = make_grid(nr, nz, nz, dx)
        _call("axb_smooth_heaviside_sphere", ctypes.byref(self.grid), ptr(self.char_func), None, ptr(self.z1d),
              ptr(self.r1d), float(Z_cm), float(R_cm), float(r_sph), float(dx * 2 ** 0.5), stream_ptr())
        if self.periodic:
            g = self.ghost
            self.solver = FastDiagonalisationStokesSolver(
                nr, nz - 2 * g, dx, bc_type="homogenous_neumann_along_r_and_periodic_along_z", basis=basis,
                r_method=r_method, z_method=z_method)
        else:
            self.solver = FastDiagonalisationStokesSolver(nr, nz, dx, basis=basis, r_method=r_method,
                                                          z_method=z_method)
        if self.periodic:
            _call("axb_periodic_ghost_comm", ctypes.byref(self.grid), ptr(self.char_func), self.ghost, 0.0, 0.0,
                  stream_ptr())
        self._graph = None
        self._graphs3 = None
        self._use_graph = use_graph
        self.graph_launches = 0          # kernels inside the captured step
        self.launches_replayed = 0       # kernels launched through graph replays so far

    # -- one step, enqueued on the current stream ---------------------------------------------
    def _enqueue(self, probe=None, parts=(0, 1, 2)):
        """parts: 0 = everything before the solve, 1 = the solve, 2 = everything after it"""
        s = stream_ptr()
        g = ctypes.byref(self.grid)
        st = self.state
        sp = lambda i: ctypes.c_void_p(st.data_ptr() + 8 * i)  # noqa: E731
        w, psi = self.vorticity, self.psi
        if 0 in parts:
            _call("axb_rigid_flow_scalars", 0, ptr(st), self.U_0, self.T_ramp, self.ur_ramp, self.dt_diff_limit,
                  self.CFL * self.dx, s)
            if not self.periodic:
                _call("axb_kill_boundary_vorticity_sine_z", g, ptr(w), ptr(self.z1d), 3, s)
            _call("axb_kill_boundary_vorticity_sine_r", g, ptr(w), ptr(self.r1d), 3, s)
        if probe is not None:
            probe[0].record()
        if 1 not in parts:
            pass
        elif self.periodic:
            gh = self.ghost
            off = 8 * gh
            _lib.call("axb_fd_solve", ctypes.byref(self.solver.plan), ctypes.c_void_p(psi.data_ptr() + off), self.nz,
                      ctypes.c_void_p(w.data_ptr() + off), self.nz, s)
            _call("axb_periodic_ghost_comm", g, ptr(psi), gh, 0.0, 0.0, s)
        else:
            _lib.call("axb_fd_solve", ctypes.byref(self.solver.plan), ptr(psi), self.nz, ptr(w), self.nz, s)
        if probe is not None:
            probe[1].record()
        if 2 not in parts:
            return
        _call("axb_velocity_from_psi", g, ptr(self.u_z_upen), ptr(self.u_r_upen), ptr(psi), ptr(self.r1d), 0.0, 0.0,
              sp(S_UZADD), sp(S_UMAX), s)
        _call("axb_rigid_flow_scalars", 1, ptr(st), self.U_0, self.T_ramp, self.ur_ramp, self.dt_diff_limit, self.CFL * self.dx, s)
        if self.periodic:
            _call("axb_periodic_ghost_comm", g, ptr(self.u_r_upen), self.ghost, 0.0, 0.0, s)
            _call("axb_periodic_ghost_comm", g, ptr(self.u_z_upen), self.ghost, 0.0, 0.0, s)
        _call("axb_penalise_update_vorticity", g, ptr(self.u_z), ptr(self.u_r), ptr(w), ptr(self.u_z_upen),
              ptr(self.u_r_upen), ptr(self.char_func), self.brink_lam, 0.0, sp(S_DT), 0.0, 0.0, None, ptr(self.r1d),
              sp(S_SUM), s)
        _call("axb_advect_vorticity_eno3", g, ptr(self._w2), ptr(w), ptr(self.u_z), ptr(self.u_r), 0.0, sp(S_DT), s)
        if self.periodic:
            _call("axb_periodic_ghost_comm", g, ptr(self._w2), self.ghost, 0.0, 0.0, s)
        if self.periodic:
            _call("axb_diffusion_rk2_stage1", g, ptr(self._tmp), ptr(self._w2), ptr(self.r1d), self.nu, 0.0, sp(S_DT), s)
            _call("axb_periodic_ghost_comm", g, ptr(self._tmp), self.ghost, 0.0, 0.0, s)
            _call("axb_diffusion_rk2_stage2", g, ptr(w), ptr(self._w2), ptr(self._tmp), ptr(self.r1d), self.nu, 0.0,
                  sp(S_DT), s)
        else:       # both RK2 stages in one pass, the intermediate field stays on chip
            _call("axb_diffusion_rk2_fused", g, ptr(w), ptr(self._w2), ptr(self._tmp), ptr(self.r1d), self.nu, 0.0,
                  sp(S_DT), s)
        _call("axb_rigid_flow_scalars", 2, ptr(st), self.U_0, self.T_ramp, self.ur_ramp, self.dt_diff_limit, self.CFL * self.dx, s)

    def step(self, n=1):
        """advance n timesteps (asynchronous; call :meth:`scalars` or synchronise to read back)"""
        if self._use_graph:
            if self._graph is None:
                self._enqueue()                      # warm-up outside capture (sets kernel attributes)
                torch.cuda.synchronize()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                self._graph = torch.cuda.CUDAGraph()
                before = _lib.launch_count()
                with torch.cuda.graph(self._graph, stream=side):
                    self._enqueue()
                self.graph_launches = _lib.launch_count() - before    # kernels one replay launches
                n -= 1
            for _ in range(n):
                self._graph.replay()
                self.launches_replayed += self.graph_launches
        else:
            for _ in range(n):
                self._enqueue()

    def _capture(self, parts):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(graph, stream=side):
            self._enqueue(parts=parts)
        return graph, _lib.launch_count() - before

    def step_probed(self):
        """one step with CUDA events around the solve (bench.py's roofline leg).  With ``use_graph`` the step
        is replayed as three graphs (before / solve / after) so that the events sit between graph launches."""
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        if not self._use_graph:
            self._enqueue(probe=ev)
            return ev
        if self._graphs3 is None:
            self._enqueue(probe=ev)                  # the first call is the warm-up (sets kernel attributes): it IS
            torch.cuda.synchronize()                 # the step, launched kernel by kernel; later calls replay
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

    def solve_flops(self):
        return self.solver.flops()

    def solve_hbm_bytes(self):
        return self.solver.hbm_bytes()

    def solver_basis(self):
        return self.solver.basis

    def solve_kernel_note(self):
        return self.solver.kernel_note()

    def seed_vorticity(self, seed=0, amplitude=1.0):
        """seeded band-limited blob (SURVEY.md 8d synthetic input (ii)): N(0,1) * exp(-((Z-1/2)^2+R^2)/0.02)"""
        gen = torch.Generator(device="cuda")
        gen.manual_seed(seed)
        noise = torch.randn((self.nr, self.nz), dtype=torch.float64, device="cuda", generator=gen)
        env = torch.exp(-((self.z1d[None, :] - 0.5) ** 2 + self.r1d[:, None] ** 2) / 0.02)
        self.vorticity.copy_(amplitude * noise * env)

    def step_host(self, vorticity_host, char_func_host, out_host):
        """End-to-end form for host-resident callers: pinned host fields in, one step, vorticity out.
        (bench.py's `e2e`: both copies are inside the timed region.)  ``char_func_host=None`` keeps the resident
        characteristic function: the body of flow_past_sphere.py is fixed, its char_func is computed once before the
        loop (:88-96) and the only per-step input is the vorticity."""
        self.vorticity.copy_(vorticity_host, non_blocking=True)
        if char_func_host is not None:
            self.char_func.copy_(char_func_host, non_blocking=True)
        self.step(1)
        out_host.copy_(self.vorticity, non_blocking=True)

    # -- diagnostics ----------------------------------------------------------------------------
    def scalars(self):
        st = self.state.cpu().numpy()
        cd = 2 * 2 * np.pi * self.dx * self.dx * self.brink_lam * st[S_SUMCOPY] / (np.pi * self.r_sph ** 2)
        return {"t": st[S_T], "dt": st[S_DT], "umax": st[S_UMAX], "iterations": int(st[S_IT]), "Cd": cd}


class HostStepPipeline:
    """Streams host-resident cases through one :class:`RigidFlowStepper`: every :meth:`submit` does what
    ``step_host`` does (pinned vorticity + characteristic function in, one step, vorticity out), but the
    H2D copy of case k+1 and the D2H copy of case k-1 run on their own streams and overlap the step of
    case k (double-buffered device staging, PCIe is full duplex).  Results are the same as calling
    ``step_host`` case by case; call :meth:`drain` before reading the last outputs.  With
    ``char_func_host=None`` the stepper's resident characteristic function is used (a fixed body, set once):
    half the H2D bytes and one device copy less per case."""

    def __init__(self, stepper):
        self.st = stepper
        like = stepper.vorticity
        self.in_w = [torch.empty_like(like) for _ in range(2)]
        self.in_c = [torch.empty_like(like) for _ in range(2)]
        self.out = [torch.empty_like(like) for _ in range(2)]
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        ev = lambda: [torch.cuda.Event() for _ in range(2)]  # noqa: E731
        self.ev_in, self.ev_in_free, self.ev_out, self.ev_out_free = ev(), ev(), ev(), ev()
        self.k = 0

    def submit(self, vorticity_host, char_func_host, out_host):
        st, b = self.st, self.k & 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_in_free[b])          # the step two cases ago has consumed staging b
            self.in_w[b].copy_(vorticity_host, non_blocking=True)
            if char_func_host is not None:
                self.in_c[b].copy_(char_func_host, non_blocking=True)
            self.ev_in[b].record(self.s_in)
        main.wait_event(self.ev_in[b])
        st.vorticity.copy_(self.in_w[b])
        if char_func_host is not None:
            st.char_func.copy_(self.in_c[b])
        self.ev_in_free[b].record(main)
        st.step(1)
        main.wait_event(self.ev_out_free[b])                  # the D2H two cases ago has left out[b]
        self.out[b].copy_(st.vorticity)
        self.ev_out[b].record(main)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_out[b])
            out_host.copy_(self.out[b], non_blocking=True)
            self.ev_out_free[b].record(self.s_out)
        self.k += 1

    def drain(self):
        self.s_out.synchronize()
        torch.cuda.current_stream().synchronize()



def _capture_step(obj, enqueue):
    """warm-up call + capture of one step as a CUDA graph; records how many kernels a replay launches"""
    enqueue()                                        # warm-up outside capture
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    n0 = _lib.launch_count()
    with torch.cuda.graph(graph, stream=side):
        enqueue()
    obj.graph_launches = _lib.launch_count() - n0
    return graph


class _FieldSet:
    """small helper: float64 CUDA fields of one (nr, nz) grid plus the ctypes grid descriptor.  With a ``pool``
    (:class:`_WidePool`) the fields are this member's column blocks of (nr, batch nz) tensors shared by an
    ensemble, so the row pitch is batch nz."""

    def __init__(self, nr, nz, dx, pool=None, member=0):
        self.nr, self.nz, self.dx = nr, nz, dx
        self.pool, self.member, self._n = pool, member, 0
        self.grid = make_grid(nr, nz, nz if pool is None else pool.batch * nz, dx)
        self.g = ctypes.byref(self.grid)

    def new(self, n=1):
        if self.pool is None:
            f = [torch.zeros((self.nr, self.nz), dtype=torch.float64, device="cuda") for _ in range(n)]
        else:
            f = [self.pool.view(self._n + i, self.member) for i in range(n)]
            self._n += n
        return f[0] if n == 1 else f


class _WidePool:
    """Storage of a batched ensemble (SURVEY.md 8e "Ensemble"): field i of all members lives in ONE (nr, batch nz)
    tensor, member m owning the columns [m nz, (m+1) nz).  Every per-member kernel sees an ordinary (nr, nz) field of
    pitch batch nz; the solve sees (batch nr) rows of nz doubles for the z transforms -- element (j, m nz + k) is
    row 8 j + m -- and batch nz independent columns for the r sweeps, so one launch serves all members."""

    def __init__(self, nr, nz, batch):
        self.nr, self.nz, self.batch = nr, nz, batch
        self.wide = []

    def view(self, i, member):
        while len(self.wide) <= i:
            self.wide.append(torch.zeros((self.nr, self.batch * self.nz), dtype=torch.float64, device="cuda"))
        return self.wide[i][:, member * self.nz:(member + 1) * self.nz]


class SoftSphereStepper:
    """Loop body of ``examples/SoftSphereStreaming/soft_sphere_streaming.py:129-274`` (config C3):
    a tethered hyperelastic sphere (reference-map solid) driven by an oscillating rigid core.

    Per step: boundary damping, streamfunction solve, velocity (+CFL max), running averages,
    ENO3 advection of both reference maps, level-set pinning, ENO3 vorticity advection, Heaviside +
    inside mask, least-squares extrapolation of the maps, solid stress (blend fused), div(tau) and
    its curl, moving-tether Heaviside, Brinkman penalisation (+curl), RK2 diffusion.
    The narrow-band re-initialisation ``skfmm.distance`` (third party, soft_sphere_streaming.py:196-199,
    SURVEY.md 8f-3) runs on the GPU when ``reinit_levelset=True`` (``csrc/reinit.cu``, band =
    ``extrap_zone`` like the driver's ``reinit_band``); by default it is left out, which is what
    BASELINE.md's config-C3 timing excludes on both the CPU and the GPU side, and ``ball_phi`` is only
    pinned from the reference map.
    One small D2H (the CFL max) per step, like the reference's ``np.amax``; the LS sweep loop
    synchronises anyway.
    """

    def __init__(self, grid_size_z=256, domain_AR=0.5, grid_size_r=None, r_ball=0.15, freq=16.0, nond_AC=0.125,
                 e=0.1, Cauchy=0.1, zeta=0.25, brink_lam=1e8, CFL=0.1, rho_f=1.0, Z_cm=0.5, R_cm=0.0, basis="auto",
                 reinit_levelset=False, device_scalars=False, use_graph=False, ls_sweeps=12, fused_solid=True,
                 overlap_ls=True):
        if not torch.cuda.is_available():
            raise _lib.AxbError("SoftSphereStepper needs a CUDA device (no CPU fallback)")
        if device_scalars and reinit_levelset:
            raise ValueError("the level-set re-initialisation polls a host flag: not available with device_scalars")
        nz = int(grid_size_z)
        nr = int(grid_size_r) if grid_size_r is not None else int(domain_AR * nz)
        dx = 1.0 / nz
        self.F = F = _FieldSet(nr, nz, dx)
        self.nr, self.nz, self.dx = nr, nz, dx
        self.CFL, self.brink_lam, self.rho_f = CFL, brink_lam, rho_f
        self.moll_zone = dx * 2
        self.extrap_zone = self.moll_zone + 4 * dx
        self.r_ball, self.Z_cm, self.R_cm, self.e = r_ball, Z_cm, R_cm, e
        self.freq = freq
        self.freqTimer_limit = 1 / freq
        self.omega = 2 * np.pi * freq
        self.U_0 = e * r_ball * self.omega
        Rs = (e / nond_AC) ** 2
        self.nu = e * self.U_0 * r_ball / Rs
        self.G = e * rho_f * (r_ball * self.omega) ** 2 / Cauchy
        self.fixed_rad = zeta * r_ball
        self.eps = np.finfo(float).eps
        z = np.linspace(0 + dx / 2, 1 - dx / 2, nz)
        r = np.linspace(0 + dx / 2, nr * dx - dx / 2, nr)
        self.z1d, self.r1d = torch.from_numpy(z).cuda(), torch.from_numpy(r).cuda()
        # grid axes of the doubled array handed to the LS routine: the reference passes (z, z)
        # (extrapolate_eta_using_least_squares_unb.py:27), which is z[:2 nr] for any aspect ratio
        self.gy = torch.from_numpy(np.linspace(dx / 2, 2 * nr * dx - dx / 2, 2 * nr)).cuda()
        (self.vorticity, self.psi, self.u_z, self.u_r, self.u_z_upen, self.u_r_upen, self._tmp, self._w2,
         self.eta1, self.eta2, self._e1b, self._e2b, self.ball_phi, self.ball_char_func, self.tether_char_func,
         self.s11, self.s12, self.s22, self.e1z, self.e1r, self.e2z, self.e2r, self.tau_z, self.tau_r,
         self.avg_psi, self.avg_phi) = F.new(26)
        self.inside_solid = torch.zeros((nr, nz), dtype=torch.uint8, device="cuda")
        s = stream_ptr()
        _call("axb_smooth_heaviside_sphere", F.g, ptr(self.ball_char_func), ptr(self.ball_phi), ptr(self.z1d),
              ptr(self.r1d), Z_cm, R_cm, r_ball, self.moll_zone, s)
        self.eta1.copy_(self.z1d[None, :].expand(nr, nz))
        self.eta2.copy_(self.r1d[:, None].expand(nr, nz))
        self.solver = FastDiagonalisationStokesSolver(nr, nz, dx, basis=basis)
        nbytes = int(_call("axb_ls_workspace_bytes", 2 * nr, nz))
        self._ls_work = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        self._ls_bytes = nbytes
        self._umax = torch.zeros(1, dtype=torch.float64, device="cuda")
        self._reinit = None
        if reinit_levelset:
            from .reinit import NarrowBandReinit

            self._reinit = NarrowBandReinit(nr, nz)
        self.t, self.freqTimer, self.it, self.dt = 0.0, 0.0, 0, 0.0
        self.cycles, self.on_cycle = 0, None      # on_cycle(stepper): called with the completed-cycle averages
        self.tEnd = 30 / freq
        # device-resident loop scalars (include/axisym_b200.h, axb_soft_sphere_scalars): no host round trip per step,
        # the LS extrapolation with device-terminated sweeps, the whole step one replayed CUDA graph
        self.device_scalars = bool(device_scalars)
        self.fused_solid = bool(fused_solid)          # device mode: one pass for the elastic stress (else 3 calls)
        # device mode: the LS sweeps (36 nearly empty launches, ~0.2 ms at 2048 x 8192) run on a high-priority side
        # stream while the vorticity advection and the tether Heaviside, which do not depend on them, use the GPU
        self.overlap_ls = bool(overlap_ls)
        self._ls_stream = None
        self._use_graph, self._graph = bool(use_graph) and self.device_scalars, None
        self.graph_launches, self.launches_replayed = 0, 0
        if self.device_scalars:
            self.state = torch.zeros(16, dtype=torch.float64, device="cuda")
            self._ls_status = torch.zeros(2, dtype=torch.int32, device="cuda")
            self._ls_sweeps = int(ls_sweeps)
            self.avg_psi_last, self.avg_phi_last = F.new(2)
            self.ls_sweeps = 0

    def step(self, n=1):
        if not self.device_scalars:
            for _ in range(n):
                self._one()
            return
        if not self._use_graph:
            for _ in range(n):
                self._one_dev()
            return
        if self._graph is None:
            self._graph = _capture_step(self, self._one_dev)
            n -= 1
        for _ in range(n):
            self._graph.replay()
            self.launches_replayed += self.graph_launches

    def _one_dev(self):
        """the step of `_one` with every host decision on the device (state block, axb_soft_sphere_scalars) and no
        buffer swap -- the advected reference maps / vorticity live in the scratch fields for the rest of the step
        and the LS extrapolation / the last RK2 stage land them back in `eta1`, `eta2` / `vorticity` -- so the launch
        sequence is the same every step"""
        s, g, dx = stream_ptr(), self.F.g, self.dx
        w, w2, e1b, e2b = self.vorticity, self._w2, self._e1b, self._e2b
        base = self.state.data_ptr()

        def sp(i):
            return ctypes.c_void_p(base + 8 * i)

        def scalars(phase):
            _call("axb_soft_sphere_scalars", phase, ptr(self.state), self.CFL * dx / np.sqrt(self.G / self.rho_f),
                  self.CFL * dx, self.eps, 0.9 * dx ** 2 / 4 / self.nu, self.freqTimer_limit, self.tEnd, self.omega,
                  self.U_0, self.Z_cm, self.e * self.r_ball, s)

        _call("axb_kill_boundary_vorticity_sine_z", g, ptr(w), ptr(self.z1d), 3, s)
        _call("axb_kill_boundary_vorticity_sine_r", g, ptr(w), ptr(self.r1d), 3, s)
        _lib.call("axb_fd_solve", ctypes.byref(self.solver.plan), ptr(self.psi), self.nz, ptr(w), self.nz, s)
        _call("axb_velocity_from_psi", g, ptr(self.u_z_upen), ptr(self.u_r_upen), ptr(self.psi), ptr(self.r1d), 0.0, 0.0,
              None, sp(2), s)                          # state[2] was zeroed by phase 2
        scalars(1)
        _call("axb_cycle_average3", g, ptr(self.avg_psi), ptr(self.psi), ptr(self.avg_psi_last), ptr(self.avg_phi),
              ptr(self.ball_phi), ptr(self.avg_phi_last), None, None, None, sp(1), sp(8), s)
        _call("axb_advect_refmap_eno3", g, ptr(e1b), ptr(e2b), ptr(self.eta1), ptr(self.eta2), ptr(self.u_z_upen),
              ptr(self.u_r_upen), 0.0, sp(1), s)
        _call("axb_pin_level_set", g, ptr(self.ball_phi), None, ptr(e1b), ptr(e2b), self.Z_cm, self.R_cm, self.r_ball,
              -3 * dx, s)

        def ls(parts, stream):
            _call("axb_ls_extrapolate_eta_device_parts", g, ptr(self.ball_phi), ptr(self.inside_solid), ptr(e1b), ptr(e2b),
                  ptr(self.eta1), ptr(self.eta2), self.extrap_zone, ptr(self.z1d), ptr(self.gy), ptr(self._ls_work),
                  self._ls_bytes, self._ls_sweeps, ptr(self._ls_status), parts, stream)

        def tether():
            _call("axb_smooth_heaviside_sphere_dev", g, ptr(self.tether_char_func), None, ptr(self.z1d), ptr(self.r1d),
                  sp(6), self.R_cm, self.fixed_rad, self.moll_zone, s)

        if self.overlap_ls:
            _call("axb_smooth_heaviside_mask", g, ptr(self.ball_char_func), ptr(self.inside_solid), ptr(self.ball_phi),
                  self.moll_zone, 0.5, 0, s)
            ls(1, s)
            # fork: the sweeps on the side stream, the independent HBM-bound passes on this one (same data flow as below)
            if self._ls_stream is None:
                self._ls_stream = torch.cuda.Stream(priority=-1)
            cur, side = torch.cuda.current_stream(), self._ls_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                ls(2, stream_ptr())
            _call("axb_advect_vorticity_eno3", g, ptr(w2), ptr(w), ptr(self.u_z_upen), ptr(self.u_r_upen), 0.0, sp(1), s)
            tether()
            cur.wait_stream(side)
            ls(4, s)
        else:
            _call("axb_advect_vorticity_eno3", g, ptr(w2), ptr(w), ptr(self.u_z_upen), ptr(self.u_r_upen), 0.0, sp(1), s)
            _call("axb_smooth_heaviside_mask", g, ptr(self.ball_char_func), ptr(self.inside_solid), ptr(self.ball_phi),
                  self.moll_zone, 0.5, 0, s)
            ls(7, s)
        if self.fused_solid:
            # sigma -> tau -> curl in one shared-memory pass (the driver never looks at the seven intermediates)
            _call("axb_solid_stress_vorticity_update", g, ptr(w2), ptr(self.eta1), ptr(self.eta2),
                  ptr(self.ball_char_func), ptr(self.r1d), self.G, 0.0, sp(1), 0, s)
        else:
            _call("axb_solid_sigma", g, ptr(self.s11), ptr(self.s12), ptr(self.s22), self.G, ptr(self.eta1),
                  ptr(self.eta2), ptr(self.e1z), ptr(self.e1r), ptr(self.e2z), ptr(self.e2r), ptr(self.ball_char_func), s)
            _call("axb_solid_tau", g, ptr(self.tau_z), ptr(self.tau_r), ptr(self.s11), ptr(self.s12), ptr(self.s22),
                  ptr(self.r1d), s)
            _call("axb_solid_vorticity_update", g, ptr(w2), ptr(self.tau_z), ptr(self.tau_r), 0.0, sp(1), s)
        if not self.overlap_ls:
            tether()
        _call("axb_penalise_update_vorticity", g, ptr(self.u_z), ptr(self.u_r), ptr(w2), ptr(self.u_z_upen),
              ptr(self.u_r_upen), ptr(self.tether_char_func), self.brink_lam, 0.0, sp(1), 0.0, 0.0, sp(4), ptr(self.r1d),
              None, s)
        _call("axb_diffusion_rk2_stage1", g, ptr(self._tmp), ptr(w2), ptr(self.r1d), self.nu, 0.0, sp(1), s)
        _call("axb_diffusion_rk2_stage2", g, ptr(w), ptr(w2), ptr(self._tmp), ptr(self.r1d), self.nu, 0.0, sp(1), s)
        scalars(2)

    def sync_scalars(self):
        """device mode: one synchronising D2H of the loop scalars into the attributes the host mode keeps; raises if
        the LS extrapolation ran out of workspace or sweeps"""
        if not self.device_scalars:
            return self
        st = self.state.cpu().numpy()
        self.t, self.dt, self.freqTimer, self.cycles, self.it = st[0], st[1], st[3], int(st[7]), int(st[9])
        status, self.ls_sweeps = (int(v) for v in self._ls_status.cpu().numpy())
        if status:
            raise _lib.AxbError(f"LS extrapolation (device loop): status {status} "
                                "(1: pending list overflow, 2: sweep budget exhausted)")
        return self

    def push_scalars(self):
        """device mode: the host attributes (e.g. after io.load_restart) -> the device block"""
        if self.device_scalars:
            st = self.state.cpu()
            st[0], st[3], st[7], st[9] = self.t, self.freqTimer, float(self.cycles), float(self.it)
            st[2], st[8] = 0.0, 0.0
            self.state.copy_(st)
        return self

    def _one(self):
        F, s, g = self.F, stream_ptr(), self.F.g
        dx, w = self.dx, self.vorticity
        _call("axb_kill_boundary_vorticity_sine_z", g, ptr(w), ptr(self.z1d), 3, s)
        _call("axb_kill_boundary_vorticity_sine_r", g, ptr(w), ptr(self.r1d), 3, s)
        _lib.call("axb_fd_solve", ctypes.byref(self.solver.plan), ptr(self.psi), self.nz, ptr(w), self.nz, s)
        self._umax.zero_()
        _call("axb_velocity_from_psi", g, ptr(self.u_z_upen), ptr(self.u_r_upen), ptr(self.psi), ptr(self.r1d), 0.0, 0.0,
              None, ptr(self._umax), s)
        umax = float(self._umax)                       # the reference's np.amax round trip
        dt = min(self.CFL * dx / np.sqrt(self.G / self.rho_f), self.CFL * dx / (umax + self.eps),
                 0.9 * dx ** 2 / 4 / self.nu)
        if self.freqTimer + dt > self.freqTimer_limit:
            dt = self.freqTimer_limit - self.freqTimer
        if self.t + dt > self.tEnd:
            dt = self.tEnd - self.t
        self.dt = dt
        _call("axb_axpy", g, ptr(self.avg_psi), ptr(self.psi), dt, None, s)
        _call("axb_axpy", g, ptr(self.avg_phi), ptr(self.ball_phi), dt, None, s)
        _call("axb_advect_refmap_eno3", g, ptr(self._e1b), ptr(self._e2b), ptr(self.eta1), ptr(self.eta2),
              ptr(self.u_z_upen), ptr(self.u_r_upen), dt, None, s)
        self.eta1, self._e1b = self._e1b, self.eta1
        self.eta2, self._e2b = self._e2b, self.eta2
        _call("axb_pin_level_set", g, ptr(self.ball_phi), None, ptr(self.eta1), ptr(self.eta2), self.Z_cm, self.R_cm,
              self.r_ball, -3 * dx, s)
        if self._reinit is not None:
            # reinit level set (soft_sphere_streaming.py:195-199): unreached cells keep their value
            self._reinit(self.ball_phi, dx, self.extrap_zone)
        _call("axb_advect_vorticity_eno3", g, ptr(self._w2), ptr(w), ptr(self.u_z_upen), ptr(self.u_r_upen), dt, None, s)
        self.vorticity, self._w2 = self._w2, self.vorticity
        w = self.vorticity
        _call("axb_smooth_heaviside_mask", g, ptr(self.ball_char_func), ptr(self.inside_solid), ptr(self.ball_phi),
              self.moll_zone, 0.5, 0, s)
        sweeps = ctypes.c_int(0)
        _call("axb_ls_extrapolate_eta", g, ptr(self.ball_phi), ptr(self.inside_solid), ptr(self.eta1), ptr(self.eta2),
              self.extrap_zone, ptr(self.z1d), ptr(self.gy), ptr(self._ls_work), self._ls_bytes, 0,
              ctypes.byref(sweeps), s)
        self.ls_sweeps = sweeps.value
        _call("axb_solid_sigma", g, ptr(self.s11), ptr(self.s12), ptr(self.s22), self.G, ptr(self.eta1), ptr(self.eta2),
              ptr(self.e1z), ptr(self.e1r), ptr(self.e2z), ptr(self.e2r), ptr(self.ball_char_func), s)
        _call("axb_solid_tau", g, ptr(self.tau_z), ptr(self.tau_r), ptr(self.s11), ptr(self.s12), ptr(self.s22),
              ptr(self.r1d), s)
        _call("axb_solid_vorticity_update", g, ptr(w), ptr(self.tau_z), ptr(self.tau_r), dt, None, s)
        Z_cm_t = self.Z_cm + self.e * self.r_ball * np.sin(self.omega * self.t)
        _call("axb_smooth_heaviside_sphere", g, ptr(self.tether_char_func), None, ptr(self.z1d), ptr(self.r1d), Z_cm_t,
              self.R_cm, self.fixed_rad, self.moll_zone, s)
        _call("axb_penalise_update_vorticity", g, ptr(self.u_z), ptr(self.u_r), ptr(w), ptr(self.u_z_upen),
              ptr(self.u_r_upen), ptr(self.tether_char_func), self.brink_lam, dt, None,
              self.U_0 * np.cos(self.omega * self.t), 0.0, None, ptr(self.r1d), None, s)
        _call("axb_diffusion_rk2_stage1", g, ptr(self._tmp), ptr(w), ptr(self.r1d), self.nu, dt, None, s)
        _call("axb_diffusion_rk2_stage2", g, ptr(w), ptr(w), ptr(self._tmp), ptr(self.r1d), self.nu, dt, None, s)
        self.t += dt
        self.freqTimer += dt
        if self.freqTimer >= self.freqTimer_limit:
            # end of an oscillation cycle (soft_sphere_streaming.py:139-165): the driver's hook sees the completed
            # cycle averages (the reference plots them here), then they start again from zero
            self.freqTimer = 0.0
            self.cycles += 1
            if self.on_cycle is not None:
                self.on_cycle(self)
            _call("axb_set_fixed_val", g, ptr(self.avg_psi), 0.0, s)
            _call("axb_set_fixed_val", g, ptr(self.avg_phi), 0.0, s)
        self.it += 1


class ParticleFlowStepper:
    """Loop body of ``examples/ParticleOscillatoryFlowCases/particle_in_bubble_oscillatory_flow.py:159-358``
    (one member of the config-C5 ensemble): a free rigid particle next to an oscillating bubble,
    remeshed-vortex (MP4 particle) advection.  ``freq`` and ``e`` are the swept parameters.
    Members of an ensemble on one GPU share the solver factors (``solver=``)."""

    def __init__(self, grid_size_z=400, domain_AR=0.5, grid_size_r=None, freq=8.0, e=0.01, lambda_part=20.0,
                 r0_bubble=0.25, rp=2.0, brink_lam=1e12, CFL=0.1, rho_f=1.0, rho_s=1.0, solver=None, basis="auto",
                 device_scalars=False, use_graph=False, trace_capacity=4096, _pool=None, _member=0):
        if not torch.cuda.is_available():
            raise _lib.AxbError("ParticleFlowStepper needs a CUDA device (no CPU fallback)")
        nz = int(grid_size_z)
        nr = int(grid_size_r) if grid_size_r is not None else int(domain_AR * nz)
        dx = 1.0 / nz
        self.F = F = _FieldSet(nr, nz, dx, _pool, _member)
        self.nr, self.nz, self.dx = nr, nz, dx
        self.CFL, self.brink_lam, self.rho_f, self.rho_s = CFL, brink_lam, rho_f, rho_s
        self.moll_zone = np.sqrt(2) * dx
        self.freqTimer_limit = 1 / freq
        self.omega = 2 * np.pi * freq
        self.r0_bubble = r0_bubble
        self.r_part = 0.2 * r0_bubble
        self.nu = self.r_part ** 2 * self.omega / 3.0 / lambda_part
        self.U_0 = e * r0_bubble * self.omega
        self.eps = np.finfo(float).eps
        z = np.linspace(0 + dx / 2, 1 - dx / 2, nz)
        r = np.linspace(0 + dx / 2, nr * dx - dx / 2, nr)
        self.z1d, self.r1d = torch.from_numpy(z).cuda(), torch.from_numpy(r).cuda()
        self.rl_double = torch.from_numpy(np.linspace(dx / 2, 2 * nr * dx - dx / 2, 2 * nr)).cuda()
        self.bubble_Z_cm, self.bubble_R_cm = 0.5 - rp * r0_bubble, 0.0
        self.part_Z_cm, self.part_R_cm = self.bubble_Z_cm + rp * r0_bubble, 0.0
        (self.vorticity, self.psi, self.u_z, self.u_r, self.u_z_upen, self.u_r_upen, self._tmp, self._w2,
         self.bubble_char_func, self.part_char_func, self.avg_psi, self.avg_vort, self.avg_part_char_func) = F.new(13)
        s = stream_ptr()
        _call("axb_smooth_heaviside_sphere", F.g, ptr(self.bubble_char_func), None, ptr(self.z1d), ptr(self.r1d),
              self.bubble_Z_cm, self.bubble_R_cm, r0_bubble, self.moll_zone, s)
        _call("axb_smooth_heaviside_sphere", F.g, ptr(self.part_char_func), None, ptr(self.z1d), ptr(self.r1d),
              self.part_Z_cm, self.part_R_cm, self.r_part, self.moll_zone, s)
        self._acc = torch.zeros(2, dtype=torch.float64, device="cuda")
        self._far_flag = torch.zeros(1, dtype=torch.int32, device="cuda")   # remesh: set when a particle moved >= a cell
        ones = F.new(1) if _pool is not None else torch.empty((nr, nz), dtype=torch.float64, device="cuda")
        ones.fill_(1.0)
        _call("axb_reduce_weighted_sum", F.g, ptr(self.r1d), ptr(self.part_char_func), ptr(ones), 0.0, ptr(self._acc), s)
        self.part_vol = float(self._acc[0])          # np.sum(part_char_func * R)
        self.part_mass = rho_s * self.part_vol
        del ones
        self.solver = solver if solver is not None else FastDiagonalisationStokesSolver(nr, nz, dx, basis=basis)
        # device-resident loop scalars (include/axisym_b200.h, axb_particle_scalars): no host round trip per step
        self.device_scalars = bool(device_scalars)
        self._use_graph = bool(use_graph) and self.device_scalars
        self._graph = None
        self.graph_launches, self.launches_replayed = 0, 0
        if self.device_scalars:
            self.state = torch.zeros(24, dtype=torch.float64, device="cuda")
            self.state[6] = self.part_Z_cm
            self.trace_dev = torch.zeros((int(trace_capacity), 5), dtype=torch.float64, device="cuda")
            self.avg_psi_last, self.avg_vort_last, self.avg_part_char_func_last = F.new(3)
            # the bubble does not move: (d^2)^1.5 and the inside flag once (axb_bubble_flow_geometry)
            self.bubble_geom = torch.empty((nr, nz), dtype=torch.float64, device="cuda")
            gg = make_grid(nr, nz, nz, dx)
            chi_b = self.bubble_char_func.contiguous()
            _call("axb_bubble_flow_geometry", ctypes.byref(gg), ptr(self.bubble_geom), ptr(chi_b), ptr(self.z1d),
                  ptr(self.r1d), self.bubble_Z_cm, self.bubble_R_cm, s)
            torch.cuda.current_stream().synchronize()
        self.t, self.it, self.dt = 0.0, 0, 0.0
        # per-cycle averages (particle_in_bubble_oscillatory_flow.py:102-109, 168-263, 297-301, 355)
        self.freqTimer, self.avg_Z_cm, self.avg_time = 0.0, 0.0, 0.0
        self.cycles, self.on_cycle = 0, None      # on_cycle(stepper): called with the completed-cycle averages
        self.avg_T, self.avg_part_trajectory = [], []
        self.U_z_cm_part, self.diff = 0.0, 0.0
        self.F_total = 0.0
        self.trace = []       # per step: (t, dt, U_z_cm_part, part_Z_cm, F_total) at the start of the step

    def step(self, n=1):
        if self.device_scalars:
            if not self._use_graph:
                for _ in range(n):
                    self._one_dev()
                return
            if self._graph is None:
                self._graph = _capture_step(self, self._one_dev)
                n -= 1
            for _ in range(n):
                self._graph.replay()
                self.launches_replayed += self.graph_launches
            return
        for _ in range(n):
            self._one()

    # ---- device-resident form: the same kernels, every host decision replaced by axb_particle_scalars ------------
    def _sp(self, i):
        return ctypes.c_void_p(self.state.data_ptr() + 8 * i)

    def _scalars_dev(self, phase):
        dx = self.dx
        _call("axb_particle_scalars", phase, ptr(self.state), ptr(self.trace_dev), self.trace_dev.shape[0],
              0.9 * dx ** 2 / 4 / self.nu, self.CFL, self.eps, self.freqTimer_limit, self.omega,
              self.rho_f * self.brink_lam, self.part_vol, self.part_mass, self.bubble_Z_cm, self.r0_bubble, stream_ptr())

    def _enqueue_dev_pre(self):
        s, g = stream_ptr(), self.F.g
        w = self.vorticity
        _call("axb_kill_boundary_vorticity_sine_z", g, ptr(w), ptr(self.z1d), 3, s)
        _call("axb_kill_boundary_vorticity_sine_r", g, ptr(w), ptr(self.r1d), 3, s)

    def _enqueue_dev_solve(self):
        ld = self.vorticity.stride(0)
        _lib.call("axb_fd_solve", ctypes.byref(self.solver.plan), ptr(self.psi), ld, ptr(self.vorticity), ld, stream_ptr())

    def _enqueue_dev_post(self):
        s, g, sp = stream_ptr(), self.F.g, self._sp
        w = self.vorticity
        _call("axb_velocity_from_psi", g, ptr(self.u_z_upen), ptr(self.u_r_upen), ptr(self.psi), ptr(self.r1d), 0.0, 0.0,
              None, None, s)
        _call("axb_reduce_max_abs_sum", g, ptr(w), None, sp(2), s)         # state[2] was zeroed by phase 2
        self._scalars_dev(1)
        _call("axb_add_bubble_flow_geom", g, ptr(self.u_z_upen), ptr(self.u_r_upen), ptr(self.bubble_geom), self.nz,
              ptr(self.z1d), ptr(self.r1d), self.bubble_Z_cm, self.bubble_R_cm, self.r0_bubble, self.U_0, None, sp(9), s)
        _call("axb_cycle_average3", g, ptr(self.avg_part_char_func), ptr(self.part_char_func),
              ptr(self.avg_part_char_func_last), ptr(self.avg_psi), ptr(self.psi), ptr(self.avg_psi_last),
              ptr(self.avg_vort), ptr(w), ptr(self.avg_vort_last), sp(10), sp(15), s)
        _call("axb_smooth_heaviside_sphere_dev", g, ptr(self.part_char_func), None, ptr(self.z1d), ptr(self.r1d), sp(6),
              self.part_R_cm, self.r_part, self.moll_zone, s)
        _call("axb_penalise_update_vorticity", g, ptr(self.u_z), ptr(self.u_r), ptr(w), ptr(self.u_z_upen),
              ptr(self.u_r_upen), ptr(self.part_char_func), self.brink_lam, 0.0, sp(1), 0.0, 0.0, sp(4), ptr(self.r1d),
              sp(3), s)
        _call("axb_advect_vorticity_particles_flagged", g, ptr(self._w2), ptr(w), ptr(self.u_z), ptr(self.u_r),
              ptr(self.z1d), ptr(self.rl_double), 0.0, sp(1), ptr(self._far_flag), s)
        # RK2 diffusion lands the result back in `vorticity` (stage 2 takes its base field from _w2): no buffer swap,
        # so the launch sequence is the same every step and can be replayed as a CUDA graph
        _call("axb_diffusion_rk2_stage1", g, ptr(self._tmp), ptr(self._w2), ptr(self.r1d), self.nu, 0.0, sp(1), s)
        _call("axb_diffusion_rk2_stage2", g, ptr(w), ptr(self._w2), ptr(self._tmp), ptr(self.r1d), self.nu, 0.0, sp(1), s)
        self._scalars_dev(2)

    def _one_dev(self):
        self._enqueue_dev_pre()
        self._enqueue_dev_solve()
        self._enqueue_dev_post()

    def sync_scalars(self):
        """device mode: copy the loop scalars (and the trace rows written so far) to the host attributes the host
        mode keeps (one synchronising D2H; call it when a value is wanted, at least once per oscillation cycle if the
        per-cycle lists ``avg_T`` / ``avg_part_trajectory`` are wanted complete)"""
        if not self.device_scalars:
            return self
        st = self.state.cpu().numpy()
        self.t, self.dt, self.U_z_cm_part, self.part_Z_cm, self.F_total = st[0], st[1], st[4], st[6], st[7]
        self.it, self.freqTimer, self.avg_Z_cm, self.avg_time, self.diff = int(st[8]), st[11], st[12], st[13], st[16]
        if int(st[14]) > self.cycles:
            self.cycles = int(st[14])
            self.avg_T.append(float(st[17]))
            self.avg_part_trajectory.append(float(st[18]))
        cap = self.trace_dev.shape[0]
        rows = self.trace_dev[:min(self.it, cap)].cpu().numpy()
        if self.it > cap:                            # the ring has wrapped: oldest kept row first
            rows = np.roll(rows, -(self.it % cap), axis=0)
        self.trace = [tuple(r) for r in rows]
        return self

    def push_scalars(self):
        """device mode: the host attributes (e.g. after io.load_restart) -> the device block"""
        if self.device_scalars:
            st = self.state.cpu()
            for i, v in ((0, self.t), (4, self.U_z_cm_part), (6, self.part_Z_cm), (7, self.F_total), (8, float(self.it)),
                         (11, self.freqTimer), (12, self.avg_Z_cm), (13, self.avg_time), (14, float(self.cycles)),
                         (16, self.diff), (2, 0.0), (3, 0.0), (15, 0.0)):
                st[i] = v
            self.state.copy_(st)
        return self

    # One step = three pieces, so that an ensemble can interleave its members (ParticleEnsemble): everything up to
    # the first host decision (dt needs the vorticity maximum), everything that needs dt, and the rigid-body update
    # on the host (needs the penalisation force sum).  `_acc` = [max |w|, sum R chi (u_z - U)].
    def _enqueue_a(self):
        s, g = stream_ptr(), self.F.g
        w = self.vorticity
        _call("axb_kill_boundary_vorticity_sine_z", g, ptr(w), ptr(self.z1d), 3, s)
        _call("axb_kill_boundary_vorticity_sine_r", g, ptr(w), ptr(self.r1d), 3, s)
        _lib.call("axb_fd_solve", ctypes.byref(self.solver.plan), ptr(self.psi), self.nz, ptr(w), self.nz, s)
        _call("axb_velocity_from_psi", g, ptr(self.u_z_upen), ptr(self.u_r_upen), ptr(self.psi), ptr(self.r1d), 0.0, 0.0,
              None, None, s)
        _call("axb_fill_scalars", ptr(self._acc), 1, 0.0, s)
        _call("axb_reduce_max_abs_sum", g, ptr(w), None, ptr(self._acc), s)

    def _enqueue_b(self, wmax):
        s, g, dx = stream_ptr(), self.F.g, self.dx
        w = self.vorticity
        if self.freqTimer >= self.freqTimer_limit:
            # a cycle is complete (particle_in_bubble_oscillatory_flow.py:168-263): hand the averages to the driver's
            # hook (the reference dumps them here), record the cycle-mean trajectory point, start again from zero
            self.freqTimer = 0.0
            self.cycles += 1
            cycle_time = self.freqTimer_limit
            self.avg_T.append(self.avg_time / cycle_time)
            self.avg_part_trajectory.append((self.avg_Z_cm / cycle_time - self.bubble_Z_cm) / self.r0_bubble)
            if self.on_cycle is not None:
                self.on_cycle(self)
            for f in (self.avg_part_char_func, self.avg_psi, self.avg_vort):
                _call("axb_set_fixed_val", g, ptr(f), 0.0, s)
            self.avg_Z_cm, self.avg_time = 0.0, 0.0
        dt = min(0.9 * dx ** 2 / 4 / self.nu, self.CFL / (wmax + self.eps), 0.01 * self.freqTimer_limit)
        self.dt = dt
        _call("axb_add_bubble_flow", g, ptr(self.u_z_upen), ptr(self.u_r_upen), ptr(self.bubble_char_func),
              ptr(self.z1d), ptr(self.r1d), self.bubble_Z_cm, self.bubble_R_cm, self.r0_bubble, self.U_0,
              np.sin(self.omega * self.t), s)
        a = dt / self.freqTimer_limit
        _call("axb_axpy", g, ptr(self.avg_part_char_func), ptr(self.part_char_func), a, None, s)
        _call("axb_axpy", g, ptr(self.avg_psi), ptr(self.psi), a, None, s)
        _call("axb_axpy", g, ptr(self.avg_vort), ptr(w), a, None, s)
        self.avg_Z_cm += self.part_Z_cm * dt
        self.avg_time += self.t * dt
        self.freqTimer += dt
        _call("axb_smooth_heaviside_sphere", g, ptr(self.part_char_func), None, ptr(self.z1d), ptr(self.r1d),
              self.part_Z_cm, self.part_R_cm, self.r_part, self.moll_zone, s)
        sum_ptr = ctypes.c_void_p(self._acc.data_ptr() + 8)
        _call("axb_fill_scalars", sum_ptr, 1, 0.0, s)
        _call("axb_penalise_update_vorticity", g, ptr(self.u_z), ptr(self.u_r), ptr(w), ptr(self.u_z_upen),
              ptr(self.u_r_upen), ptr(self.part_char_func), self.brink_lam, dt, None, self.U_z_cm_part, 0.0, None,
              ptr(self.r1d), sum_ptr, s)
        _call("axb_advect_vorticity_particles_flagged", g, ptr(self._w2), ptr(w), ptr(self.u_z), ptr(self.u_r),
              ptr(self.z1d), ptr(self.rl_double), dt, None, ptr(self._far_flag), s)
        self.vorticity, self._w2 = self._w2, self.vorticity
        w = self.vorticity
        _call("axb_diffusion_rk2_stage1", g, ptr(self._tmp), ptr(w), ptr(self.r1d), self.nu, dt, None, s)
        _call("axb_diffusion_rk2_stage2", g, ptr(w), ptr(w), ptr(self._tmp), ptr(self.r1d), self.nu, dt, None, s)

    def _finish(self, pen_sum):
        # rigid-body update on the host (compute_forces.py:4-17, particle_in_bubble_oscillatory_flow.py:323-351)
        dt = self.dt
        F_pen = self.rho_f * self.brink_lam * pen_sum
        F_un = (self.diff * self.part_vol) / dt
        F_total = F_pen + F_un
        self.F_total = F_total
        self.trace.append((self.t, dt, self.U_z_cm_part, self.part_Z_cm, F_total))
        U_old = self.U_z_cm_part
        self.U_z_cm_part += 0.5 * dt * (self.diff / dt + (F_total / self.part_mass))
        self.diff = dt * F_total / self.part_mass
        self.part_Z_cm += U_old * dt + (0.5 * dt * dt * F_total / self.part_mass)
        self.t += dt
        self.it += 1

    def _one(self):
        self._enqueue_a()
        self._enqueue_b(float(self._acc[0]))
        self._finish(float(self._acc[1]))


class _BatchedFdSolver:
    """One fast-diagonalisation solve for all members of a :class:`_WidePool` ensemble (cosine-transform z path +
    tridiagonal r path): DCT-II over batch nr rows, factored sweeps over batch nz columns (the pivots of the nz
    z-modes tiled batch times), DCT-III -- 3 + 1 launches whatever the ensemble size, and the r sweeps, which march
    row by row and are latency bound on a single 2048-column member, get batch times the columns."""

    def __init__(self, nr, nz, batch, factors):
        f, tri = factors, factors["tri"]
        self.nr, self.nz, self.batch, self.tables = nr, nz, batch, f["zfft"]["tables"]
        dev = f["lam_z"].device
        lam = f["lam_z"].repeat(batch).contiguous()
        cols = batch * nz
        self.inv = torch.empty((nr, cols), dtype=torch.float64, device=dev)
        self.rc = torch.empty((nr, 4), dtype=torch.float64, device=dev)
        _call("axb_tridiag_factor_columns", nr, cols, ptr(tri["sub"]), ptr(tri["diag"]), ptr(tri["sup"]), ptr(lam),
              ptr(tri["scale"]), float(f["c0"]), float(f["c1"]), ptr(self.inv), ptr(self.rc), stream_ptr())
        self.spec = torch.empty((nr, cols), dtype=torch.float64, device=dev)

    def solve(self, psi_wide, rhs_wide):
        nr, nz, b = self.nr, self.nz, self.batch
        s = stream_ptr()
        rows = b * nr
        _call("axb_dct2_rows", rows, nz, ptr(rhs_wide), nz, ptr(self.spec), nz, ptr(self.tables), 1.0 / nz, 2.0 / nz, s)
        _call("axb_tridiag_solve_factored", nr, b * nz, ptr(self.spec), b * nz, ptr(self.inv), ptr(self.rc), s)
        _call("axb_dct3_rows", rows, nz, ptr(self.spec), nz, ptr(psi_wide), nz, ptr(self.tables), s)


class ParticleEnsemble:
    """Config C5 (SURVEY.md 8e "Ensemble"): independent ``ParticleFlowStepper`` members on one GPU, no communication.

    The members' kernels are small (a 1024 x 2048 field is 16 MiB) and each member has two host decisions per step
    (dt from the vorticity maximum, the rigid-body update from the penalisation force), so run one after the other
    they leave the GPU idle most of the time.  Here every member has its own stream and its own solve work fields
    (``solver.fork()``: same factors), the pieces of a step are enqueued member after member, and the force of
    step n is read together with the vorticity maximum of step n+1 -- one host read per member and step, while
    the other members' kernels run.  Each member performs exactly the operations of ``ParticleFlowStepper.step``.
    """

    def __init__(self, members):
        self.members = list(members)
        self.batched = False
        first = self.members[0].solver
        for i, m in enumerate(self.members):
            if i > 0 and m.solver is first:
                m.solver = first.fork()
        self.streams = [torch.cuda.Stream() for _ in self.members]
        self._pending = [False] * len(self.members)

    @classmethod
    def batched_ensemble(cls, params, grid_size_z, grid_size_r=None, use_graph=True, branches=4, solver=None,
                         launch="batched", **kw):
        """SURVEY.md 8e "Ensemble": ``params`` = [(freq, e), ...]; all members share one set of (nr, batch nz) field
        tensors (:class:`_WidePool`) and keep their loop scalars in one (batch, 24) device block; a step of the WHOLE
        ensemble is a fixed launch sequence captured once and replayed -- no host round trip.  ``launch="batched"``:
        every operation is ONE launch over all members (``axb_grid_t.batch``: the member is the grid's z dimension,
        per-member dt / U / omega / nu / ... read from the scalar block), 14 launches + the 4 of the solve whatever
        the ensemble size.  ``launch="members"``: one launch per member and operation on `branches` parallel graph
        branches around the shared solve (the cross-check of the batched kernels)."""
        nz = int(grid_size_z)
        nr = int(grid_size_r) if grid_size_r is not None else int(kw.get("domain_AR", 0.5) * nz)
        pool = _WidePool(nr, nz, len(params))
        members = []
        for i, (f, e) in enumerate(params):
            m = ParticleFlowStepper(nz, grid_size_r=nr, freq=f, e=e, solver=solver, device_scalars=True,
                                    _pool=pool, _member=i, **kw)
            solver = m.solver
            members.append(m)
        self = cls.__new__(cls)
        self.members, self.batched, self.pool = members, True, pool
        self._use_graph, self._graph, self._branches = bool(use_graph), None, max(1, int(branches))
        self.graph_launches, self.launches_replayed = 0, 0
        self._solver = _BatchedFdSolver(nr, nz, len(params), solver.factors) if solver.factors.get("zfft") is not None \
            and solver.factors.get("tri") is not None else None
        self._w_wide = pool.wide[0]        # field 0 = vorticity, field 1 = psi (ParticleFlowStepper's F.new order)
        self._psi_wide = pool.wide[1]
        self._side = [torch.cuda.Stream() for _ in range(self._branches)]
        # one scalar block / trace ring for the ensemble; the members keep views of their rows
        b, m0 = len(members), members[0]
        self.state = torch.zeros((b, 24), dtype=torch.float64, device="cuda")
        self.trace_dev = torch.zeros((b,) + tuple(m0.trace_dev.shape), dtype=torch.float64, device="cuda")
        for i, m in enumerate(members):
            self.state[i].copy_(m.state)
            self.state[i, 19:24] = torch.tensor([m.omega, m.freqTimer_limit, m.U_0, m.nu, 0.9 * m.dx ** 2 / 4 / m.nu],
                                                dtype=torch.float64)
            m.state, m.trace_dev = self.state[i], self.trace_dev[i]
            m.bubble_geom = m0.bubble_geom               # same bubble for every member
            if abs(m.part_vol - m0.part_vol) > 1e-14 * m0.part_vol or (m.bubble_Z_cm, m.r0_bubble, m.r_part) != \
                    (m0.bubble_Z_cm, m0.r0_bubble, m0.r_part):
                raise ValueError("the members of a batched ensemble differ in (freq, e) only")
        if launch not in ("batched", "members"):
            raise ValueError(f"launch {launch!r}")
        self._launch = launch
        self._grid_b = make_grid(nr, nz, b * nz, m0.dx, batch=(b, nz, 24))
        return self

    def _enqueue_one_launch_per_op(self):
        """the step of every member, each operation one launch (grid z dimension = member)"""
        m, b = self.members[0], len(self.members)
        g, s = ctypes.byref(self._grid_b), stream_ptr()
        base = self.state.data_ptr()

        def sp(i):
            return ctypes.c_void_p(base + 8 * i)

        def scalars(phase):
            _call("axb_particle_scalars_batched", phase, b, 24, ptr(self.state), ptr(self.trace_dev),
                  self.trace_dev.shape[1], m.CFL, m.eps, m.rho_f * m.brink_lam, m.part_vol, m.part_mass, m.bubble_Z_cm,
                  m.r0_bubble, s)

        w = m.vorticity
        _call("axb_kill_boundary_vorticity_sine_z", g, ptr(w), ptr(m.z1d), 3, s)
        _call("axb_kill_boundary_vorticity_sine_r", g, ptr(w), ptr(m.r1d), 3, s)
        if self._solver is not None:
            self._solver.solve(self._psi_wide, self._w_wide)
        else:
            for mm in self.members:
                mm._enqueue_dev_solve()
        _call("axb_velocity_from_psi", g, ptr(m.u_z_upen), ptr(m.u_r_upen), ptr(m.psi), ptr(m.r1d), 0.0, 0.0, None, None, s)
        _call("axb_reduce_max_abs_sum", g, ptr(w), None, sp(2), s)
        scalars(1)
        _call("axb_add_bubble_flow_geom", g, ptr(m.u_z_upen), ptr(m.u_r_upen), ptr(m.bubble_geom), m.nz, ptr(m.z1d),
              ptr(m.r1d), m.bubble_Z_cm, m.bubble_R_cm, m.r0_bubble, 0.0, sp(21), sp(9), s)
        _call("axb_cycle_average3", g, ptr(m.avg_part_char_func), ptr(m.part_char_func), ptr(m.avg_part_char_func_last),
              ptr(m.avg_psi), ptr(m.psi), ptr(m.avg_psi_last), ptr(m.avg_vort), ptr(w), ptr(m.avg_vort_last), sp(10),
              sp(15), s)
        _call("axb_smooth_heaviside_sphere_dev", g, ptr(m.part_char_func), None, ptr(m.z1d), ptr(m.r1d), sp(6),
              m.part_R_cm, m.r_part, m.moll_zone, s)
        _call("axb_penalise_update_vorticity", g, ptr(m.u_z), ptr(m.u_r), ptr(w), ptr(m.u_z_upen), ptr(m.u_r_upen),
              ptr(m.part_char_func), m.brink_lam, 0.0, sp(1), 0.0, 0.0, sp(4), ptr(m.r1d), sp(3), s)
        _call("axb_advect_vorticity_particles_flagged", g, ptr(m._w2), ptr(w), ptr(m.u_z), ptr(m.u_r), ptr(m.z1d),
              ptr(m.rl_double), 0.0, sp(1), ptr(m._far_flag), s)
        _call("axb_diffusion_rk2_stage1_dev", g, ptr(m._tmp), ptr(m._w2), ptr(m.r1d), sp(22), sp(1), s)
        _call("axb_diffusion_rk2_stage2_dev", g, ptr(w), ptr(m._w2), ptr(m._tmp), ptr(m.r1d), sp(22), sp(1), s)
        scalars(2)

    def _enqueue_batched(self):
        if self._launch == "batched":
            return self._enqueue_one_launch_per_op()
        for m in self.members:
            m._enqueue_dev_pre()
        if self._solver is not None:
            self._solver.solve(self._psi_wide, self._w_wide)
        else:
            for m in self.members:
                m._enqueue_dev_solve()
        cur = torch.cuda.current_stream()
        if self._branches == 1:
            for m in self.members:
                m._enqueue_dev_post()
            return
        for st in self._side:
            st.wait_stream(cur)
        for i, m in enumerate(self.members):
            with torch.cuda.stream(self._side[i % self._branches]):
                m._enqueue_dev_post()
        for st in self._side:
            cur.wait_stream(st)

    def sync_scalars(self):
        for m in self.members:
            m.sync_scalars()
        return self

    def step(self, n=1):
        if self.batched:
            if not self._use_graph:
                for _ in range(n):
                    self._enqueue_batched()
                return
            if self._graph is None:
                self._graph = _capture_step(self, self._enqueue_batched)
                n -= 1
            for _ in range(n):
                self._graph.replay()
                self.launches_replayed += self.graph_launches
            return
        cur = torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(cur)
        for _ in range(n):
            for m, st in zip(self.members, self.streams):
                with torch.cuda.stream(st):
                    m._enqueue_a()
            for i, (m, st) in enumerate(zip(self.members, self.streams)):
                with torch.cuda.stream(st):
                    vals = m._acc.cpu()                 # waits for this member's stream only
                    if self._pending[i]:
                        m._finish(float(vals[1]))
                    m._enqueue_b(float(vals[0]))
                    self._pending[i] = True
        for i, (m, st) in enumerate(zip(self.members, self.streams)):
            with torch.cuda.stream(st):
                m._finish(float(m._acc[1]))
            self._pending[i] = False
            cur.wait_stream(st)
