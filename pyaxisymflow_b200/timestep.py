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
    Launch sequence of one step (18 launches):
      scalars(0) -> kill_z -> kill_r -> 4 x DGEMM (psi) -> velocity(+U, max) -> scalars(1: dt)
      -> penalise+curl+drag -> ENO3 advect (w -> w2) -> RK2 stage 1 (w2 -> tmp)
      -> RK2 stage 2 (w2, tmp -> w) -> scalars(2: t += dt)
    """

    def __init__(self, grid_size_z, domain_AR=0.5, Re=100.0, U_0=1.0, r_sph=0.1, Z_cm=0.25, R_cm=0.0,
                 brink_lam=1e12, CFL=0.1, T_ramp=None, periodic=False, ghost_size=2, basis="auto",
                 grid_size_r=None, use_graph=False):
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
                nr, nz - 2 * g, dx, bc_type="homogenous_neumann_along_r_and_periodic_along_z", basis=basis)
        else:
            self.solver = FastDiagonalisationStokesSolver(nr, nz, dx, basis=basis)
        if self.periodic:
            _call("axb_periodic_ghost_comm", ctypes.byref(self.grid), ptr(self.char_func), self.ghost, 0.0, 0.0,
                  stream_ptr())
        self._graph = None
        self._use_graph = use_graph

    # -- one step, enqueued on the current stream ---------------------------------------------
    def _enqueue(self, probe=None):
        s = stream_ptr()
        g = ctypes.byref(self.grid)
        st = self.state
        sp = lambda i: ctypes.c_void_p(st.data_ptr() + 8 * i)  # noqa: E731
        w, psi = self.vorticity, self.psi
        _call("axb_rigid_flow_scalars", 0, ptr(st), self.U_0, self.T_ramp, self.ur_ramp, self.dt_diff_limit, self.CFL * self.dx, s)
        if not self.periodic:
            _call("axb_kill_boundary_vorticity_sine_z", g, ptr(w), ptr(self.z1d), 3, s)
        _call("axb_kill_boundary_vorticity_sine_r", g, ptr(w), ptr(self.r1d), 3, s)
        if probe is not None:
            probe[0].record()
        if self.periodic:
            gh = self.ghost
            off = 8 * gh
            _lib.call("axb_fd_solve", ctypes.byref(self.solver.plan), ctypes.c_void_p(psi.data_ptr() + off), self.nz,
                      ctypes.c_void_p(w.data_ptr() + off), self.nz, s)
            _call("axb_periodic_ghost_comm", g, ptr(psi), gh, 0.0, 0.0, s)
        else:
            _lib.call("axb_fd_solve", ctypes.byref(self.solver.plan), ptr(psi), self.nz, ptr(w), self.nz, s)
        if probe is not None:
            probe[1].record()
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
        _call("axb_diffusion_rk2_stage1", g, ptr(self._tmp), ptr(self._w2), ptr(self.r1d), self.nu, 0.0, sp(S_DT), s)
        if self.periodic:
            _call("axb_periodic_ghost_comm", g, ptr(self._tmp), self.ghost, 0.0, 0.0, s)
        _call("axb_diffusion_rk2_stage2", g, ptr(w), ptr(self._w2), ptr(self._tmp), ptr(self.r1d), self.nu, 0.0,
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
                with torch.cuda.graph(self._graph, stream=side):
                    self._enqueue()
                n -= 1
            for _ in range(n):
                self._graph.replay()
        else:
            for _ in range(n):
                self._enqueue()

    def step_probed(self):
        """one step with CUDA events around the four GEMMs of the solve (bench.py's roofline leg)"""
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._enqueue(probe=ev)
        return ev

    def solve_flops(self):
        nr, nz = self.solver.grid_size_r, self.solver.grid_size_z
        return 4.0 * nr * nz * (nr + nz)

    def solver_basis(self):
        return self.solver.basis

    def seed_vorticity(self, seed=0, amplitude=1.0):
        """seeded band-limited blob (SURVEY.md 8d synthetic input (ii)): N(0,1) * exp(-((Z-1/2)^2+R^2)/0.02)"""
        gen = torch.Generator(device="cuda")
        gen.manual_seed(seed)
        noise = torch.randn((self.nr, self.nz), dtype=torch.float64, device="cuda", generator=gen)
        env = torch.exp(-((self.z1d[None, :] - 0.5) ** 2 + self.r1d[:, None] ** 2) / 0.02)
        self.vorticity.copy_(amplitude * noise * env)

    def step_host(self, vorticity_host, char_func_host, out_host):
        """End-to-end form for host-resident callers: pinned host fields in, one step, vorticity out.
        (bench.py's `e2e`: both copies are inside the timed region.)"""
        self.vorticity.copy_(vorticity_host, non_blocking=True)
        self.char_func.copy_(char_func_host, non_blocking=True)
        self.step(1)
        out_host.copy_(self.vorticity, non_blocking=True)

    # -- diagnostics ----------------------------------------------------------------------------
    def scalars(self):
        st = self.state.cpu().numpy()
        cd = 2 * 2 * np.pi * self.dx * self.dx * self.brink_lam * st[S_SUMCOPY] / (np.pi * self.r_sph ** 2)
        return {"t": st[S_T], "dt": st[S_DT], "umax": st[S_UMAX], "iterations": int(st[S_IT]), "Cd": cd}
