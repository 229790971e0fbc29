#!/usr/bin/env python
"""bench.py -- grid-point updates / s per full timestep (FP64), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W                   (our arm, one process per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W  (reference arm: the CPU path, rank 0 only)

Default workload (``--config c4``): one complete loop iteration of the rigid-flow driver
(examples/FlowPastSphere/flow_past_sphere.py:107-207: boundary damping, streamfunction solve, velocity, CFL
reduction, penalisation + drag, ENO3 advection, RK2 diffusion) on the synthetic 4096 x 16384 FP64 grid the metric
is quoted on (BASELINE.json configs[3]; it fits one GPU).  With one GPU the same line also carries, under
``"configs"``, the other BASELINE.json configurations (c1, c2, c3, c5), each with its own value, roofline object,
CPU baseline and end-to-end number (``--no-configs`` skips them).

Timing: CUDA events on the launching stream, max over ranks.  4096 x 16384 and 2048 x 8192 fields (512 / 128 MiB
each) exceed the 126 MB L2, so those steps need no flush; for the small grids (c1, c2, c5) every step is timed
with its own event pair and a 256 MiB buffer is rewritten between the steps (L2 flush outside the timed pairs).

The CPU legs (``cpu_baseline`` and ``--impl reference``) run the oracle port of the reference's algorithm -- four
OpenBLAS GEMMs with the eigenbases (numpy.linalg.multi_dot, all cores), C/OpenMP ports of the numba and
pystencils kernels (all cores; the reference runs the numba ones on one thread), NumPy driver glue -- on a
bounded sample of the workload: a band of rows times the whole z extent for the stencils and the z GEMMs and the
matching column block for the r GEMMs, so that every sample point costs what a point of the full step costs.
When the band is the whole grid (c1, c2) the step is a full, un-extrapolated reference step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv and any(a.startswith("--impl") for a in sys.argv):
    # the reference arm uses every host core; torchrun pins OMP_NUM_THREADS=1 for its workers, and the BLAS /
    # OpenMP runtimes read the variables when they are loaded (i.e. before `import numpy`)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

# the CPU legs mix OpenBLAS threads with OpenMP regions: spinning waiters of one runtime starve the other
# (measured here: 87 ms -> 3 ms per 128x256 step), so idle threads must sleep
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "grid-pt updates/s per full timestep (FP64)"
UNIT = "grid-pt updates/s"
L2_BYTES = 126e6
GRIDS = {"c4": (4096, 16384), "c1": (128, 256), "c2": (1024, 4096), "c3": (2048, 8192), "c5": (1024, 2048)}
WORKLOADS = {
    "c4": "rigid-flow timestep (FlowPastSphere loop body) at {nr}x{nz}",
    "c1": "rigid-flow timestep (FlowPastSphere CLI default) at {nr}x{nz}",
    "c2": "PeriodicFlowPastSphere loop body at {nr}x{nz} (periodic z, ghost 2)",
    "c3": "SoftSphereStreaming loop body at {nr}x{nz} (reference-map solid)",
    "c5": "ensemble of {cases} ParticleOscillatoryFlowCases per GPU at {nr}x{nz} (MP4 remeshing)",
}


def config_dict(name, nr, nz, world, cases_per_gpu, reinit=False):
    """the `config` object -- built by this one function for both arms, so the two lines describe the same job"""
    flush = needs_flush(name, nr, nz)
    slab = "z-slab" if os.environ.get("AXB_SLAB_Z") else "r-slab"
    par = "single GPU" if world == 1 else (f"{slab} x{world}" if name in ("c4", "c2") else f"{world} independent replicas")
    wl = WORKLOADS[name].format(nr=nr, nz=nz, cases=cases_per_gpu)
    if name == "c3":
        wl += ("; narrow-band level-set re-initialisation included" if reinit
               else "; skfmm re-initialisation excluded on both arms")
    return {"workload": wl, "name": name, "cases": cases_per_gpu * world if name == "c5" else 1, "grid": [nr, nz],
            "parallelism": par,
            "l2": ("a 256 MiB buffer is rewritten between the timed steps (L2 flush); every step has its own event "
                   "pair" if flush else "fields (%.0f MiB each) exceed the 126 MB L2; no flush needed"
                   % (nr * nz * 8 / 2 ** 20))}


def needs_flush(name, nr, nz):
    return nr * nz * 8 < L2_BYTES


# ----------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index=0):
        self.rows, self.proc, self.idx = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm_sorted = sorted(sm)                      # "under load" = the upper half of the samples
        return {"sm_mhz": float(np.median(sm_sorted[len(sm_sorted) // 2:])) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU legs: the reference's CPU path (oracle port) on a bounded sample of the same workload
# ----------------------------------------------------------------------------------------------
class CpuStepSample:
    """One loop iteration of a configured driver on the host, restricted to a band of ``rows`` grid rows (all of
    z) -- the whole grid when it is small enough.  Same kernel sequence as the reference's loop body:

      c1 / c4  examples/FlowPastSphere/flow_past_sphere.py:107-207
      c2       examples/PeriodicFlowPastSphere/periodic_flow_past_sphere.py:95-183
      c3       examples/SoftSphereStreaming/soft_sphere_streaming.py:129-274 (without skfmm, like the GPU arm)
      c5       examples/ParticleOscillatoryFlowCases/particle_in_bubble_oscillatory_flow.py:159-358 (one member)

    The solve is the reference's algorithm (FastDiagonalisationStokesSolver.py:137-156): S = Lr rhs Rz, S *= 1/lam,
    psi = Lrb S Rzb with the dense eigenbases (closed-form bases built by the solver's host set-up; la.eig at
    Nz = 16384 would take tens of minutes and gives the same matrices up to rounding), evaluated by threaded
    BLAS.  On a band the two z products run on the band's rows with the full Nz x Nz matrices, the two r products
    with the full Nr x Nr matrices on as many columns as make up the band's point count -- per-point cost of the
    full products."""

    TARGET_POINTS = 1 << 22

    def __init__(self, name, nr, nz):
        from oracle import axisym_oracle as ox
        from pyaxisymflow_b200 import fd          # host-side set-up only (closed-form eigenbases)

        self.ox, self.name, self.nr, self.nz = ox, name, nr, nz
        rows = nr if nr * nz <= self.TARGET_POINTS else max(64, self.TARGET_POINTS // nz)
        self.rows = rows = min(rows, nr)
        self.points = rows * nz
        self.full = rows == nr
        self.cores = os.cpu_count() or 1
        dx = self.dx = 1.0 / nz
        periodic = name == "c2"
        self.g = 2 if periodic else 0
        bc = ("homogenous_neumann_along_r_and_periodic_along_z" if periodic else "homogenous_neumann_along_z_and_r")
        nzs = nz - 2 * self.g
        f = fd.build_factors("stokes", bc, nr, nzs, dx, basis="analytic", device="cpu", split=0, r_method="eigen",
                             z_method="gemm")
        self.Lr, self.Lrb = f["Lr"].numpy(), f["Lrb"].numpy()
        self.Rz, self.Rzb = f["Rz"].numpy(), f["Rzb"].numpy()
        lam = f["lam_z"].numpy()[None, :] + f["lam_r"].numpy()[:, None]
        self.inv_lam_band = np.ascontiguousarray(1.0 / lam[:rows])
        self.cols = nzs if self.full else max(64, self.points // nr)
        self.inv_lam_cols = None if self.full else np.ascontiguousarray(1.0 / lam[:, :self.cols])
        rng = np.random.default_rng(0)
        z = np.linspace(dx / 2, 1 - dx / 2, nz)
        r = np.linspace(dx / 2, nr * dx - dx / 2, nr)[:rows]
        self.Z, self.R = np.meshgrid(z, r)
        Z, R = self.Z, self.R
        blob = np.exp(-((Z - 0.5) ** 2 + R ** 2) / 0.02)
        self.w = rng.standard_normal((rows, nz)) * blob
        self.psi = np.zeros_like(Z)
        self.chi = np.zeros_like(Z)
        r_sph, z_cm = (0.075, 0.85) if periodic else (0.1, 0.25)
        ox.smooth_Heaviside(self.chi, -np.sqrt((Z - z_cm) ** 2 + R ** 2) + r_sph, dx * 2 ** 0.5)
        self.uz, self.ur, self.tmp, self.pv = (np.zeros_like(Z) for _ in range(4))
        self.rhs_cols = None if self.full else rng.standard_normal((nr, self.cols))
        self.nu, self.lam, self.t = 2e-3, 1e12, 0.0
        if name == "c3":
            self.eta1, self.eta2 = Z.copy(), R.copy()
            self.phi = 0.15 - np.sqrt((Z - 0.5) ** 2 + R ** 2)
            self.phi_orig = self.phi.copy()
            self.bchi, self.tchi = np.zeros_like(Z), np.zeros_like(Z)
            self.el = {k: np.zeros_like(Z) for k in ("s11", "s12", "s22", "e1z", "e1r", "e2z", "e2r", "tz", "tr")}
            self.avg_psi, self.avg_phi = np.zeros_like(Z), np.zeros_like(Z)
            self.lam = 1e8
            # the LS routine needs a square doubled array (the driver hard-codes 2 Nr == Nz): run it on a
            # 2n x 2n crop around the sphere's band; its cost is O(band), not O(grid)
            self.ls_n = min(rows, nz // 2, 512)
        if name == "c5":
            self.avg = [np.zeros_like(Z) for _ in range(3)]
            self.pchi = np.zeros_like(Z)
            nrd = 2 * rows
            self.Zd, self.Rd = np.meshgrid(z, np.linspace(dx / 2, nrd * dx - dx / 2, nrd))
            self.zp, self.rp, self.wp = self.Zd.copy(), self.Rd.copy(), 0 * self.Zd

    def describe(self):
        if self.full:
            what = f"one FULL un-extrapolated step at {self.nr}x{self.nz}"
        else:
            what = (f"band of {self.rows} rows x {self.nz} columns = {self.points} of {self.nr * self.nz} points "
                    f"(stencils + z GEMMs ({self.rows}x{self.Rz.shape[0]})({self.Rz.shape[0]}x{self.Rz.shape[0]}) on the "
                    f"band, r GEMMs ({self.nr}x{self.nr})({self.nr}x{self.cols}) on the matching column block; "
                    f"full step = x{self.nr * self.nz / self.points:.0f})")
        return what + ("; reference algorithm: 4 dense GEMMs (threaded BLAS), C/OpenMP ports of the numba / pystencils "
                       "kernels on all cores, NumPy glue")

    # -- the reference's solve on the sample ------------------------------------------------------------------
    def solve(self):
        g = self.g
        rhs = self.w[:, g:self.nz - g] if g else self.w
        if self.full:
            spec = np.linalg.multi_dot([self.Lr, rhs, self.Rz])
            np.multiply(spec, self.inv_lam_band, out=spec)
            out = np.linalg.multi_dot([self.Lrb, spec, self.Rzb])
            if g:
                self.psi[:, g:self.nz - g] = out
                self.ox.periodic_ghost_comm(self.psi, g)
            else:
                self.psi[...] = out
            return
        spec = rhs @ self.Rz                                   # z products on the band's rows
        np.multiply(spec, self.inv_lam_band, out=spec)
        out = spec @ self.Rzb
        if g:
            self.psi[:, g:self.nz - g] = 1e-6 * out
        else:
            self.psi[...] = 1e-6 * out                          # band result is not a solution: keep it small
        c = self.Lr @ self.rhs_cols                              # r products on the matching column block
        np.multiply(c, self.inv_lam_cols, out=c)
        self.rhs_cols[...] = 1e-3 * (self.Lrb @ c) + 1.0

    def _penalise(self, chi, Uz, dt):
        ox, ck = self.ox, self.ox.c_kernels
        uzu, uru = self.uz.copy(), self.ur.copy()
        ck.brinkmann_penalize(self.lam, dt, chi, Uz, 0.0, uzu, uru, self.uz, self.ur)
        ck.compute_vorticity_from_velocity(self.pv, self.uz - uzu, self.ur - uru, self.dx)
        self.w += self.pv

    def step(self):
        ox, ck, dx, g = self.ox, self.ox.c_kernels, self.dx, self.g
        Z, R, w = self.Z, self.R, self.w
        eps = np.finfo(float).eps
        t0 = time.perf_counter()
        ox.kill_boundary_vorticity_sine_z(w, Z, 3, dx)
        ox.kill_boundary_vorticity_sine_r(w, R, 3, dx)
        self.solve()
        ck.compute_velocity_from_psi(self.uz, self.ur, self.psi, R, dx)
        if self.name in ("c1", "c2", "c4"):
            self.uz += 1.0
            dt = min(0.9 * dx ** 2 / 4 / self.nu, 0.1 * dx / (np.amax(np.fabs(self.uz) + np.fabs(self.ur)) + eps))
            if g:
                ox.periodic_ghost_comm(self.ur, g)
                ox.periodic_ghost_comm(self.uz, g)
            self._penalise(self.chi, 0.0, dt)
            _cd = np.sum(R * self.chi * self.uz)
            ox.advect_vorticity_via_eno3(w, self.uz, self.ur, dt, dx, periodic_ghost=0)
            if g:
                ox.periodic_ghost_comm(w, g)
                self.tmp[...] = w                    # periodic RK2: ghost copy between the stages
                ck.diffusion_RK2(w, self.tmp, R, self.nu, dt, dx)
                ox.periodic_ghost_comm(w, g)
            else:
                ck.diffusion_RK2(w, self.tmp, R, self.nu, dt, dx)
        elif self.name == "c3":
            a = self.el
            dt = min(0.1 * dx / 3.0, 0.1 * dx / (np.amax(np.fabs(self.uz) + np.fabs(self.ur)) + eps),
                     0.9 * dx ** 2 / 4 / self.nu)
            self.avg_psi += self.psi * dt
            self.avg_phi += self.phi * dt
            ox.advect_refmap_via_eno3(self.eta1, self.eta2, self.uz, self.ur, dt, dx)
            band = np.where(self.phi > -3 * dx)
            self.phi_orig[...] = 0.15 - np.sqrt((self.eta1 - 0.5) ** 2 + self.eta2 ** 2)
            self.phi[band] = self.phi_orig[band]
            ox.advect_vorticity_via_eno3(w, self.uz, self.ur, dt, dx)
            ox.smooth_Heaviside(self.bchi, self.phi, 2 * dx)
            n = self.ls_n
            k0 = self.nz // 2 - n
            inside = (self.bchi > 0.5)[:n, k0:k0 + 2 * n]
            e1c = np.ascontiguousarray(self.eta1[:n, k0:k0 + 2 * n])
            e2c = np.ascontiguousarray(self.eta2[:n, k0:k0 + 2 * n])
            ox.extrapolate_eta_with_least_squares(inside, np.ascontiguousarray(self.phi[:n, k0:k0 + 2 * n]), e1c, e2c,
                                                  6 * dx, n, Z[0, k0:k0 + 2 * n])
            self.eta1[:n, k0:k0 + 2 * n], self.eta2[:n, k0:k0 + 2 * n] = e1c, e2c
            ox.solid_sigma(a["s11"], a["s12"], a["s22"], 2.0, dx, self.eta1, self.eta2, a["e1z"], a["e1r"], a["e2z"],
                           a["e2r"])
            for k in ("s11", "s12", "s22"):
                a[k][...] = self.bchi * a[k]
            ox.update_vorticity_from_solid_stress(w, a["tz"], a["tr"], a["s11"], a["s12"], a["s22"], R, dt, dx)
            ox.smooth_Heaviside(self.tchi, -np.sqrt((Z - 0.5) ** 2 + R ** 2) + 0.0375, 2 * dx)
            self._penalise(self.tchi, 0.7, dt)
            ck.diffusion_RK2(w, self.tmp, R, self.nu, dt, dx)
        else:   # c5
            dt = min(0.9 * dx ** 2 / 4 / self.nu, 0.1 / (np.amax(np.fabs(w)) + eps), 1e-3)
            d = np.sqrt((Z - 0.0) ** 2 + R ** 2)                   # analytic bubble potential flow (8 field expressions)
            fac = 0.25 ** 2 * 0.3 * np.sin(self.t) / (d ** 2 + eps) ** 1.5
            self.uz += (1 - self.chi) * fac * Z
            self.ur += (1 - self.chi) * fac * R
            for acc, f in zip(self.avg, (self.pchi, self.psi, w)):
                acc += f * dt
            ox.smooth_Heaviside(self.pchi, -np.sqrt((Z - 0.5) ** 2 + R ** 2) + 0.05, dx * 2 ** 0.5)
            self._penalise(self.pchi, 0.1, dt)
            _f = ox.compute_force_on_body(R, self.pchi, 1.0, self.lam, self.uz, 0.1, 1e-3, dt, 0.0)
            ox.advect_vorticity_via_particles(self.zp, self.rp, self.wp, w, self.Zd, self.Rd, self.rows, self.uz,
                                              self.ur, dx, dt)
            ck.diffusion_RK2(w, self.tmp, R, self.nu, dt, dx)
        self.t += dt
        np.clip(w, -1e3, 1e3, out=w)                 # a band is not a closed flow problem: keep the data finite
        return time.perf_counter() - t0


def cpu_baseline(name, nr, nz, budget_s=12.0, sample=None):
    """cpu_baseline object: median time of a few sample steps (bounded by ``budget_s``) -> updates/s"""
    s = sample if sample is not None else CpuStepSample(name, nr, nz)
    s.step()                                          # warm-up (page faults, BLAS thread start)
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < 2 or (time.perf_counter() < t_end and len(times) < 8):
        times.append(s.step())
    secs = float(np.median(times))
    return {"value": s.points / secs, "unit": UNIT, "cores": s.cores, "kind": "port",
            "sample": s.describe() + f"; {len(times)} sample steps, median {secs * 1e3:.0f} ms"
                      + ("" if s.full else f" -> {secs * nr * nz / s.points:.2f} s per extrapolated full step"),
            "measured_full_step": bool(s.full), "ms_per_sample_step": secs * 1e3}


def run_reference_arm(args, name, nr, nz):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    s = CpuStepSample(name, nr, nz)
    for _ in range(args.warmup):
        s.step()
    t0 = time.perf_counter()
    times = [s.step() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    secs = wall / args.steps                           # exactly K timed steps, wall clock around them
    value = s.points / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
            "scaling": "strong" if name in ("c4", "c2") else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(name, nr, nz, world, args.cases, args.reinit),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": s.cores, "kind": "port",
                             "sample": s.describe() + f"; each of the {args.steps} timed steps is one sample step "
                                       f"({s.points} points), median {float(np.median(times)) * 1e3:.0f} ms",
                             "measured_full_step": bool(s.full)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "points_per_step": s.points,
            "full_step_ms_estimate": secs * 1e3 * nr * nz / s.points,
            "cpu_threads": {"blas": os.environ.get("OPENBLAS_NUM_THREADS"), "openmp": os.environ.get("OMP_NUM_THREADS")}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def measure_fp64_peak(torch):
    """cuBLAS DGEMM 8192^3 (burst, best of 5): the FP64 tensor-pipe yardstick MEASURED_PEAKS.json lacks."""
    n = 8192
    a = torch.randn((n, n), dtype=torch.float64, device="cuda")
    b = torch.randn((n, n), dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


class _ConfigRunner:
    """BASELINE.json configs other than the rigid ones, behind the stepper interface bench uses:
    c3 soft-sphere streaming, c5 particle ensemble (cases per GPU)."""

    def __init__(self, name, nr, nz, basis, cases, reinit=False):
        import torch

        from pyaxisymflow_b200.timestep import ParticleFlowStepper, SoftSphereStepper

        self.torch = torch
        self.cases = 1
        self.ensemble = None
        if name == "c3":
            # Z_cm off the grid's mirror plane when re-initialising: exact ties make the marcher (and the GPU
            # iteration) order/rounding dependent and cost extra sweeps
            # without the re-initialisation the step runs device resident (loop scalars and the LS sweep loop on the
            # device) as one replayed CUDA graph; AXB_SOFT_HOST=1 keeps the host-driven loop
            dev = not reinit and not os.environ.get("AXB_SOFT_HOST")
            self.members = [SoftSphereStepper(nz, grid_size_r=nr, basis=basis, reinit_levelset=reinit,
                                              Z_cm=0.47 if reinit else 0.5, device_scalars=dev, use_graph=dev,
                                              overlap_ls=not os.environ.get("AXB_SOFT_NO_OVERLAP"))]
        else:
            from pyaxisymflow_b200.timestep import ParticleEnsemble

            freqs = [4.0, 8.0, 12.0, 16.0, 20.0, 24.0, 28.0, 32.0]
            params = [(8.0, 0.01)] + [(freqs[i % 8], 0.005 * (1 + (i // 8) % 8)) for i in range(1, cases)]
            self.cases = cases
            mode = os.environ.get("AXB_ENSEMBLE", "batched")
            if mode in ("batched", "members"):
                # SURVEY 8e "Ensemble": one (nr, cases nz) tensor per field, every operation ONE launch over all
                # members (axb_grid_t.batch), one solve for all members, loop scalars on the device, the whole ensemble
                # step replayed as one CUDA graph ("members": one launch per member and operation instead)
                self.ensemble = ParticleEnsemble.batched_ensemble(params, nz, nr, use_graph=True, basis=basis,
                                                                  launch=mode)
                self.members = self.ensemble.members
            else:
                first = ParticleFlowStepper(nz, grid_size_r=nr, basis=basis, freq=params[0][0], e=params[0][1])
                self.members = [first] + [ParticleFlowStepper(nz, grid_size_r=nr, freq=f, e=e, solver=first.solver)
                                          for f, e in params[1:]]
                self.ensemble = ParticleEnsemble(self.members) if mode == "streams" else None   # else "serial"
        self.name = name
        self.solver = self.members[0].solver
        self.vorticity = self.members[0].vorticity
        self.char_func = None
        self.periodic = False

    def step(self, n=1):
        if self.ensemble is not None:
            self.ensemble.step(n)
            return
        for _ in range(n):
            for m in self.members:
                m.step(1)

    def step_probed(self):
        self.step(1)
        return None

    @property
    def launches_replayed(self):
        """kernels launched through CUDA-graph replays (the library's launch counter only sees eager launches)"""
        objs = [self.ensemble] if getattr(self.ensemble, "batched", False) else self.members
        return sum(getattr(o, "launches_replayed", 0) for o in objs)

    def host_state(self):
        """device fields a host-resident caller has to supply for a step / gets back from it (e2e leg)"""
        if self.name == "c3":
            m = self.members[0]
            return [m.vorticity, m.eta1, m.eta2, m.ball_phi], [m.vorticity, m.eta1, m.eta2, m.ball_phi]
        if getattr(self.ensemble, "batched", False):
            w = self.ensemble._w_wide                        # all members' vorticity, one (nr, cases nz) tensor
            return [w], [w]
        return [m.vorticity for m in self.members], [m.vorticity for m in self.members]

    def solve_flops(self):
        return self.solver.flops()

    def solve_hbm_bytes(self):
        return self.members[0].solver.hbm_bytes()

    def solver_basis(self):
        return self.solver.basis

    def solve_kernel_note(self):
        return self.members[0].solver.kernel_note()


def solve_probe(stepper, torch, reps=3):
    """device time of one solve (of one member), measured outside the timed steps"""
    ens = getattr(stepper, "ensemble", None)
    if getattr(ens, "batched", False) and ens._solver is not None:
        # batched ensemble: ONE solve serves all members; report the per-member share of its time
        keep = ens._psi_wide.clone()
        ens._solver.solve(ens._psi_wide, ens._w_wide)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            ens._solver.solve(ens._psi_wide, ens._w_wide)
        b.record()
        torch.cuda.synchronize()
        ens._psi_wide.copy_(keep)
        return a.elapsed_time(b) / reps / len(ens.members)
    m = stepper.members[0] if hasattr(stepper, "members") else stepper
    per = getattr(m, "periodic", False)
    sol = m.psi[:, 2:-2] if per else m.psi
    rhs = m.vorticity[:, 2:-2] if per else m.vorticity
    keep = sol.clone()
    m.solver.solve(sol, rhs)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        m.solver.solve(sol, rhs)
    b.record()
    torch.cuda.synchronize()
    sol.copy_(keep)
    return a.elapsed_time(b) / reps


# algorithmic HBM bytes per grid point of one WHOLE step (DESIGN.md sections 3 and 6: every field a pass must read or
# write, counted once per pass; O(band) work and halos not counted):
#   c4 / c1 / c2 rigid flow: solve 80 + velocity 24 + penalisation 56 + ENO3 32 + one-pass RK2 16 = 208
#   c3 soft sphere (device-resident step): solve 80 + velocity 24 + 2 running averages 48 + reference-map ENO3 48 +
#      level-set pin 32 + vorticity ENO3 32 + Heaviside/mask 17 + masked map write-back of the LS extrapolation 33 +
#      fused elastic stress 40 + tether Heaviside 8 + penalisation 56 + two-stage RK2 40 = 458
#   c5 particle case: solve 80 + velocity 24 + max reduction 8 + bubble flow 40 + 3 running averages 72 + particle
#      Heaviside 8 + penalisation 56 + lattice remesh 32 + two-stage RK2 40 = 360
STEP_BYTES_PER_POINT = {"c4": 208, "c1": 208, "c2": 208, "c3": 458, "c5": 360}


def step_roofline(name, points, ms_per_step, mp):
    """whole-step HBM roofline: algorithmic bytes of all passes / measured copy bandwidth against the step time"""
    peak = mp.get("hbm_gbs") or 6650.0
    bpp = STEP_BYTES_PER_POINT[name]
    floor_ms = bpp * points / (peak * 1e9) * 1e3
    return {"bound": "hbm", "algorithmic_bytes_per_point": bpp, "points": points, "peak": peak, "unit": "GB/s",
            "achieved": bpp * points / (ms_per_step * 1e-3) / 1e9, "floor_ms": floor_ms, "frac": floor_ms / ms_per_step}


def roofline(stepper, mp, fp64_peak, tflops, hbm_bytes, traffic, solve_ms, share):
    """roofline object of the dominant operation, the fast-diagonalisation solve: tensor bound on the GEMM
    paths (FP64 DMMA, peak = cuBLAS DGEMM measured in this run), HBM bound on the transform + sweep paths (peak =
    MEASURED_PEAKS.json's copy bandwidth, else the profiling guide's fallback)."""
    if hbm_bytes is None:
        return {"bound": "tensor", "kernel": stepper.solve_kernel_note(), "achieved": tflops, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": tflops / fp64_peak, "traffic": traffic,
                "peak_source": "cuBLAS DGEMM 8192^3 burst measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "solve_ms": solve_ms, "solve_share_of_step": share, "hbm_peak_gbs": mp.get("hbm_gbs")}
    peak = mp.get("hbm_gbs")
    src = "MEASURED_PEAKS.json hbm_gbs (copy bandwidth measured on this pool)"
    if not peak:
        peak, src = 6650.0, "fallback of B200_PROFILING.md (6.65 TB/s)"
    gbs = hbm_bytes / (solve_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": stepper.solve_kernel_note(), "achieved": gbs, "peak": peak, "unit": "GB/s",
            "frac": gbs / peak, "traffic": traffic, "peak_source": src, "algorithmic_bytes": hbm_bytes,
            "solve_ms": solve_ms, "solve_share_of_step": share}


def make_stepper(name, nr, nz, args, world):
    from pyaxisymflow_b200.timestep import RigidFlowStepper

    if name == "c1":
        # the reference's own CPU-runnable case (FlowPastSphere CLI default 128x256): launch bound, so the
        # whole step is replayed as one CUDA graph.  --r-method / --z-method left on "auto" select the cosine-transform
        # + tridiagonal solve here as well (the library's own "auto" keeps the reference's four small GEMMs on LAPACK
        # bases below 1536 points a side: 0.13 ms of a 0.19 ms step on two CTAs)
        rm = "tridiagonal" if args.r_method == "auto" else args.r_method
        zm = "fft" if args.z_method == "auto" else args.z_method
        st = RigidFlowStepper(nz, grid_size_r=nr, use_graph=True, r_method=rm, z_method=zm)
        st.seed_vorticity()
        return st
    if name == "c2":
        if world > 1:                                # rows split over the ranks; the periodic wrap stays inside a row
            from pyaxisymflow_b200.rowslab import RowSlabRigidFlowStepper

            st = RowSlabRigidFlowStepper(nz, grid_size_r=nr, periodic=True, r_sph=0.075, Z_cm=0.85,
                                         use_graph=not args.no_graph)
        else:
            st = RigidFlowStepper(nz, grid_size_r=nr, periodic=True, r_sph=0.075, Z_cm=0.85,
                                  use_graph=not args.no_graph)
        st.seed_vorticity()
        return st
    if name == "c4":
        if world > 1:
            if os.environ.get("AXB_SLAB_Z"):         # the z-slab flow (two transposes per solve), for comparison
                from pyaxisymflow_b200.slab import SlabRigidFlowStepper

                st = SlabRigidFlowStepper(nz, grid_size_r=nr, r_method=args.r_method, z_method=args.z_method)
            else:                                    # rows split over the ranks: no transposes
                from pyaxisymflow_b200.rowslab import RowSlabRigidFlowStepper

                st = RowSlabRigidFlowStepper(nz, grid_size_r=nr, use_graph=not args.no_graph)
        else:
            st = RigidFlowStepper(nz, grid_size_r=nr, r_method=args.r_method, z_method=args.z_method,
                                  use_graph=not args.no_graph)
        # synthetic start: seeded band-limited vorticity blob so every kernel sees
        # non-trivial data from the first step on
        st.seed_vorticity()
        return st
    return _ConfigRunner(name, nr, nz, "auto", args.cases, reinit=args.reinit)


def measure_gpu(name, nr, nz, args, steps, warmup, world, rank, dist, sampler=None, with_cpu=True):
    """time `steps` steps of configuration `name`; returns the fields of its JSON object"""
    import torch

    from pyaxisymflow_b200 import _lib

    stepper = make_stepper(name, nr, nz, args, world)
    cases = getattr(stepper, "cases", 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if needs_flush(name, nr, nz) else None
    whole_graph = name == "c1"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # N > 1: the timed steps replay the whole step as ONE graph on every rank; the solve is probed on separate
    # steps right after the timed region (three graphs with events in between cost 0.1 ms of a 0.6 ms step)
    split_probe = world > 1 and getattr(stepper, "_use_graph", False)

    def one():
        if whole_graph or split_probe:
            stepper.step(1)
            return None
        return stepper.step_probed()

    for _ in range(warmup):
        one()
    barrier()
    if sampler is not None:
        sampler.start()
    # ---- timed region: exactly K steps, device-resident
    launches0 = _lib.launch_count() + getattr(stepper, "launches_replayed", 0)
    probes = []
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            probes.append(one())
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    else:
        pairs = []
        for _ in range(steps):
            flush.fill_(1)                                  # evict the fields from L2 (outside the event pair)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            probes.append(one())
            b.record()
            pairs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in pairs)
    launches = _lib.launch_count() + getattr(stepper, "launches_replayed", 0) - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    clocks = sampler.stop() if sampler is not None else None
    ms_per_step = ms / steps
    per_gpu_cases = cases
    slabbed = name in ("c4", "c2")                         # one domain split over the ranks (strong scaling)
    value = nr * nz / (ms_per_step * 1e-3) if slabbed else per_gpu_cases * world * nr * nz / (ms_per_step * 1e-3)
    warm_ms = None
    if flush is not None:                                   # the same steps back to back, L2-warm, for comparison
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        barrier()
        warm_ms = e0.elapsed_time(e1) / steps

    # ---- roofline of the dominant operation (the solve): algorithmic bytes or flops / its device time
    if probes and probes[0] is not None:
        solve_ms = float(np.mean([a.elapsed_time(b) for a, b in probes]))
    elif split_probe:
        for _ in range(2):
            stepper.step_probed()
        extra_probes = [stepper.step_probed() for _ in range(steps)]
        barrier()
        solve_ms = float(np.mean([a.elapsed_time(b) for a, b in extra_probes]))
    else:
        solve_ms = solve_probe(stepper, torch)
    flops = stepper.solve_flops()
    hbm_bytes = stepper.solve_hbm_bytes() if hasattr(stepper, "solve_hbm_bytes") else None
    out = {"value": value, "ms_per_step": ms_per_step, "gpu_launches": int(launches), "clocks": clocks,
           "solve_ms": solve_ms, "flops": flops, "hbm_bytes": hbm_bytes, "cases": cases, "stepper": stepper,
           "ms_per_step_l2_warm": warm_ms}
    return out


def e2e_rigid(stepper, nr, nz, steps, torch):
    """host-resident caller through RigidFlowStepper.step_host / HostStepPipeline: pinned host buffers, H2D of the
    step's inputs and D2H of its result inside the timed region"""
    from pyaxisymflow_b200.timestep import HostStepPipeline

    def pinned():
        return torch.empty((nr, nz), dtype=torch.float64).pin_memory()

    hw, ho = [pinned(), pinned()], [pinned(), pinned()]
    hc = pinned()
    for i in range(2):
        hw[i].copy_(stepper.vorticity)
    hc.copy_(stepper.char_func)
    stepper.step_host(hw[0], hc, ho[0])
    torch.cuda.synchronize()
    k = max(2, min(steps, 6))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(k):
        stepper.step_host(hw[i & 1], hc, ho[i & 1])
    b.record()
    torch.cuda.synchronize()
    serial_ms = a.elapsed_time(b) / k
    pipe = HostStepPipeline(stepper)

    def piped(char):
        for i in range(2):
            pipe.submit(hw[i & 1], char, ho[i & 1])
        pipe.drain()
        t0 = time.perf_counter()
        for i in range(k):
            pipe.submit(hw[i & 1], char, ho[i & 1])
        pipe.drain()
        return (time.perf_counter() - t0) * 1e3 / k       # three streams: wall clock around a full drain

    both_ms = piped(hc)          # every case brings its own characteristic function (a different body per case)
    e2e_ms = piped(None)         # the reference loop: fixed body, char_func set once, per-step input = the vorticity
    return {"value": nr * nz / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nr * nz * 8,
            "d2h_bytes_per_step": nr * nz * 8, "ms_per_step": e2e_ms,
            "mode": "HostStepPipeline: pinned host buffers, vorticity in / vorticity out every step (the characteristic "
                    "function of the fixed body is resident, set once before the loop like flow_past_sphere.py:88-96); "
                    "H2D of step k+1 and D2H of step k-1 overlap step k",
            "with_char_func_upload_ms_per_step": both_ms, "with_char_func_upload_value": nr * nz / (both_ms * 1e-3),
            "serial_ms_per_step": serial_ms, "serial_value": nr * nz / (serial_ms * 1e-3)}


def e2e_generic(runner, nr, nz, steps, torch):
    """c3 / c5: every step copies the evolving state in from pinned host memory and the result back (serial)"""
    ins, outs = runner.host_state()
    h_in = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in ins]
    h_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs]
    for h, t in zip(h_in, ins):
        h.copy_(t)
    torch.cuda.synchronize()

    def once():
        cur_in, _ = runner.host_state()                  # buffers are swapped inside the steppers
        for h, t in zip(h_in, cur_in):
            t.copy_(h, non_blocking=True)
        runner.step(1)
        _, cur_out = runner.host_state()
        for h, t in zip(h_out, cur_out):
            h.copy_(t, non_blocking=True)

    once()
    torch.cuda.synchronize()
    k = max(2, min(steps, 6))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        once()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / k
    cases = runner.cases
    return {"value": cases * nr * nz / (ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in ins)),
            "d2h_bytes_per_step": int(sum(t.numel() * t.element_size() for t in outs)), "ms_per_step": ms,
            "mode": "serial per step: state fields H2D from pinned host memory, one step, state fields D2H"}


def slab_vs_single(world, rank, dist, torch, nz=2048, steps=4, periodic=False):
    """correctness of the multi-GPU path, outside the timed region: the slab stepper against the single-GPU
    stepper on a reduced grid (nz/4 x nz), relative L-infinity of the vorticity after `steps` steps"""
    from pyaxisymflow_b200.timestep import RigidFlowStepper

    pk = {"periodic": True, "r_sph": 0.075, "Z_cm": 0.85} if periodic else {}
    if periodic:
        nz += 4                                      # inner width 2^11 + 2 x 2 ghost columns
    if os.environ.get("AXB_SLAB_Z") and not periodic:
        from pyaxisymflow_b200.slab import SlabRigidFlowStepper as Stepper
    else:
        from pyaxisymflow_b200.rowslab import RowSlabRigidFlowStepper as Stepper
    kw = {} if (os.environ.get("AXB_SLAB_Z") and not periodic) else {"use_graph": True}
    s = Stepper(nz, grid_size_r=(nz - 4 if periodic else nz) // 4, **kw, **pk)
    s.seed_vorticity()
    s.step(steps)
    w = s.gather_vorticity()
    err = None
    nr = (nz - 4 if periodic else nz) // 4
    if rank == 0:
        ref = RigidFlowStepper(nz, grid_size_r=nr, **pk)
        ref.seed_vorticity()
        ref.step(steps)
        torch.cuda.synchronize()
        err = ((w - ref.vorticity).abs().max() / ref.vorticity.abs().max()).item()
        del ref
    dist.barrier()
    if hasattr(s, "close"):
        s.close()
    del s
    torch.cuda.empty_cache()
    return err, f"{nr}x{nz}{' periodic' if periodic else ''}, {steps} steps"


def run_gpu_arm(args, name, nr, nz):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        mp = {}

    def config_line(cname, cnr, cnz, m, peak_holder, e2e, cpu):
        st = m["stepper"]
        share = m["cases"] * m["solve_ms"] / m["ms_per_step"]
        if m["hbm_bytes"] is None and peak_holder[0] is None:
            peak_holder[0] = measure_fp64_peak(torch)
        achieved = m["flops"] / (m["solve_ms"] * 1e-3) / 1e12
        traffic = None
        try:   # per-launch DRAM bytes of the dominant kernel from the committed ncu capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "solve_traffic.json")))
            if world == 1 and tj.get("grid") == [cnr, cnz] and m["hbm_bytes"] is not None:
                traffic = tj["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        cfg = config_dict(cname, cnr, cnz, world, args.cases, args.reinit)
        d = {"value": m["value"], "unit": UNIT, "ms_per_step": m["ms_per_step"],
             "config": cfg, "solver_basis": st.solver_basis(),
             "roofline": roofline(st, mp, peak_holder[0], achieved, m["hbm_bytes"], traffic, m["solve_ms"], share),
             "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": m["gpu_launches"], "clocks": m["clocks"]}
        if world == 1:
            d["step_roofline"] = step_roofline(cname, m["cases"] * cnr * cnz, m["ms_per_step"], mp)
        if m["ms_per_step_l2_warm"] is not None:
            d["ms_per_step_l2_warm"] = m["ms_per_step_l2_warm"]
        return d

    peak_holder = [None]
    sampler = ClockSampler(local) if rank == 0 else None
    m = measure_gpu(name, nr, nz, args, args.steps, args.warmup, world, rank, dist, sampler)
    stepper = m["stepper"]

    # ---- e2e: host-resident caller, H2D of the step's inputs + D2H of its result inside the timing
    e2e = None
    if world == 1:
        if name in ("c4", "c1", "c2"):
            e2e = e2e_rigid(stepper, nr, nz, args.steps, torch)
        else:
            e2e = e2e_generic(stepper, nr, nz, args.steps, torch)
    elif name in ("c4", "c2"):
        # every rank stages its own slab through its own PCIe link (serial per call: copy in, step, copy out)
        L = stepper.L
        own = tuple(L.owned(stepper.vorticity).shape)        # (nr, nz/P) columns or (nr/P, nz) rows
        hw = torch.empty(own, dtype=torch.float64).pin_memory()
        hc = torch.empty(own, dtype=torch.float64).pin_memory()
        ho = torch.empty(own, dtype=torch.float64).pin_memory()
        hw.copy_(L.owned(stepper.vorticity))
        hc.copy_(L.owned(stepper.char_func))
        stepper.step_host(hw, hc, ho)                        # sets the (fixed) body once
        torch.cuda.synchronize()
        dist.barrier()
        k = max(2, min(args.steps, 6))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            stepper.step_host(hw, None, ho)                  # per-step input = the vorticity
        b.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([a.elapsed_time(b) / k], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
        e2e = {"value": nr * nz / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nr * nz * 8,
               "d2h_bytes_per_step": nr * nz * 8, "ms_per_step": e2e_ms,
               "mode": f"step_host on every rank: each of the {world} ranks stages its own slab of the vorticity (pinned "
                       "host buffers) over its own PCIe link, serial per call; the fixed body's characteristic "
                       "function is resident"}
        del hw, hc, ho

    extra = {}
    if world > 1 and name in ("c4", "c2"):
        phases = stepper.phase_times() if hasattr(stepper, "phase_times") else None
        err, what = slab_vs_single(world, rank, dist, torch, periodic=(name == "c2"))
        extra = {"slab_vs_single_rel_linf": err, "slab_vs_single_case": what, "phases_ms": phases}

    if hasattr(stepper, "close"):
        torch.cuda.synchronize()
        stepper.close()                      # captured graphs go before the process group does
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(name, nr, nz)
        line = {"metric": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "higher_is_better": True, "scaling": "strong" if name in ("c4", "c2") else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic"}
        line.update(config_line(name, nr, nz, m, peak_holder, e2e, cpu))
        line.update(extra)
        del m, stepper
        torch.cuda.empty_cache()
        # ---- the other BASELINE.json configurations, same measurement each (1 GPU, default run only)
        if world == 1 and name == "c4" and not args.no_configs:
            line["configs"] = {}
            for cname in ("c1", "c2", "c3", "c5"):
                cnr, cnz = GRIDS[cname]
                try:
                    cs = ClockSampler(local)
                    cm = measure_gpu(cname, cnr, cnz, args, min(args.steps, 10), 3, 1, 0, dist, cs)
                    if cname in ("c1", "c2"):
                        ce = e2e_rigid(cm["stepper"], cnr, cnz, args.steps, torch)
                    else:
                        ce = e2e_generic(cm["stepper"], cnr, cnz, args.steps, torch)
                    ccpu = None if args.no_cpu else cpu_baseline(cname, cnr, cnz, budget_s=8.0)
                    line["configs"][cname] = config_line(cname, cnr, cnz, cm, peak_holder, ce, ccpu)
                    del cm
                except Exception as exc:  # noqa: BLE001 -- a failing side configuration must not void the headline
                    line["configs"][cname] = {"error": f"{type(exc).__name__}: {exc}"}
                torch.cuda.empty_cache()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nz", type=int, default=None, help="grid_size_z (default: the configuration's own)")
    ap.add_argument("--nr", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-configs", action="store_true", help="c4 only: skip the c1/c2/c3/c5 objects under `configs`")
    ap.add_argument("--config", default="c4", choices=["c4", "c1", "c2", "c3", "c5"],
                    help="c4 (default): 4096x16384 rigid flow, the configuration the metric is quoted on; "
                         "c1: 128x256 rigid flow; c2: periodic 1024x4096; c3: soft sphere 2048x8192; "
                         "c5: particle ensemble 1024x2048")
    ap.add_argument("--cases", type=int, default=8, help="ensemble members per GPU (c5)")
    ap.add_argument("--reinit", action="store_true",
                    help="c3: include the narrow-band level-set re-initialisation (csrc/reinit.cu) in the step")
    ap.add_argument("--no-graph", action="store_true",
                    help="c4 / c2, 1 GPU: launch the step kernel by kernel instead of replaying CUDA graphs (before / "
                         "solve / after, with the roofline events between them)")
    ap.add_argument("--r-method", default="auto", choices=["auto", "eigen", "tridiagonal"],
                    help="r direction of the solve: eigen-decomposition GEMMs (the reference's algorithm) or a "
                         "batched tridiagonal solve per z-mode (auto: tridiagonal on grids too large for la.eig)")
    ap.add_argument("--z-method", default="auto", choices=["auto", "gemm", "fft"],
                    help="z transforms of the solve: GEMMs with the eigenvector matrix (parity-split) or "
                         "shared-memory FFTs (auto: fft where available with the tridiagonal r solve)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    nr, nz = GRIDS[args.config]
    if args.nz is not None:
        nr, nz = (args.nz // 2 if args.config in ("c5", "c1") else args.nz // 4), args.nz
    if args.nr is not None:
        nr = args.nr
    if args.impl == "reference":
        run_reference_arm(args, args.config, nr, nz)
    else:
        run_gpu_arm(args, args.config, nr, nz)


if __name__ == "__main__":
    main()
