#!/usr/bin/env python
"""bench.py -- grid-point updates / s per full timestep (FP64), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            (our arm, one process per GPU)
    python bench.py --impl reference --steps K --warmup W    (reference arm: the CPU path)

A "step" is one complete loop iteration of the rigid-flow driver
(examples/FlowPastSphere/flow_past_sphere.py:107-207: boundary damping, streamfunction solve,
velocity, CFL reduction, penalisation + drag, ENO3 advection, RK2 diffusion) on the synthetic
4096 x 16384 FP64 grid the metric is quoted on (BASELINE.json configs[3]; it fits one GPU).
Fields (512 MiB each) are far larger than the 126 MB L2, so no explicit flush is needed
between timed iterations.  Timing: CUDA events on the launching stream, max over ranks.
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "grid-pt updates/s per full timestep (FP64)"
UNIT = "grid-pt updates/s"


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
        # "under load" = the upper half of the samples
        sm_sorted = sorted(sm)
        return {"sm_mhz": float(np.median(sm_sorted[len(sm_sorted) // 2:])) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path (oracle port) on a bounded sample of the same workload
# ----------------------------------------------------------------------------------------------
def cpu_step_seconds(nr, nz, repeats=1):
    """One full rigid-flow timestep on the host, extrapolated from a bounded sample.

    The reference's path at this grid is: four OpenBLAS dgemms (numpy.linalg.multi_dot) +
    numba/NumPy stencil passes + the pystencils ENO3 kernels.  pystencils cannot be installed,
    and la.eig at Nz=16384 takes tens of minutes, so:
      * stencils: the oracle's NumPy restatement (single thread, like numba) and its C/OpenMP
        ENO3 (all cores, like pystencils with num_threads=nproc) on a z-window of the grid,
        scaled by nz / window;
      * solve: numpy matmul (threaded BLAS) of each of the four products on a column / row
        block, scaled by the block count (BLAS cost is linear in the blocked dimension).
    Returns (seconds per full step, description of the sample).
    """
    from oracle import axisym_oracle as ox

    dx = 1.0 / nz
    nzw = min(nz, 1024)
    rng = np.random.default_rng(0)
    z = np.linspace(dx / 2, nzw * dx - dx / 2, nzw)
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    Z, R = np.meshgrid(z, r)
    blob = np.exp(-((Z - 0.5 * nzw * dx) ** 2 + R ** 2) / 0.02)
    w = rng.standard_normal((nr, nzw)) * blob
    psi = 1e-3 * rng.standard_normal((nr, nzw)) * blob
    chi = np.zeros_like(Z)
    ox.smooth_Heaviside(chi, -np.sqrt((Z - 0.25 * nzw * dx) ** 2 + R ** 2) + 0.1 * nzw * dx, dx * 2 ** 0.5)
    uz, ur, tmp, pv = (np.zeros_like(Z) for _ in range(4))
    eps = np.finfo(float).eps
    nu, lam = 2e-3, 1e12
    t_st = []
    for _ in range(repeats + 1):
        t0 = time.perf_counter()
        ox.kill_boundary_vorticity_sine_z(w, Z, 3, dx)
        ox.kill_boundary_vorticity_sine_r(w, R, 3, dx)
        ox.compute_velocity_from_psi(uz, ur, psi, R, dx)
        uz += 1.0
        dt = min(0.9 * dx ** 2 / 4 / nu, 0.1 * dx / (np.amax(np.fabs(uz) + np.fabs(ur)) + eps))
        uzu, uru = uz.copy(), ur.copy()
        ox.brinkmann_penalize(lam, dt, chi, 0.0, 0.0, uzu, uru, uz, ur)
        ox.compute_vorticity_from_velocity(pv, uz - uzu, ur - uru, dx)
        w += pv
        _cd = np.sum(R * chi * uz)
        ox.advect_vorticity_via_eno3(w, uz, ur, dt, dx)
        ox.diffusion_RK2(w, tmp, R, nu, dt, dx)
        t_st.append(time.perf_counter() - t0)
    stencil = min(t_st[1:]) * (nz / nzw)

    def gemm_time(m, k, n):
        a, b = np.ones((m, k)), np.ones((k, n))
        a @ b
        best = 1e30
        for _ in range(max(1, repeats)):
            t0 = time.perf_counter()
            a @ b
            best = min(best, time.perf_counter() - t0)
        return best

    nzc = min(nz, 512)      # column block for the r-transforms  (nr x nr)(nr x nzc)
    nrc = min(nr, 128)      # row block for the z-transforms     (nrc x nz)(nz x nz)
    t_r = gemm_time(nr, nr, nzc) * (nz / nzc)
    t_z = gemm_time(nrc, nz, nz) * (nr / nrc)
    solve = 2 * t_r + 2 * t_z
    sample = (f"stencils on a {nr}x{nzw} z-window x{nz // nzw}; dgemm blocks ({nr}x{nr})({nr}x{nzc}) x{nz // nzc} "
              f"and ({nrc}x{nz})({nz}x{nz}) x{nr // nrc}, each twice per step; "
              f"stencil {stencil:.2f} s + solve {solve:.2f} s per extrapolated step")
    return stencil + solve, sample


def run_reference_arm(args, nr, nz):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    secs, sample = None, ""
    for _ in range(max(0, args.warmup) and 1):
        cpu_step_seconds(nr, nz)
    times = []
    for _ in range(max(1, min(args.steps, 3))):
        s, sample = cpu_step_seconds(nr, nz)
        times.append(s)
    secs = float(np.median(times))
    value = nr * nz / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"rigid-flow timestep (FlowPastSphere loop body) at {nr}x{nz}", "grid": [nr, nz]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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


def run_gpu_arm(args, nr, nz):
    import torch
    import torch.distributed as dist

    from pyaxisymflow_b200 import _lib
    from pyaxisymflow_b200.timestep import RigidFlowStepper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    basis = "auto"
    cases = 1
    workload = f"rigid-flow timestep (FlowPastSphere loop body) at {nr}x{nz}"
    if args.config == "c1":
        # the reference's own CPU-runnable case (FlowPastSphere CLI default 128x256): launch bound, so the
        # whole step is replayed as one CUDA graph
        stepper = RigidFlowStepper(nz, grid_size_r=nr, basis=basis, use_graph=True)
        stepper.seed_vorticity()
        scaling = "weak"
        workload = f"rigid-flow timestep (FlowPastSphere CLI default) at {nr}x{nz}, one CUDA graph per step"
    elif args.config == "c4":
        if world > 1:
            from pyaxisymflow_b200.slab import SlabRigidFlowStepper

            stepper = SlabRigidFlowStepper(nz, grid_size_r=nr, r_method=args.r_method, z_method=args.z_method)
        else:
            stepper = RigidFlowStepper(nz, grid_size_r=nr, basis=basis, r_method=args.r_method,
                                       z_method=args.z_method, use_graph=not args.no_graph)
        scaling = "strong"
        # synthetic start: seeded band-limited vorticity blob so every kernel sees
        # non-trivial data from the first step on
        stepper.seed_vorticity()
    else:
        stepper = _ConfigRunner(args.config, nr, nz, basis, args.cases, reinit=args.reinit)
        scaling = "weak"          # independent cases / replicas per GPU, no communication
        cases = stepper.cases * world
        workload = stepper.workload

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if args.config == "c1":
        stepper.step(args.warmup)
    else:                                   # the same call as the timed loop (graph capture happens here)
        for _ in range(args.warmup):
            stepper.step_probed()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region: exactly K steps, device-resident
    launches0 = _lib.launch_count() + getattr(stepper, "launches_replayed", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probes = []
    e0.record()
    for _ in range(args.steps):
        if args.config == "c1":
            stepper.step(1)
            probes.append(None)
        else:
            probes.append(stepper.step_probed())
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() + getattr(stepper, "launches_replayed", 0) - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = cases * nr * nz / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (k_dgemm): flops of the four GEMMs / their device time
    if probes and probes[0] is not None:
        solve_ms = float(np.mean([a.elapsed_time(b) for a, b in probes]))
    elif hasattr(stepper, "solve_probe"):
        solve_ms = stepper.solve_probe()
    else:
        sol = stepper.psi.clone()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stepper.solver.solve(sol, stepper.vorticity)
        a.record()
        for _ in range(5):
            stepper.solver.solve(sol, stepper.vorticity)
        b.record()
        torch.cuda.synchronize()
        solve_ms = a.elapsed_time(b) / 5
    solves_per_step = getattr(stepper, "cases", 1)
    flops = stepper.solve_flops()
    achieved = flops / (solve_ms * 1e-3) / 1e12
    # DCT + sweep solve: bandwidth bound, so the roofline is algorithmic HBM bytes / device time
    hbm_bytes = stepper.solve_hbm_bytes() if hasattr(stepper, "solve_hbm_bytes") else None

    # ---- e2e: host-resident caller, H2D of the step's inputs + D2H of its result inside the timing.
    # Cases are independent host buffer sets streamed through HostStepPipeline (copies of neighbouring
    # cases overlap the step); the strictly serial step_host() time is reported beside it.
    e2e = None
    if world == 1 and args.config in ("c4", "c1"):
        from pyaxisymflow_b200.timestep import HostStepPipeline

        def pinned():
            return torch.empty((nr, nz), dtype=torch.float64).pin_memory()

        hw, hc, ho = [pinned(), pinned()], [pinned(), pinned()], [pinned(), pinned()]
        for i in range(2):
            hw[i].copy_(stepper.vorticity)
            hc[i].copy_(stepper.char_func)
        stepper.step_host(hw[0], hc[0], ho[0])
        torch.cuda.synchronize()
        k = max(2, min(args.steps, 6))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(k):
            stepper.step_host(hw[i & 1], hc[i & 1], ho[i & 1])
        b.record()
        torch.cuda.synchronize()
        serial_ms = a.elapsed_time(b) / k
        pipe = HostStepPipeline(stepper)
        for i in range(2):
            pipe.submit(hw[i & 1], hc[i & 1], ho[i & 1])
        pipe.drain()
        t0 = time.perf_counter()
        for i in range(k):
            pipe.submit(hw[i & 1], hc[i & 1], ho[i & 1])
        pipe.drain()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / k       # three streams: wall clock around a full drain
        e2e = {"value": nr * nz / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * nr * nz * 8,
               "d2h_bytes_per_step": nr * nz * 8, "ms_per_step": e2e_ms,
               "mode": "HostStepPipeline: pinned host buffers, H2D of case k+1 and D2H of case k-1 overlap the step "
                       "of case k",
               "serial_ms_per_step": serial_ms, "serial_value": nr * nz / (serial_ms * 1e-3)}
        del hw, hc, ho, pipe

    if world > 1 and args.config == "c4":
        # every rank stages its own z-slab through its own PCIe link (serial per call: copy in, step, copy out)
        L = stepper.L
        hw = torch.empty((nr, L.nzl), dtype=torch.float64).pin_memory()
        hc = torch.empty((nr, L.nzl), dtype=torch.float64).pin_memory()
        ho = torch.empty((nr, L.nzl), dtype=torch.float64).pin_memory()
        hw.copy_(L.owned(stepper.vorticity))
        hc.copy_(L.owned(stepper.char_func))
        stepper.step_host(hw, hc, ho)
        barrier()
        k = max(2, min(args.steps, 6))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            stepper.step_host(hw, hc, ho)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / k], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
        e2e = {"value": nr * nz / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * nr * nz * 8,
               "d2h_bytes_per_step": nr * nz * 8, "ms_per_step": e2e_ms,
               "mode": f"step_host on every rank: each of the {world} ranks stages its own z-slab (pinned host "
                       "buffers) over its own PCIe link, serial per call"}
        del hw, hc, ho

    if rank == 0:
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            mp = {}
        peak = measure_fp64_peak(torch) if hbm_bytes is None else None
        cpu = None
        if world == 1 and not args.no_cpu and args.config in ("c4", "c1"):
            secs, sample = cpu_step_seconds(nr, nz)
            cpu = {"value": nr * nz / secs, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": sample}
        traffic = None
        try:   # per-launch DRAM bytes of the dominant kernel from the committed ncu capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles",
                                             "gemm_traffic.json" if hbm_bytes is None else "solve_traffic.json")))
            if world == 1 and tj.get("grid") == [nr, nz]:
                traffic = tj["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "name": args.config, "cases": cases,
                       "grid": [nr, nz], "parallelism": "single GPU" if world == 1 else (
                           f"z-slab x{world}" if args.config == "c4" else f"{world} independent replicas"),
                       "l2": "fields (%.0f MiB each) exceed the 126 MB L2; no flush needed" % (nr * nz * 8 / 2 ** 20),
                       "basis": stepper.solver_basis()},
            "roofline": roofline(stepper, mp, peak, achieved, hbm_bytes, traffic, solve_ms,
                                 solves_per_step * solve_ms / ms_per_step),
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _ConfigRunner:
    """BASELINE.json configs other than the headline one, behind the stepper interface bench uses:
    c2 periodic flow past a sphere, c3 soft-sphere streaming, c5 particle ensemble (cases per GPU)."""

    def __init__(self, name, nr, nz, basis, cases, reinit=False):
        import torch

        from pyaxisymflow_b200.timestep import ParticleFlowStepper, RigidFlowStepper, SoftSphereStepper

        self.torch = torch
        self.cases = 1
        if name == "c2":
            self.members = [RigidFlowStepper(nz, grid_size_r=nr, periodic=True, r_sph=0.075, Z_cm=0.85, basis=basis)]
            self.members[0].seed_vorticity()
            self.workload = f"PeriodicFlowPastSphere loop body at {nr}x{nz} (periodic z, ghost 2)"
        elif name == "c3":
            # Z_cm off the grid's mirror plane when re-initialising: exact ties make the marcher (and the GPU
            # iteration) order/rounding dependent and cost extra sweeps
            self.members = [SoftSphereStepper(nz, grid_size_r=nr, basis=basis, reinit_levelset=reinit,
                                              Z_cm=0.47 if reinit else 0.5)]
            self.workload = (f"SoftSphereStreaming loop body at {nr}x{nz} (reference-map solid; "
                             + ("narrow-band level-set re-initialisation on the GPU included)" if reinit else
                                "skfmm re-initialisation excluded on both arms)"))
        else:
            first = ParticleFlowStepper(nz, grid_size_r=nr, basis=basis)
            self.members = [first]
            freqs = [4.0, 8.0, 12.0, 16.0, 20.0, 24.0, 28.0, 32.0]
            for i in range(1, cases):
                self.members.append(ParticleFlowStepper(nz, grid_size_r=nr, freq=freqs[i % 8],
                                                        e=0.005 * (1 + (i // 8) % 8), solver=first.solver))
            self.cases = cases
            self.workload = f"ensemble of {cases} ParticleOscillatoryFlowCases per GPU at {nr}x{nz} (MP4 remeshing)"
        self.ensemble = None
        if name == "c5" and not os.environ.get("AXB_ENSEMBLE_SERIAL"):
            from pyaxisymflow_b200.timestep import ParticleEnsemble

            self.ensemble = ParticleEnsemble(self.members)
            self.workload += "; members interleaved on their own streams"
        self.solver = self.members[0].solver
        self.vorticity = self.members[0].vorticity
        self.char_func = None

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

    def solve_probe(self, reps=3):
        """device time of one solve of one member, measured outside the timed steps"""
        t = self.torch
        m = self.members[0]
        per = getattr(m, "periodic", False)
        sol = m.psi[:, 2:-2] if per else m.psi
        rhs = m.vorticity[:, 2:-2] if per else m.vorticity
        keep = sol.clone()
        m.solver.solve(sol, rhs)
        a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            m.solver.solve(sol, rhs)
        b.record()
        t.cuda.synchronize()
        sol.copy_(keep)
        return a.elapsed_time(b) / reps

    def solve_flops(self):
        return self.solver.flops()

    def solve_hbm_bytes(self):
        return self.members[0].solver.hbm_bytes()

    def solver_basis(self):
        return self.solver.basis

    def solve_kernel_note(self):
        return self.members[0].solver.kernel_note()


def roofline(stepper, mp, fp64_peak, tflops, hbm_bytes, traffic, solve_ms, share):
    """roofline object of the dominant operation, the fast-diagonalisation solve: tensor bound on the GEMM
    paths (FP64 DMMA, peak = cuBLAS DGEMM measured in this run), HBM bound on the DCT + sweep path (peak =
    MEASURED_PEAKS.json's copy bandwidth, else the profiling guide's fallback)."""
    if hbm_bytes is None:
        return {"bound": "tensor", "kernel": stepper.solve_kernel_note(), "achieved": tflops, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": tflops / fp64_peak, "traffic": traffic,
                "peak_source": "cuBLAS DGEMM 8192^3 burst measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "solve_ms": solve_ms, "solve_share_of_step": share, "hbm_peak_gbs": mp.get("hbm_gbs")}
    peak = mp.get("hbm_gbs")
    src = "MEASURED_PEAKS.json hbm_gbs (copy bandwidth measured on this pool)"
    if not peak:
        peak, src = 6650.0, "of fallback (B200_PROFILING.md: 6.65 TB/s)"
    gbs = hbm_bytes / (solve_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": stepper.solve_kernel_note(), "achieved": gbs, "peak": peak, "unit": "GB/s",
            "frac": gbs / peak, "traffic": traffic, "peak_source": src, "launches": 4,
            "algorithmic_bytes": hbm_bytes, "solve_ms": solve_ms, "solve_share_of_step": share}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nz", type=int, default=None, help="grid_size_z (default: the configuration's own)")
    ap.add_argument("--nr", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--config", default="c4", choices=["c4", "c1", "c2", "c3", "c5"],
                    help="c4 (default): 4096x16384 rigid flow, the configuration the metric is quoted on; "
                         "c2: periodic 1024x4096; c3: soft sphere 2048x8192; c5: particle ensemble 1024x2048")
    ap.add_argument("--cases", type=int, default=8, help="ensemble members per GPU (c5)")
    ap.add_argument("--reinit", action="store_true",
                    help="c3: include the narrow-band level-set re-initialisation (csrc/reinit.cu) in the step")
    ap.add_argument("--no-graph", action="store_true",
                    help="c4, 1 GPU: launch the step kernel by kernel instead of replaying CUDA graphs (before / "
                         "solve / after, with the roofline events between them)")
    ap.add_argument("--r-method", default="auto", choices=["auto", "eigen", "tridiagonal"],
                    help="r direction of the solve: eigen-decomposition GEMMs (the reference's algorithm) or a "
                         "batched tridiagonal solve per z-mode (auto: tridiagonal on grids too large for la.eig)")
    ap.add_argument("--z-method", default="auto", choices=["auto", "gemm", "fft"],
                    help="z transforms of the solve: GEMMs with the eigenvector matrix (parity-split) or "
                         "shared-memory FFT cosine transforms (auto: fft where available with the tridiagonal r solve)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    defaults = {"c4": (4096, 16384), "c1": (128, 256), "c2": (1024, 4096), "c3": (2048, 8192), "c5": (1024, 2048)}
    nz = args.nz if args.nz is not None else defaults[args.config][1]
    nr = args.nr if args.nr is not None else (nz // 4 if args.config not in ("c5", "c1") else nz // 2)
    if args.impl == "reference":
        run_reference_arm(args, nr, nz)
    else:
        run_gpu_arm(args, nr, nz)


if __name__ == "__main__":
    main()
