"""Device-time the SURVEY 8f kernels at the BASELINE grid sizes (CUDA events, inputs resident in HBM, fields larger
than L2) and print one JSON line per kernel: algorithmic bytes / time against MEASURED_PEAKS.json's HBM figure."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyaxisymflow_b200 import _lib  # noqa: E402
from pyaxisymflow_b200.device import make_grid, ptr, stream_ptr  # noqa: E402
from pyaxisymflow_b200.reinit import NarrowBandReinit  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    peak = 6456.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    nr, nz = 4096, 16384
    dx = 1.0 / nz
    g = make_grid(nr, nz, nz, dx)
    f = [torch.rand(nr, nz, dtype=torch.float64, device="cuda") + 0.5 for _ in range(9)]
    r1d = torch.linspace(dx / 2, nr * dx - dx / 2, nr, dtype=torch.float64, device="cuda")
    s = stream_ptr()
    rows = []

    def line(name, bpp, ms, n=nr * nz, extra=None):
        gbs = bpp * n / ms / 1e6
        d = {"kernel": name, "grid": [nr, nz], "ms": round(ms, 4), "B_per_pt": bpp, "GBps": round(gbs, 1),
             "frac_of_hbm_peak": round(gbs / peak, 3), "peak": peak}
        d.update(extra or {})
        rows.append(d)
        print(json.dumps(d), flush=True)

    line("k_velocity_phi", 24, timed(lambda: _lib.call(
        "axb_velocity_from_phi", ctypes.byref(g), ptr(f[0]), ptr(f[1]), ptr(f[2]), s)))
    for mode, bpp in ((0, 56), (1, 72), (2, 72)):
        line(f"k_baroclinic<{mode}>", bpp, timed(lambda: _lib.call(
            "axb_baroclinic_vorticity_update", ctypes.byref(g), ptr(f[0]), ptr(f[1]), ptr(f[2]), ptr(f[3]), ptr(f[4]),
            ptr(f[5]), ptr(f[6]), ptr(f[7]), ptr(r1d), 1e-3, 1e-3, mode, s)))
    del f
    torch.cuda.empty_cache()
    # re-initialisation at the config-C3 size: deformed sphere, band 6 dx (soft_sphere_streaming.py:48)
    nr, nz = 2048, 8192
    dx = 1.0 / nz
    z = torch.linspace(dx / 2, 1 - dx / 2, nz, dtype=torch.float64, device="cuda")
    r = torch.linspace(dx / 2, nr * dx - dx / 2, nr, dtype=torch.float64, device="cuda")
    true = 0.15 - torch.sqrt((z[None, :] - 0.47) ** 2 + r[:, None] ** 2)
    phi0 = true * (1 + 0.2 * torch.sin(9 * z[None, :] + 5 * r[:, None]))
    rn = NarrowBandReinit(nr, nz)
    phi = phi0.clone()

    def go():
        phi.copy_(phi0)
        rn(phi, dx, 6 * dx)

    t_all = timed(go, reps=5, warm=2)
    t_copy = timed(lambda: phi.copy_(phi0), reps=5, warm=2)
    line("reinit (front + sweeps + ring, host-checked)", 8, t_all - t_copy, n=nr * nz,
         extra={"grid": [nr, nz], "sweeps": rn.sweeps, "note": "8 B/pt = the one unavoidable read of phi; the "
                "band itself is O(band) work"})
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_widen.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
