"""tools/bench_sweeps.py : device time of axb_tridiag_solve_factored (forward + backward sweep, 48 algorithmic
B/pt) at the row x column counts of the BASELINE configs, single-warp TMA kernel against the warp-specialised one
(CUDA events, 20 solves each, right-hand sides larger than L2 or an L2 flush in between)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyaxisymflow_b200 import _lib, fd  # noqa: E402
from pyaxisymflow_b200.device import ptr, stream_ptr  # noqa: E402

shapes = [(1024, 4096, "c2"), (2048, 8192, "c3"), (1024, 16384, "c5 x8"), (4096, 16384, "c4"), (512, 16384, "c4 / 8 ranks"),
          (64, 256, "c1")]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda")
for nr, nz, tag in shapes:
    dx = 1.0 / nz
    sub, diag, sup, r = fd.radial_tridiagonal("stokes", "homogenous_neumann_along_z_and_r", nr, dx)
    lam = fd.axial_natural_eigenvalues("neumann", 1.0, nz, dx)
    lam[0] = lam[1]
    dev = [torch.from_numpy(a).cuda() for a in (sub, diag, sup, lam, r)]
    inv = torch.empty((nr, nz), dtype=torch.float64, device="cuda")
    rc = torch.empty((nr, 4), dtype=torch.float64, device="cuda")
    _lib.call("axb_tridiag_factor_columns", nr, nz, ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), ptr(dev[3]), ptr(dev[4]), 0.0,
              1.0, ptr(inv), ptr(rc), stream_ptr())
    x0 = torch.randn((nr, nz), dtype=torch.float64, device="cuda")
    res = {}
    for one_warp in (1, 0):
        _lib.call("axb_set_tridiag_sweep", one_warp)
        x = x0.clone()
        for _ in range(3):
            _lib.call("axb_tridiag_solve_factored", nr, nz, ptr(x), nz, ptr(inv), ptr(rc), stream_ptr())
        tot = 0.0
        n = 20
        for _ in range(n):
            x.copy_(x0)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("axb_tridiag_solve_factored", nr, nz, ptr(x), nz, ptr(inv), ptr(rc), stream_ptr())
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[one_warp] = (tot / n, x.clone())
    _lib.call("axb_set_tridiag_sweep", 0)
    same = bool(torch.equal(res[0][1], res[1][1]))
    gb = 48.0 * nr * nz / 1e9
    print(f"{tag:14s} {nr}x{nz}: single-warp {res[1][0]:.4f} ms ({gb / res[1][0] * 1e3 / 6543.1:.2f} of HBM)   "
          f"warp-specialised {res[0][0]:.4f} ms ({gb / res[0][0] * 1e3 / 6543.1:.2f})   bit-identical {same}")
