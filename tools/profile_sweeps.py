"""tools/profile_sweeps.py nr nz [solves] : a few factored tridiagonal solves for ncu (see tools/gpu_ncu_sweeps.sh)"""
import sys

import torch

sys.path.insert(0, ".")
from pyaxisymflow_b200 import _lib, fd  # noqa: E402
from pyaxisymflow_b200.device import ptr, stream_ptr  # noqa: E402

nr, nz = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dx = 1.0 / nz
sub, diag, sup, r = fd.radial_tridiagonal("stokes", "homogenous_neumann_along_z_and_r", nr, dx)
lam = fd.axial_natural_eigenvalues("neumann", 1.0, nz, dx)
lam[0] = lam[1]
dev = [torch.from_numpy(a).cuda() for a in (sub, diag, sup, lam, r)]
inv = torch.empty((nr, nz), dtype=torch.float64, device="cuda")
rc = torch.empty((nr, 4), dtype=torch.float64, device="cuda")
_lib.call("axb_tridiag_factor_columns", nr, nz, ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), ptr(dev[3]), ptr(dev[4]), 0.0, 1.0,
          ptr(inv), ptr(rc), stream_ptr())
x = torch.randn((nr, nz), dtype=torch.float64, device="cuda")
for _ in range(n):
    _lib.call("axb_tridiag_solve_factored", nr, nz, ptr(x), nz, ptr(inv), ptr(rc), stream_ptr())
torch.cuda.synchronize()
