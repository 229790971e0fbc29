"""tiny driver for ncu: one narrow-band re-initialisation at the config-C3 size (2048 x 8192, band 6 dx)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyaxisymflow_b200.reinit import NarrowBandReinit  # noqa: E402

nr, nz = 2048, 8192
dx = 1.0 / nz
z = torch.linspace(dx / 2, 1 - dx / 2, nz, dtype=torch.float64, device="cuda")
r = torch.linspace(dx / 2, nr * dx - dx / 2, nr, dtype=torch.float64, device="cuda")
true = 0.15 - torch.sqrt((z[None, :] - 0.47) ** 2 + r[:, None] ** 2)
phi = true * (1 + 0.2 * torch.sin(9 * z[None, :] + 5 * r[:, None]))
rn = NarrowBandReinit(nr, nz)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    p = phi.clone()
    rn(p, dx, 6 * dx)
torch.cuda.synchronize()
print("launches", rn.sweeps)
