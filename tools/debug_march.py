import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from pyaxisymflow_b200 import _lib
import pyaxisymflow_b200.ops as K

def grid(nr, nz):
    dx = 1.0 / nz
    z = np.linspace(dx / 2, nz * dx - dx / 2, nz); r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    Z, R = np.meshgrid(z, r); return dx, Z, R

rng = np.random.default_rng(0)
for nr, nz in [(24, 56), (24, 64), (16, 256), (40, 600)]:
    dx, Z, R = grid(nr, nz)
    psi = rng.standard_normal((nr, nz))
    res = {}
    for path in (1, 0):
        _lib.call("axb_set_stencil_path", path)
        uz, ur = np.zeros((nr, nz)), np.zeros((nr, nz))
        K.compute_velocity_from_psi_unb(uz, ur, psi, R, dx)
        w, tmp = psi.copy(), np.zeros((nr, nz))
        K.diffusion_RK2_unb(w, tmp, R, 1e-3, 0.1 * dx * dx / 1e-3, dx)
        res[path] = (uz, ur, w, tmp)
    for name, a, b in zip(("uz", "ur", "w", "tmp"), res[0], res[1]):
        bad = np.argwhere(np.abs(a - b) > 1e-12 * np.abs(b).max())
        print(nr, nz, name, "mismatches", len(bad), "rows", sorted(set(bad[:, 0]))[:12], "cols", sorted(set(bad[:, 1]))[:40])
        if len(bad):
            j, k = bad[0]
            print("   first", j, k, a[j, k], b[j, k])
