"""Generate the fixtures of the SURVEY.md 8f rows from the UNMODIFIED reference (build container only).

Run from the repo root:   python tests/golden/make_golden_widen.py
Same rules as make_golden.py: /root/reference is imported, never copied; the GPU box only reads
the committed .npz files.  Kept separate so that the round-1 fixtures stay byte-identical.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, grid, save, smooth  # noqa: E402

warnings.filterwarnings("ignore")
sys.path.insert(0, REF)


def main():
    from pyaxisymflow.kernels.compute_velocity_from_phi import compute_velocity_from_phi_unb
    from pyaxisymflow.kernels.update_baroclinic_vorticity import (
        update_baroclinic_vorticity, update_baroclinic_vorticity_diff_penal, update_baroclinic_vorticity_penal)

    rng = np.random.default_rng(20261018)
    nr, nz = 24, 56
    dx, z, r, Z, R = grid(nr, nz)

    # ---- 8f-2 velocity of a potential ----------------------------------------------------
    phi = smooth(rng, Z, R)
    uz, ur = rng.standard_normal(Z.shape), rng.standard_normal(Z.shape)
    compute_velocity_from_phi_unb(uz, ur, phi, dx)
    save("velocity_from_phi", phi=phi, dx=dx, uz=uz, ur=ur)

    # ---- 8f-4 baroclinic vorticity source (three variants) -------------------------------
    u_z, u_r = smooth(rng, Z, R), smooth(rng, Z, R)
    o_z, o_r = u_z + 1e-3 * smooth(rng, Z, R), u_r + 1e-3 * smooth(rng, Z, R)
    rho = 1.0 + 0.5 * np.clip(smooth(rng, Z, R) + 0.5, 0, 1)
    p_z, p_r = smooth(rng, Z, R, 5.0), smooth(rng, Z, R, 5.0)
    w0 = smooth(rng, Z, R, 3.0)
    dt, nu = 2e-3, 1e-2
    w_a, w_b, w_c = w0.copy(), w0.copy(), w0.copy()
    update_baroclinic_vorticity(w_a, u_z, u_r, o_z, o_r, rho, dt, dx)
    update_baroclinic_vorticity_penal(w_b, u_z, u_r, o_z, o_r, rho, p_z, p_r, dt, dx)
    update_baroclinic_vorticity_diff_penal(w_c, u_z, u_r, o_z, o_r, rho, p_z, p_r, R, nu, dt, dx)
    save("baroclinic", w0=w0, u_z=u_z, u_r=u_r, o_z=o_z, o_r=o_r, rho=rho, p_z=p_z, p_r=p_r, dt=dt, nu=nu, dx=dx,
         w_plain=w_a, w_penal=w_b, w_diff_penal=w_c)


if __name__ == "__main__":
    main()
