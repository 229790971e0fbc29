"""Generate tests/golden/eno3.npz by running the reference's OWN pystencils kernel definitions
(pyst_kernels/advection_flux.py, advection_timestep.py, elementwise_ops.py and the wrappers
kernels/advect_vorticity_via_eno3.py, elasto_kernels/advect_refmap_via_eno3.py), imported
unmodified from /root/reference, through tests/pystencils_shim.py (pystencils 1.0.1 itself is not
installable offline).  Build container only:

    python tests/golden/make_golden_eno3.py

Every entry stores the seeded inputs and the outputs, so the tests never re-create inputs.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AXB_REFERENCE", "/root/reference")
warnings.filterwarnings("ignore")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import pystencils_shim

    shimmed = pystencils_shim.install()
    from pyaxisymflow.pyst_kernels.advection_flux import (
        gen_advection_flux_conservative_eno3_pyst_kernel, gen_advection_flux_non_conservative_eno3_pyst_kernel)
    from pyaxisymflow.pyst_kernels.advection_timestep import (
        gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel,
        gen_advection_timestep_euler_forward_non_conservative_eno3_pyst_kernel)
    from pyaxisymflow.pyst_kernels.elementwise_ops import (
        gen_elementwise_sum_pyst_kernel, gen_set_fixed_val_pyst_kernel)
    from pyaxisymflow.kernels.advect_vorticity_via_eno3 import (
        gen_advect_vorticity_via_eno3, gen_advect_vorticity_via_eno3_periodic)
    from pyaxisymflow.elasto_kernels.advect_refmap_via_eno3 import (
        gen_advect_refmap_via_eno3, gen_advect_refmap_via_eno3_periodic)
    from pyaxisymflow.kernels.periodic_boundary_ghost_comm import (
        gen_periodic_boundary_ghost_comm, gen_periodic_boundary_ghost_comm_eta)

    rng = np.random.default_rng(20261018)
    out = {"generated_with_shim": np.array(shimmed)}

    # ---- a3 / a4: raw flux closures on plain arrays; flux pre-filled so that `+=` and the untouched rim show ----
    n0, n1 = 22, 30
    y, x = np.meshgrid(np.arange(n0) / n0, np.arange(n1) / n1, indexing="ij")
    field = np.sin(2 * np.pi * (x + 0.4 * y)) + 0.3 * rng.standard_normal((n0, n1))
    vel = np.empty((2, n0, n1))
    vel[0] = np.cos(2 * np.pi * (2 * x - y)) + 0.2 * rng.standard_normal((n0, n1))      # sign changes: both branches
    vel[1] = np.sin(2 * np.pi * (x + 3 * y)) + 0.2 * rng.standard_normal((n0, n1))
    flux0 = rng.standard_normal((n0, n1))
    inv_dx = 0.37
    out.update(raw_field=field, raw_vel=vel, raw_flux0=flux0, raw_inv_dx=inv_dx)
    for tag, gen in (("cons", gen_advection_flux_conservative_eno3_pyst_kernel),
                     ("noncons", gen_advection_flux_non_conservative_eno3_pyst_kernel)):
        for fixed in (False, (n0, n1)):
            k = gen(real_t=np.float64, num_threads=False, fixed_grid_size=fixed)
            flux = flux0.copy()
            k(advection_flux=flux, field=field, velocity=vel, inv_dx=inv_dx)
            out[f"raw_flux_{tag}" + ("_fixed" if fixed else "")] = flux
        assert np.array_equal(out[f"raw_flux_{tag}"], out[f"raw_flux_{tag}_fixed"])
        del out[f"raw_flux_{tag}_fixed"]
        # a fixed grid size that does not match raises ValueError (pystencils shape check)
        try:
            gen(fixed_grid_size=(n0 + 1, n1))(advection_flux=flux0.copy(), field=field, velocity=vel, inv_dx=inv_dx)
            raise SystemExit("expected ValueError")
        except ValueError:
            pass

    # ---- a5 / a6: Euler-forward closures ----
    dt_by_dx = 0.21
    out["step_dt_by_dx"] = dt_by_dx
    for tag, gen in (("cons", gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel),
                     ("noncons", gen_advection_timestep_euler_forward_non_conservative_eno3_pyst_kernel)):
        k = gen(real_t=np.float64, num_threads=False, fixed_grid_size=False)
        f, flux = field.copy(), flux0.copy()
        k(field=f, advection_flux=flux, velocity=vel, dt_by_dx=dt_by_dx)
        out[f"step_field_{tag}"], out[f"step_flux_{tag}"] = f, flux

    # ---- a1 / a2: elementwise closures (scalar and vector flavours) ----
    a, b = rng.standard_normal((n0, n1)), rng.standard_normal((n0, n1))
    s = rng.standard_normal((n0, n1))
    gen_elementwise_sum_pyst_kernel()(sum_field=s, field_1=a, field_2=b)
    va, vb = rng.standard_normal((2, n0, n1)), rng.standard_normal((2, n0, n1))
    vs = np.zeros((2, n0, n1))
    gen_elementwise_sum_pyst_kernel(field_type="vector")(sum_field=vs, field_1=va, field_2=vb)
    alias = a.copy()
    gen_elementwise_sum_pyst_kernel()(sum_field=alias, field_1=alias, field_2=b)       # in-place use of the timestep
    fill = rng.standard_normal((n0, n1))
    gen_set_fixed_val_pyst_kernel()(field=fill, fixed_val=-2.5)
    vfill = rng.standard_normal((2, n0, n1))
    gen_set_fixed_val_pyst_kernel(field_type="vector")(vector_field=vfill, fixed_vals=[1.25, -0.75])
    out.update(ew_a=a, ew_b=b, ew_sum=s, ew_va=va, ew_vb=vb, ew_vsum=vs, ew_alias=alias, ew_fill=fill,
               ew_vfill=vfill)

    # ---- a7: vorticity advection wrappers (mirrored domain) ----
    nr, nz = 24, 56
    dx = 1.0 / nz
    z = np.linspace(dx / 2, 1 - dx / 2, nz)
    r = np.linspace(dx / 2, nr * dx - dx / 2, nr)
    Z, R = np.meshgrid(z, r)
    w0 = 3.0 * np.sin(2 * np.pi * (Z + 0.3 * R)) * np.exp(-((Z - 0.5) ** 2 + R ** 2) / 0.05) \
        + 0.1 * rng.standard_normal(Z.shape)
    uz0 = np.cos(2 * np.pi * Z) * (1 - R) + 0.05 * rng.standard_normal(Z.shape)
    ur0 = np.sin(4 * np.pi * Z) * R + 0.05 * rng.standard_normal(Z.shape)
    dt = 0.3 * dx / np.amax(np.abs(uz0) + np.abs(ur0))
    out.update(adv_w0=w0, adv_uz0=uz0, adv_ur0=ur0, adv_dt=dt, adv_dx=dx)
    w = w0.copy()
    gen_advect_vorticity_via_eno3(dx, nr, nz, num_threads=False)(w, uz0.copy(), ur0.copy(), dt)
    out["adv_w_unb"] = w
    per = gen_periodic_boundary_ghost_comm(2)
    w, uz, ur = w0.copy(), uz0.copy(), ur0.copy()
    gen_advect_vorticity_via_eno3_periodic(dx, nr, nz, per)(w, uz, ur, dt)
    out.update(adv_w_per=w, adv_uz_per=uz, adv_ur_per=ur)
    # three steps in a row (the flux scratch and the doubled arrays are reused between calls)
    w = w0.copy()
    adv = gen_advect_vorticity_via_eno3(dx, nr, nz)
    for _ in range(3):
        adv(w, uz0, ur0, dt)
    out["adv_w_unb_3steps"] = w

    # ---- a17: reference-map advection wrappers ----
    e1 = Z + 0.02 * np.sin(2 * np.pi * Z) * np.cos(3 * R) + 0.002 * rng.standard_normal(Z.shape)
    e2 = R * (1 + 0.05 * np.cos(6 * Z)) + 0.002 * rng.standard_normal(Z.shape)
    out.update(ref_e1_0=e1, ref_e2_0=e2)
    a1, a2 = e1.copy(), e2.copy()
    gen_advect_refmap_via_eno3(dx, nr, nz)(a1, a2, uz0.copy(), ur0.copy(), dt)
    out.update(ref_e1_unb=a1, ref_e2_unb=a2)
    per_eta = gen_periodic_boundary_ghost_comm_eta(2, 1.0, dx)
    a1, a2, uz, ur = e1.copy(), e2.copy(), uz0.copy(), ur0.copy()
    gen_advect_refmap_via_eno3_periodic(dx, nr, nz, per, per_eta)(a1, a2, uz, ur, dt)
    out.update(ref_e1_per=a1, ref_e2_per=a2, ref_uz_per=uz, ref_ur_per=ur, ref_z_max=1.0)

    np.savez_compressed(os.path.join(HERE, "eno3.npz"), **out)
    print("wrote eno3.npz", sum(np.asarray(v).nbytes for v in out.values()), "B raw; shim used:", shimmed)


if __name__ == "__main__":
    main()
