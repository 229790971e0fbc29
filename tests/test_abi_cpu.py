"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/axisym_b200.h declares, the ctypes table binds exactly that set, and the product path
fails loudly (no CPU fallback) when there is no CUDA device.  No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from pyaxisymflow_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "axisym_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(axb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/axisym_b200.h but not exported"
    assert names == _lib.exported_names(), "ctypes table and header disagree"
    assert _lib.load().axb_version() == 100
    assert _lib.launch_count() == 0


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_lib.AxbGrid) == 64        # 2*i32, i64, f64, 6*i32, 2*i32 (batch, scalar_stride), i64
    assert _lib.AxbGrid.ld.offset == 8 and _lib.AxbGrid.dx.offset == 16 and _lib.AxbGrid.kz0.offset == 24
    assert _lib.AxbGrid.batch.offset == 48 and _lib.AxbGrid.batch_stride.offset == 56
    assert ctypes.sizeof(_lib.AxbFdPlan) == 8 + 6 * 8 + 2 * 8 + 8 + 8 + 3 * 32 + 2 * 64 + 8 + 4 * 8 + 8 + 8 + 8 + 8 + 8
    assert _lib.AxbFdPlan.leaf_fwd.offset == 8 + 6 * 8 + 2 * 8 + 8 + 8 + 3 * 32


def test_argument_validation_without_a_gpu():
    """error paths that return before any CUDA call"""
    lib = _lib.load()
    g = _lib.AxbGrid(8, 8, 4, 1.0, 0, 8, 0, 8)     # ld < nz
    assert lib.axb_set_fixed_val(ctypes.byref(g), ctypes.c_void_p(16), 1.0, None) == -1
    g = _lib.AxbGrid(8, 8, 8, 1.0, 0, 8, 0, 8)
    assert lib.axb_set_fixed_val(ctypes.byref(g), None, 1.0, None) == -1
    assert lib.axb_set_fixed_val(ctypes.byref(g), ctypes.c_void_p(12), 1.0, None) == -2   # misaligned
    assert lib.axb_dgemm(0, 4, 4, None, 4, None, 4, None, 4, None, None, 0.0, 0.0, None) == -1
    assert lib.axb_ls_workspace_bytes(100, 100) > 100 * 100 * 20
    # the solve's building blocks reject unsupported shapes before touching the device
    p16 = ctypes.c_void_p(16)
    assert lib.axb_dct2_rows(4, 96, p16, 96, ctypes.c_void_p(4096), 96, p16, 1.0, 1.0, None) == -1      # not 2^p
    assert lib.axb_dct3_rows(4, 32768, p16, 32768, ctypes.c_void_p(1 << 20), 32768, p16, None) == -1    # > 16384
    assert lib.axb_dct2_rows(4, 256, p16, 128, ctypes.c_void_p(4096), 256, p16, 1.0, 1.0, None) == -1   # pitch < n
    assert lib.axb_tridiag_solve_factored(1, 64, p16, 64, p16, p16, None) == -1                         # nr < 2
    assert lib.axb_tridiag_solve_factored(8, 40, p16, 40, p16, p16, None) == -1                         # nz % 16
    assert lib.axb_tridiag_partition_correct(8, 64, p16, 32, p16, p16, p16, p16, p16, 4, None) == -1    # pitch < nz
    ptrs = (ctypes.c_uint64 * 2)(16, 0)
    assert lib.axb_peer_block_put(2, 0, ptrs, 0, 64, p16, 0, 64, 2, 64, None) == -1                     # null peer
    assert lib.axb_peer_block_put(17, 0, ptrs, 0, 64, p16, 0, 64, 2, 64, None) == -1                    # > AXB_MAX_PEERS
    # ensembles (axb_grid_t.batch > 1): entries that are not batched refuse, batched ones validate the strides
    g = _lib.AxbGrid(8, 8, 32, 1.0, 0, 8, 0, 8, 0, 0, 4, 24, 8)
    assert lib.axb_set_fixed_val(ctypes.byref(g), p16, 1.0, None) == -3
    assert lib.axb_diffusion_rk2_fused(ctypes.byref(g), p16, ctypes.c_void_p(32), p16, p16, 1.0, 1.0, None, None) == -3
    assert lib.axb_diffusion_rk2_stage1_dev(ctypes.byref(g), p16, ctypes.c_void_p(32), p16, None, p16, None) == -1
    g.batch_stride = 4                                                                                  # < nz
    assert lib.axb_diffusion_rk2_stage1_dev(ctypes.byref(g), p16, ctypes.c_void_p(32), p16, p16, p16, None) == -1
    assert lib.axb_particle_scalars_batched(1, 4, 8, p16, None, 0, .1, .1, 1., 1., 1., 0., 1., None) == -1   # stride < 24
    # SURVEY 8f entries
    g = _lib.AxbGrid(8, 8, 8, 1.0, 0, 8, 0, 8)
    gb = ctypes.byref(g)
    assert lib.axb_velocity_from_phi(gb, p16, p16, None, None) == -1
    assert lib.axb_velocity_from_phi(gb, p16, ctypes.c_void_p(20), p16, None) == -2
    assert lib.axb_baroclinic_vorticity_update(gb, p16, p16, p16, p16, p16, p16, None, None, None, 0.0, 1.0, 3, None) == -1
    assert lib.axb_baroclinic_vorticity_update(gb, p16, p16, p16, p16, p16, p16, None, None, None, 0.0, 1.0, 1, None) == -1
    q32 = ctypes.c_void_p(32)                                                                          # w aliases u_z
    assert lib.axb_baroclinic_vorticity_update(gb, q32, q32, p16, p16, p16, p16, None, None, None, 0.0, 1.0, 0, None) == -1
    info = (ctypes.c_int * 2)()
    need = lib.axb_reinit_workspace_bytes(8, 8)
    assert need >= 2 * 8 * 64 + 64 and lib.axb_reinit_workspace_bytes(0, 8) == 0
    w256 = ctypes.c_void_p(256)
    assert lib.axb_reinit_distance(gb, p16, 0.5, 2, None, None, need, info, None) == -1                 # no workspace
    assert lib.axb_reinit_distance(gb, p16, 0.5, 3, None, w256, need, info, None) == -1                 # order
    assert lib.axb_reinit_distance(gb, p16, 0.0, 2, None, w256, need, info, None) == -1                 # narrow <= 0
    assert lib.axb_reinit_distance(gb, p16, 0.5, 2, None, ctypes.c_void_p(264), need, info, None) == -2 # work alignment
    assert lib.axb_reinit_distance(gb, p16, 0.5, 2, None, w256, need - 1, info, None) == -4             # work too small
    slab = _lib.AxbGrid(8, 8, 8, 1.0, 4, 16, 2, 6)
    assert lib.axb_reinit_distance(ctypes.byref(slab), p16, 0.5, 2, None, w256, need, info, None) == -3 # z-slab


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from pyaxisymflow_b200 import ops
    from pyaxisymflow_b200.fd import FastDiagonalisationStokesSolver
    from pyaxisymflow_b200.timestep import RigidFlowStepper

    a = np.zeros((8, 8))
    with pytest.raises(_lib.AxbError):
        ops.brinkmann_penalize(1.0, 1.0, a, 0.0, 0.0, a, a, a.copy(), a.copy())
    with pytest.raises(_lib.AxbError):
        FastDiagonalisationStokesSolver(8, 16, 1 / 16)
    with pytest.raises(_lib.AxbError):
        RigidFlowStepper(16)


def test_mirrored_module_tree_matches_the_reference_layout():
    import importlib

    for mod, names in {
        "kernels.brinkmann_penalize": ["brinkmann_penalize"],
        "kernels.compute_velocity_from_psi": ["compute_velocity_from_psi_unb", "compute_velocity_from_psi_periodic"],
        "kernels.compute_vorticity_from_velocity": ["compute_vorticity_from_velocity_unb"],
        "kernels.diffusion_RK2": ["diffusion_RK2_unb", "diffusion_RK2_periodic"],
        "kernels.kill_boundary_vorticity_sine": ["kill_boundary_vorticity_sine_r", "kill_boundary_vorticity_sine_z"],
        "kernels.periodic_boundary_ghost_comm": ["gen_periodic_boundary_ghost_comm"],
        "kernels.smooth_Heaviside": ["smooth_Heaviside"],
        "kernels.FastDiagonalisationStokesSolver": ["FastDiagonalisationStokesSolver"],
        "kernels.FastDiagonalisationPotentialSolver": ["FastDiagonalisationPotentialSolver"],
        "kernels.implicit_diffusion_solver": ["ImplicitEulerDiffusionStepper"],
        "kernels.advect_vorticity_via_eno3": ["gen_advect_vorticity_via_eno3"],
        "kernels.advect_particle": ["advect_vorticity_via_particles"],
        "kernels.compute_forces": ["compute_force_on_body"],
        "kernels.force_projection": ["force_projection"],
        "kernels.vortex_stretching": ["vortex_stretching"],
        "kernels.compute_velocity_from_phi": ["compute_velocity_from_phi_unb"],
        "kernels.update_baroclinic_vorticity": ["update_baroclinic_vorticity", "update_baroclinic_vorticity_penal",
                                                "update_baroclinic_vorticity_diff_penal"],
        "utils.dump_vtk": ["vtk_init", "vtk_write"],
        "pyst_kernels.advection_flux": ["gen_advection_flux_conservative_eno3_pyst_kernel"],
        "pyst_kernels.advection_timestep": ["gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel"],
        "pyst_kernels.elementwise_ops": ["gen_elementwise_sum_pyst_kernel", "gen_set_fixed_val_pyst_kernel"],
        "elasto_kernels.advect_refmap_via_eno3": ["gen_advect_refmap_via_eno3"],
        "elasto_kernels.solid_sigma": ["solid_sigma"],
        "elasto_kernels.div_tau": ["update_vorticity_from_solid_stress"],
        "elasto_kernels.extrapolate_eta_using_least_squares_unb": ["extrapolate_eta_with_least_squares"],
        # every name of core/src/instantiate.yml:1-34
        "core.mesh_to_particles": [
            "mesh_to_particles_1D_mp4", "wrap_particles_around_1D_domain", "mesh_to_particles_2D_linear_kernel",
            "mesh_to_particles_2D_mp4", "mesh_to_particles_2D_yang_smooth_three_point_kernel",
            "mesh_to_particles_2D_mp6", "wrap_particles_around_2D_domain",
            "mesh_to_particles_2D_unbounded_linear_kernel",
            "mesh_to_particles_2D_unbounded_yang_smooth_three_point_kernel", "mesh_to_particles_2D_unbounded_mp4",
            "mesh_to_particles_2D_unbounded_mp6"],
        "core.particles_to_mesh": [
            "particles_to_mesh_1D_mp4", "particles_to_mesh_2D_linear_kernel",
            "particles_to_mesh_2D_yang_smooth_three_point_kernel", "particles_to_mesh_2D_mp4",
            "particles_to_mesh_2D_mp6", "particles_to_mesh_2D_unbounded_linear_kernel",
            "particles_to_mesh_2D_unbounded_yang_smooth_three_point_kernel", "particles_to_mesh_2D_unbounded_mp4",
            "particles_to_mesh_2D_unbounded_mp6"],
        "core.extrapolate_using_least_squares": ["extrapolate_using_least_squares_till_first_order",
                                                 "extrapolate_using_least_squares_till_second_order"],
    }.items():
        m = importlib.import_module("pyaxisymflow_b200." + mod)
        for n in names:
            assert callable(getattr(m, n)), f"{mod}.{n}"


def test_edge_block_enumeration_of_the_marching_stencil_passes():
    """The row-marching stencil passes launch their edge kernel on a compact grid of the edge blocks only, enumerated
    on the host with the kernels' own predicates (csrc/stencils_march.cu: edge_map / edge_decode).  A block the
    enumeration missed would be computed by nobody, so the launched set is compared with the definition -- every
    (column block, fine row chunk) whose parent chunk is not interior -- on single-GPU grids, z-slabs (owned column
    window, global offset) and r-slabs (owned row window, which only the kernels with fused reductions honour)."""
    from pyaxisymflow_b200.device import make_grid

    lib = _lib.load()
    MT = 128

    def interior(g, bx, p0, p1, rb, hr, hz, vec, owned):
        kb0 = 2 * bx * MT
        kb1 = kb0 + 2 * MT
        ju0, ju1 = (g.ju0, g.ju1) if g.ju1 else (0, g.nr)
        cols = (kb0 >= g.ku0 and kb1 <= g.ku1 and kb0 + g.kz0 >= hz and kb1 - 1 + g.kz0 <= g.nz_global - 1 - hz
                and kb0 - hz >= 0 and kb1 - 1 + hz < g.nz)
        rows = p1 - p0 == rb and p0 >= hr and p1 + hr <= g.nr and (not owned or (p0 >= ju0 and p1 <= ju1))
        return bool(vec and cols and rows)

    rng = np.random.default_rng(5)
    grids = [make_grid(4096, 16384, 16384, 1.0), make_grid(128, 256, 256, 1.0), make_grid(1024, 4096, 4096, 1.0),
             make_grid(516, 16384, 16384, 1.0, rows=(2, 514)), make_grid(2052, 16384, 16384, 1.0, rows=(2, 2050)),
             make_grid(1024, 2052, 2052, 1.0, slab=(0, 16384, 0, 2050)),
             make_grid(1024, 2052, 2052, 1.0, slab=(14332, 16384, 2, 2052)),
             make_grid(1024, 2048, 8 * 2048, 1.0, batch=(8, 2048, 24)), make_grid(3, 12, 12, 1.0),
             make_grid(37, 1030, 1030, 1.0), make_grid(1000, 3000, 3000, 1.0, rows=(100, 433))]
    for _ in range(40):
        nr, nz = int(rng.integers(3, 700)), int(rng.integers(4, 5000))
        grids.append(make_grid(nr, nz, nz, 1.0, rows=(int(rng.integers(0, nr // 2)), int(rng.integers(nr // 2 + 1, nr + 1)))))
    seen_compact = 0
    for g in grids:
        for hr, hz in ((1, 1), (2, 2)):
            for vec in (1, 0):
                for owned in (0, 1):
                    cap = 1 << 18
                    pairs = np.zeros(2 * cap, dtype=np.int32)
                    info = np.zeros(4, dtype=np.int32)
                    rc = lib.axb_debug_edge_blocks(ctypes.byref(g), hr, hz, vec, owned, pairs.ctypes.data, cap,
                                                   info.ctypes.data)
                    assert rc == 0
                    rb, re, compact, n = (int(v) for v in info)
                    assert rb % re == 0 and re in (4, 8, 16, 32) and n <= cap
                    seen_compact += compact
                    launched = [(int(pairs[2 * i]), int(pairs[2 * i + 1])) for i in range(n)]
                    nbx, nfy = ((g.nz + 1) // 2 + MT - 1) // MT, (g.nr + re - 1) // re
                    want = set()
                    for bx in range(nbx):
                        for fy in range(nfy):
                            p0 = (fy * re // rb) * rb
                            if not interior(g, bx, p0, min(p0 + rb, g.nr), rb, hr, hz, vec, owned):
                                want.add((bx, fy))
                    edge = [p for p in launched
                            if not interior(g, p[0], (p[1] * re // rb) * rb, min((p[1] * re // rb) * rb + rb, g.nr), rb, hr,
                                            hz, vec, owned)]
                    assert len(edge) == len(set(edge)), "an edge block is launched twice"
                    assert set(edge) == want, (g.nr, g.nz, hr, hz, vec, owned)
                    if compact:
                        assert len(launched) == len(want), "the compact grid launches interior blocks"
    assert seen_compact > 50
