/*
 * axisym_b200.h -- C ABI of libaxisym_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for the per-timestep hot path of PyAxisymFlow.
 * The reference has no FFI of its own on this path: its "kernels" are Python callables
 * (numba / pystencils closures, NumPy classes, three pybind11 modules).  Each entry point
 * below therefore cites the reference *callable* it replaces (file:line under
 * /root/reference); the ctypes stubs that bind them live in pyaxisymflow_b200/_lib.py and
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; fields are float64,
 *     C-order (nr, nz), axis 1 (z) contiguous, row pitch `ld` elements;
 *   - no entry point allocates, synchronises or throws; work is enqueued on `stream`;
 *   - return value: 0 = ok, negative = AXB_E* argument error, positive = cudaError_t;
 *   - scalars that a device-resident driver keeps on the GPU come in pairs
 *     (`double x, const double* x_dev`): when x_dev != NULL the kernel reads *x_dev.
 *   - `r1d` / `z1d` are the 1-D cell-centre coordinates (R[:,0], Z[0,:] of the reference's
 *     meshgrid arrays), nr resp. nz doubles.
 */
#ifndef AXISYM_B200_H
#define AXISYM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* axb_stream_t;

#define AXB_OK 0
#define AXB_EINVAL (-1)   /* null pointer / bad shape */
#define AXB_EALIGN (-2)   /* pointer or pitch not 8-byte aligned */
#define AXB_ENOSUP (-3)   /* unsupported configuration */
#define AXB_EWORK (-4)    /* workspace too small */

/* Shape + z-slab placement of one (nr, nz) field.  Single GPU: kz0 = 0, nz_global = nz,
 * ku0 = 0, ku1 = nz.  Under z-slab decomposition `nz` counts the locally STORED columns
 * (owned + halos), `kz0` is the global z index of local column 0 (negative on the first
 * rank's left halo), and [ku0, ku1) is the locally owned range the kernel may write.
 * Under r-slab (row) decomposition every rank's (owned + halo rows, nz) block is an ordinary field for the
 * kernels (r1d holds the block's own radii); [ju0, ju1) are the OWNED rows, the only ones that count in the
 * fused reductions (CFL maximum, drag sum).  ju1 == 0 means all rows (single GPU, z-slabs).
 * Ensembles: batch > 1 describes `batch` independent members of the same shape whose fields
 * lie batch_stride ELEMENTS apart (member m of a field = pointer + m * batch_stride; batch_stride = nz for members
 * stored as column blocks of one (nr, batch nz) array with ld = batch nz) and whose device-resident scalars
 * (dt_dev, U_dev, reduction targets, ...) lie scalar_stride doubles apart.  One launch then serves every member
 * (the member index is the grid's z dimension); the 1-D coordinate arrays are shared.  Entries that are not
 * documented as "batched" refuse batch > 1 with AXB_ENOSUP.  batch = 0 or 1: a single field. */
typedef struct axb_grid {
  int32_t nr;
  int32_t nz;
  int64_t ld;
  double dx;
  int32_t kz0;
  int32_t nz_global;
  int32_t ku0;
  int32_t ku1;
  int32_t ju0;
  int32_t ju1;
  int32_t batch;
  int32_t scalar_stride;
  int64_t batch_stride;
} axb_grid_t;

int axb_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t axb_launch_count(void);
/* 0 (default): row-marching kernels for G-VEL / G-PEN / G-ADV / G-REF / G-DIF (stencils_march.cu);
 * 1: the 2-D tiled kernels (stencils.cu, eno3.cu) that repeat the reference's divisions bit for bit. */
int axb_set_stencil_path(int legacy_tiled);
/* 1: row-marching kernels for G-SOL-1 / G-SOL-2 (axb_solid_sigma, axb_solid_tau) on grids with >= 258 columns;
 * 0 (default): the 2-D tiled kernels.  Bit-identical results; see DESIGN.md section 6.1. */
int axb_set_solid_march(int on);
/* factored tridiagonal sweeps (axb_tridiag_solve_factored and the solves built on it): 0 (default): the
 * warp-specialised kernel -- a producer lane feeds the TMA ring (8 / 16 / 32 boxes by column count), a second warp
 * runs the chain and stores the rows; 1 (or the environment variable AXB_TRI_ONE_WARP=1 at load): the single-warp TMA
 * kernel.  Bit-identical results. */
int axb_set_tridiag_sweep(int one_warp);
/* Test hook, no kernel launch (runs without a GPU): the blocks the EDGE kernel of a row-marching stencil pass
 * (G-VEL / G-PEN / G-ADV / G-REF / G-DIF; halo hr rows, hz columns) is launched on for grid g -- (column block of 256
 * columns, fine row chunk) pairs -- and info = {rows per interior chunk, rows per edge chunk, 1 if the compact edge grid
 * is used, launched blocks}.  The interior blocks of the same decomposition belong to the interior kernel. */
int axb_debug_edge_blocks(const axb_grid_t* g, int hr, int hz, int vec, int owned_window, int32_t* pairs, int cap,
                          int32_t* info);

/* ---- G-BND: kernels/kill_boundary_vorticity_sine.py:4-14 and :17-27 ------------------ */
int axb_kill_boundary_vorticity_sine_z(const axb_grid_t* g, double* w, const double* z1d, int width,
                                       axb_stream_t s);
int axb_kill_boundary_vorticity_sine_r(const axb_grid_t* g, double* w, const double* r1d, int width,
                                       axb_stream_t s);
/* the two halves of kill_boundary_vorticity_sine_r for r-slabs: parts bit 0 = the sine ramp over the last
 * `width` rows (the rank that holds r_max), bit 1 = row 0 := 0 (the rank that holds the axis) */
int axb_kill_boundary_vorticity_sine_r_parts(const axb_grid_t* g, double* w, const double* r1d, int width,
                                             int parts, axb_stream_t s);

/* ---- a12: kernels/periodic_boundary_ghost_comm.py:4-15 (z_max = two_g_dx = 0) and the
 *      reference-map form :18-33: left ghosts = (src - z_max) + two_g_dx,
 *      right ghosts = (src + z_max) - two_g_dx  (same operation order as the reference) ---- */
int axb_periodic_ghost_comm(const axb_grid_t* g, double* f, int ghost, double z_max, double two_g_dx,
                            axb_stream_t s);

/* ---- G-VEL: kernels/compute_velocity_from_psi.py:4-17.  Fused extras for the device
 *      driver: adds the free stream (examples/FlowPastSphere/flow_past_sphere.py:117-121),
 *      and, if umax_out != NULL, atomically maxes |u_z|+|u_r| into it (the dt reduction of
 *      flow_past_sphere.py:150-153).  add_dev, if given, holds {uz_add, ur_add}. ---------- */
int axb_velocity_from_psi(const axb_grid_t* g, double* u_z, double* u_r, const double* psi,
                          const double* r1d, double uz_add, double ur_add, const double* add_dev,
                          double* umax_out, axb_stream_t s);

/* ---- a9: kernels/brinkmann_penalize.py:4-16.  U_*_field non-NULL selects the field form
 *      (examples/TorusParticleTransport passes arrays), else the scalars are used. --------- */
int axb_brinkmann_penalize(const axb_grid_t* g, double lam, double dt, const double* chi, double U_z,
                           double U_r, const double* U_z_field, const double* U_r_field,
                           const double* grid_u_z, const double* grid_u_r, double* pen_u_z,
                           double* pen_u_r, axb_stream_t s);

/* ---- a10: kernels/compute_vorticity_from_velocity.py:4-13.  If u_z_sub / u_r_sub are
 *      non-NULL the curl of (u - u_sub) is taken, which is how every driver calls it
 *      (flow_past_sphere.py:161-163).  accumulate != 0 does vort += curl instead of =. ----- */
int axb_vorticity_from_velocity(const axb_grid_t* g, double* vort, const double* u_z, const double* u_r,
                                const double* u_z_sub, const double* u_r_sub, int accumulate,
                                axb_stream_t s);

/* ---- G-PEN: the whole penalisation block of flow_past_sphere.py:155-175 in one pass:
 *      u = pen(u_upen), w += curl(u - u_upen) on the interior, and (sum_out != NULL)
 *      sum_out += sum(R * chi * (u_z - U_z))  [drag numerator / compute_forces.py:14]. ----- */
int axb_penalise_update_vorticity(const axb_grid_t* g, double* u_z, double* u_r, double* w,
                                  const double* u_z_upen, const double* u_r_upen, const double* chi,
                                  double lam, double dt, const double* dt_dev, double U_z, double U_r,
                                  const double* U_dev, const double* r1d, double* sum_out,
                                  axb_stream_t s);

/* ---- G-ADV: kernels/advect_vorticity_via_eno3.py:21-42 = axis mirror + conservative ENO3
 *      Euler step (pyst_kernels/advection_timestep.py:45-54) + copy back, as ONE kernel with
 *      the reflection done by index.  Out of place: w_out must not alias w_in. ------------- */
int axb_advect_vorticity_eno3(const axb_grid_t* g, double* w_out, const double* w_in, const double* u_z,
                              const double* u_r, double dt, const double* dt_dev, axb_stream_t s);

/* ---- G-REF: elasto_kernels/advect_refmap_via_eno3.py:21-52 (two non-conservative ENO3
 *      steps sharing the velocity), one kernel, out of place. ------------------------------ */
int axb_advect_refmap_eno3(const axb_grid_t* g, double* eta1_out, double* eta2_out, const double* eta1,
                           const double* eta2, const double* u_z, const double* u_r, double dt,
                           const double* dt_dev, axb_stream_t s);

/* ---- a1-a6: the pystencils closures on plain (n0, n1) arrays, no mirroring -------------
 *      axb_eno3_flux         : pyst_kernels/advection_flux.py:133-161 / :302-330 (flux += ...)
 *      axb_eno3_euler_step   : pyst_kernels/advection_timestep.py:37-54 / :85-102
 *      axb_elementwise_sum   : pyst_kernels/elementwise_ops.py:28-33
 *      axb_set_fixed_val     : pyst_kernels/elementwise_ops.py:73-77
 *      (g->nr, g->nz) is the array shape; only [2:-2, 2:-2] is touched by the flux. ------- */
int axb_eno3_flux(const axb_grid_t* g, double* flux, const double* field, const double* vel0,
                  const double* vel1, double inv_dx, int conservative, axb_stream_t s);
int axb_eno3_euler_step(const axb_grid_t* g, double* field_out, const double* field_in, const double* vel0,
                        const double* vel1, double dt_by_dx, int conservative, axb_stream_t s);
int axb_elementwise_sum(const axb_grid_t* g, double* sum, const double* f1, const double* f2,
                        axb_stream_t s);
int axb_set_fixed_val(const axb_grid_t* g, double* f, double val, axb_stream_t s);

/* ---- G-DIF: kernels/diffusion_RK2.py:4-45 as two launches.
 *      stage1: tmp = w; tmp[int] += 0.5*nu*dt*L(w).   stage2: w[int] += nu*dt*L(tmp). ------- */
int axb_diffusion_rk2_stage1(const axb_grid_t* g, double* tmp, const double* w, const double* r1d,
                             double nu, double dt, const double* dt_dev, axb_stream_t s);
/* stage2 writes w = w_src + nu*dt*L(tmp); w_src == w is the reference's in-place form, a distinct
 * w_src lets a ping-pong driver land the result in another buffer at no extra traffic. */
int axb_diffusion_rk2_stage2(const axb_grid_t* g, double* w, const double* w_src, const double* tmp,
                             const double* r1d, double nu, double dt, const double* dt_dev,
                             axb_stream_t s);
/* The same two stages with nu AND dt read from device memory (batched: member m reads nu_dev[m scalar_stride],
 * dt_dev[m scalar_stride]) -- the members of a particle ensemble differ in nu. */
int axb_diffusion_rk2_stage1_dev(const axb_grid_t* g, double* tmp, const double* w, const double* r1d,
                                 const double* nu_dev, const double* dt_dev, axb_stream_t s);
int axb_diffusion_rk2_stage2_dev(const axb_grid_t* g, double* w, const double* w_src, const double* tmp,
                                 const double* r1d, const double* nu_dev, const double* dt_dev, axb_stream_t s);
/* Both stages in one pass over HBM (16 instead of 40 B/pt): w = w_src + nu dt L(tmp), tmp = w_src + nu dt/2
 * L(w_src) held on chip (row-marching kernels; needs a width-2 halo of w_src on z-slabs).  Same bits as
 * stage1 followed by stage2.  tmp is only used (as scratch) by the 2-D tiled code path; w != w_src. */
int axb_diffusion_rk2_fused(const axb_grid_t* g, double* w, const double* w_src, double* tmp, const double* r1d,
                            double nu, double dt, const double* dt_dev, axb_stream_t s);

/* ---- G-HEAV: kernels/smooth_Heaviside.py:5-14; the _sphere form builds
 *      phi = radius - sqrt((Z-z_cm)^2 + (R-r_cm)^2) in-kernel (flow_past_sphere.py:80-82). -- */
int axb_smooth_heaviside(const axb_grid_t* g, double* H, const double* phi, double blend_w,
                         axb_stream_t s);
int axb_smooth_heaviside_sphere(const axb_grid_t* g, double* H, double* phi_out, const double* z1d,
                                const double* r1d, double z_cm, double r_cm, double radius,
                                double blend_w, axb_stream_t s);

/* ---- kernels/vortex_stretching.py:4-11 -------------------------------------------------- */
int axb_vortex_stretching(const axb_grid_t* g, double* w, const double* u_r, const double* r1d, double dt,
                          axb_stream_t s);

/* ---- kernels/compute_velocity_from_phi.py:4-17 (velocity of a potential:
 *      u_z = d(phi)/dz, u_r = d(phi)/dr; centred inside, 2nd-order one-sided at the ends). ---- */
int axb_velocity_from_phi(const axb_grid_t* g, double* u_z, double* u_r, const double* phi, axb_stream_t s);

/* ---- kernels/update_baroclinic_vorticity.py.  mode 0 = update_baroclinic_vorticity
 *      (:5-35), 1 = ..._penal (:38-67; penal_z/penal_r required), 2 = ..._diff_penal (:70-127;
 *      also r1d and nu).  w[1:-1,1:-1] += dt (Du_z/Dt d(rho)/dr - Du_r/Dt d(rho)/dz) / rho. ---- */
int axb_baroclinic_vorticity_update(const axb_grid_t* g, double* w, const double* u_z, const double* u_r,
                                    const double* old_u_z, const double* old_u_r, const double* density,
                                    const double* penal_z, const double* penal_r, const double* r1d, double nu,
                                    double dt, int mode, axb_stream_t s);

/* ---- narrow-band level-set re-initialisation, the drivers' third-party call
 *      `skfmm.distance(phi, dx=dx, narrow=narrow)` (soft_sphere_streaming.py:196-199; scikit-fmm
 *      2022.8.15, poetry.lock:628-629 -- not vendored, parity unpinned).  In place: cells the marcher
 *      would return unmasked (accepted |distance| <= narrow plus the tentative ring) are overwritten with
 *      their signed distance, the others keep their value (`ball_phi[mask] = bad_phi[mask]`).
 *      mask_out (may be NULL): nr*nz bytes, 1 = masked.  order = 1 or 2 (skfmm default 2).
 *      work: axb_reinit_workspace_bytes(nr, nz) bytes, 256-byte aligned.  Whole-domain grids only.
 *      SYNCHRONISES the stream (convergence flag).  info_host[0] = sweeps, info_host[1] = status bits:
 *      1 no zero contour (skfmm: ValueError), 2 negative discriminant (skfmm: RuntimeError),
 *      4 no fixed point within the sweep bound. ---- */
int64_t axb_reinit_workspace_bytes(int nr, int nz);
int axb_reinit_distance(const axb_grid_t* g, double* phi, double narrow, int order, unsigned char* mask_out,
                        void* work, int64_t work_bytes, int* info_host, axb_stream_t s);

/* ---- a15 diagnostics.  out is a device double; the caller zeroes it (axb_fill_scalars).
 *      max_abs_sum : max(|a| + |b|)           (flow_past_sphere.py:152; b may be NULL)
 *      max         : max(a)                   (flow_past_sphere.py:191)
 *      weighted_sum: sum(r * c * (a - off))   (compute_forces.py:14, force_projection.py:12-13) */
int axb_reduce_max_abs_sum(const axb_grid_t* g, const double* a, const double* b, double* out,
                           axb_stream_t s);
int axb_reduce_max(const axb_grid_t* g, const double* a, double* out, axb_stream_t s);
int axb_reduce_weighted_sum(const axb_grid_t* g, const double* r1d, const double* c, const double* a,
                            double off, double* out, axb_stream_t s);
int axb_fill_scalars(double* dst, int n, double val, axb_stream_t s);

/* ---- device-side scalar glue of the rigid-flow loop (flow_past_sphere.py:117-121,
 *      :150-153, :186-188) so that a whole timestep needs no host round trip.
 *      state (device doubles): [0]=t [1]=dt [2]=umax [3]=sum(R chi u_z) [4]=uz_add [5]=ur_add
 *      [6]=iteration count [7]=last Cd numerator copy.
 *      phase 0: uz_add = U0 * (t < T_ramp ? sin(pi/2 t/T_ramp) : 1);
 *               ur_add = U0 * (t < T_ramp ? ur_ramp * sin(pi t/T_ramp) : 0)
 *               (periodic_flow_past_sphere.py:108-115); umax = 0; sum = 0
 *      phase 1: dt = min(dt_diff_limit, CFL*dx/(umax + eps))
 *      phase 2: t += dt; it += 1; state[7] = state[3] ------------------------------------------ */
int axb_rigid_flow_scalars(int phase, double* state, double U0, double T_ramp, double ur_ramp,
                           double dt_diff_limit, double cfl_dx, axb_stream_t s);

/* ---- driver glue of the soft-sphere and particle loops ---------------
 *      axb_axpy                  : y += a*x  (running averages, soft_sphere_streaming.py:179-180,
 *                                  particle_in_bubble_oscillatory_flow.py:297-299)
 *      axb_pin_level_set         : phi_orig = r_ball - sqrt((eta1-z_cm)^2 + (eta2-r_cm)^2);
 *                                  phi[phi > thresh] = phi_orig      (soft_sphere_streaming.py:191-193)
 *      axb_smooth_heaviside_mask : H = smooth_Heaviside(phi), mask = H > thresh (or >=) as a dense
 *                                  (nr, nz) uint8                   (soft_sphere_streaming.py:205-206)
 *      axb_add_bubble_flow       : breathing mode inside the bubble + exterior potential flow added
 *                                  to (u_z, u_r)         (particle_in_bubble_oscillatory_flow.py:273-294) */
int axb_axpy(const axb_grid_t* g, double* y, const double* x, double a, const double* a_dev, axb_stream_t s);
int axb_pin_level_set(const axb_grid_t* g, double* phi, double* phi_orig, const double* eta1, const double* eta2,
                      double z_cm, double r_cm, double r_ball, double thresh, axb_stream_t s);
int axb_smooth_heaviside_mask(const axb_grid_t* g, double* H, uint8_t* mask, const double* phi, double blend_w,
                              double thresh, int greater_equal, axb_stream_t s);
int axb_add_bubble_flow(const axb_grid_t* g, double* u_z, double* u_r, const double* bubble_char_func,
                        const double* z1d, const double* r1d, double bubble_z_cm, double bubble_r_cm, double r0_bubble,
                        double U_0, double sin_omega_t, axb_stream_t s);
/* ---- the particle driver without host round trips: the loop scalars of
 *      particle_in_bubble_oscillatory_flow.py live in a device block `state` of >= 19 doubles
 *        [0] t  [1] dt  [2] max|w| (reduction target)  [3] penalisation sum (reduction target)  [4] U_z_cm_part
 *        [5] 0 (U_r)  [6] part_Z_cm  [7] F_total  [8] it  [9] sin(omega t)  [10] dt / cycle  [11] freqTimer
 *        [12] avg_Z_cm  [13] avg_time  [14] cycles  [15] wrap flag of this step  [16] diff  [17], [18] last cycle's
 *        avg_T / avg_part_trajectory point
 *      axb_particle_scalars phase 1 = :168-170 + :255-270 + :297-301 (cycle wrap, dt from max|w|, averages of the host
 *      scalars), phase 2 = compute_forces.py:4-17 + :323-355 (force, rigid-body update, t += dt); `trace`, if given, is
 *      a ring of trace_cap rows (t, dt, U_z_cm_part, part_Z_cm, F_total) written at the start-of-step values.
 *      axb_cycle_average3 = the three running averages of :297-299 in one pass, restarted (completed averages kept in
 *      last_i, may be NULL) when state[15] is set.  The _dev forms of the bubble flow and the sphere Heaviside read
 *      sin(omega t) (and U_0, if U_0_dev is given) / the sphere's z centre from device memory.
 *      Batched (axb_grid_t.batch > 1, one launch for every member of an ensemble): axb_kill_boundary_vorticity_sine_z/_r,
 *      axb_velocity_from_psi, axb_reduce_max_abs_sum, axb_add_bubble_flow_dev, axb_cycle_average3,
 *      axb_smooth_heaviside_sphere_dev, axb_penalise_update_vorticity, axb_advect_vorticity_particles,
 *      axb_diffusion_rk2_stage1/2(_dev) -- the row-marching stencil path only. ---------------------------------- */
int axb_particle_scalars(int phase, double* state, double* trace, int trace_cap, double dt_diff_limit, double cfl,
                         double eps, double cycle, double omega, double rho_lam, double part_vol, double part_mass,
                         double bubble_z_cm, double r0_bubble, axb_stream_t s);
/* the soft-sphere driver's scalars (soft_sphere_streaming.py:139-176, 201-203, 236-241, 262-264)
 * on a device block of >= 10 doubles:  [0] t  [1] dt  [2] max(|u_z|+|u_r|) (reduction target)  [3] freqTimer
 * [4] U_0 cos(omega t)  [5] 0  [6] Z_cm + amplitude sin(omega t)  [7] cycles  [8] wrap flag of the previous step  [9] it.
 * phase 1 (after the velocity maximum): dt with the cycle / tEnd clamps and the tether's velocity and position;
 * phase 2 (end of the step): t, the cycle timer and its wrap.  axb_cycle_average3 accepts fewer than three fields
 * (avg1 / avg2 NULL). */
int axb_soft_sphere_scalars(int phase, double* state, double dt_wave_limit, double cfl_dx, double eps, double dt_diff_limit,
                            double cycle, double t_end, double omega, double U_0, double Z_cm, double amplitude,
                            axb_stream_t s);
int axb_cycle_average3(const axb_grid_t* g, double* avg0, const double* x0, double* last0, double* avg1, const double* x1,
                       double* last1, double* avg2, const double* x2, double* last2, const double* a_dev,
                       const double* wrap_dev, axb_stream_t s);
int axb_add_bubble_flow_dev(const axb_grid_t* g, double* u_z, double* u_r, const double* bubble_char_func,
                            const double* z1d, const double* r1d, double bubble_z_cm, double bubble_r_cm,
                            double r0_bubble, double U_0, const double* U_0_dev, const double* sin_omega_t_dev,
                            axb_stream_t s);
/* The bubble does not move: axb_bubble_flow_geometry fills geom = -/+ ((Z - z_cm)^2 + (R - r_cm)^2)^1.5 (minus inside the
 * bubble, bubble_char_func >= 0.5) once, axb_add_bubble_flow_geom adds the flow of :273-294 from it with the bits of
 * axb_add_bubble_flow (the term that does not apply to a cell is an exact zero there) -- no pow() per step.  Batched;
 * geom (pitch ld_geom) is shared by the members. */
int axb_bubble_flow_geometry(const axb_grid_t* g, double* geom, const double* bubble_char_func, const double* z1d,
                             const double* r1d, double bubble_z_cm, double bubble_r_cm, axb_stream_t s);
int axb_add_bubble_flow_geom(const axb_grid_t* g, double* u_z, double* u_r, const double* geom, int64_t ld_geom,
                             const double* z1d, const double* r1d, double bubble_z_cm, double bubble_r_cm,
                             double r0_bubble, double U_0, const double* U_0_dev, const double* sin_omega_t_dev,
                             axb_stream_t s);
/* One launch for all members of an ensemble (batched): member m's scalar block is state + m scalar_stride
 * (scalar_stride >= 24) and additionally holds the constants that differ between members,
 *   [19] omega  [20] cycle time  [21] U_0  [22] nu  [23] diffusive dt limit,
 * its trace ring is trace + m 5 trace_cap. */
int axb_particle_scalars_batched(int phase, int batch, int scalar_stride, double* state, double* trace, int trace_cap,
                                 double cfl, double eps, double rho_lam, double part_vol, double part_mass,
                                 double bubble_z_cm, double r0_bubble, axb_stream_t s);
int axb_smooth_heaviside_sphere_dev(const axb_grid_t* g, double* H, double* phi_out, const double* z1d,
                                    const double* r1d, const double* z_cm_dev, double r_cm, double radius,
                                    double blend_w, axb_stream_t s);

/* ---- a18: elasto_kernels/solid_sigma.py:4-29 (all seven caller-visible outputs).
 *      chi != NULL additionally applies the driver's blend sigma *= chi
 *      (examples/SoftSphereStreaming/soft_sphere_streaming.py:236-238). --------------------- */
int axb_solid_sigma(const axb_grid_t* g, double* s11, double* s12, double* s22, double G, const double* eta1,
                    const double* eta2, double* eta1z, double* eta1r, double* eta2z, double* eta2r,
                    const double* chi, axb_stream_t s);
/* ---- a19: elasto_kernels/div_tau.py:4-34 as two launches (tau must be globally complete
 *      before its curl is taken). ------------------------------------------------------------ */
int axb_solid_tau(const axb_grid_t* g, double* tau_z, double* tau_r, const double* t11, const double* t12,
                  const double* t22, const double* r1d, axb_stream_t s);
int axb_solid_vorticity_update(const axb_grid_t* g, double* w, const double* tau_z, const double* tau_r,
                               double dt, const double* dt_dev, axb_stream_t s);
/* a18 + a19 in one pass for a driver that does not look at the intermediate arrays (soft_sphere_streaming.py:208-234):
 * w[1:-1, 1:-1] += dt curl(div(chi sigma(eta1, eta2))) with the bits of axb_solid_sigma -> axb_solid_tau ->
 * axb_solid_vorticity_update run on ZERO-INITIALISED gradient / stress / tau work arrays (the cells those calls leave
 * untouched then hold zeros for ever).  eta1, eta2, chi read once, w updated: ~50 instead of 152 B per cell.
 * exact_divisions = 1: those bits exactly (true FP64 divisions, issue bound); 0: divisions by 2 dx and r become
 * multiplications by reciprocals (<= 1e-13 relative to the exact form, several times faster). */
int axb_solid_stress_vorticity_update(const axb_grid_t* g, double* w, const double* eta1, const double* eta2,
                                      const double* chi, const double* r1d, double G, double dt, const double* dt_dev,
                                      int exact_divisions, axb_stream_t s);

/* ---- a20: core/src/extrapolate_using_least_squares.hpp:450-467 (order 1, 3x3 patch).
 *      cur/tgt int16 (n0, n1); eta_x / eta_y (n0, n1); gx[n1], gy[n0].  work holds
 *      axb_ls_workspace_bytes(n0, n1) bytes.  Sweeps run until no cell is added; the count of
 *      sweeps is written to *sweeps_host (this one call synchronises the stream, because
 *      the reference's do/while is data dependent). Bit-exact with the reference. ------------ */
int64_t axb_ls_workspace_bytes(int n0, int n1);
int axb_ls_extrapolate_order1(int n0, int n1, int16_t* cur, const int16_t* tgt, double* eta_x,
                              double* eta_y, const double* gx, const double* gy, void* work,
                              int64_t work_bytes, int max_sweeps, int* sweeps_host, axb_stream_t s);
/* core/src/extrapolate_using_least_squares.hpp:469-486 (extrapolate_using_least_squares_till_second_order):
 * same wavefront, quadratic basis [1, x, y, x^2, xy, y^2] on the 3x3 patch (6 x 6 normal equations). */
int axb_ls_extrapolate_order2(int n0, int n1, int16_t* cur, const int16_t* tgt, double* eta_x,
                              double* eta_y, const double* gx, const double* gy, void* work,
                              int64_t work_bytes, int max_sweeps, int* sweeps_host, axb_stream_t s);
/* the wrapper elasto_kernels/extrapolate_eta_using_least_squares_unb.py:7-30 fused: mirror by
 * index, flags from phi thresholds, extrapolate, write the physical half back. */
int axb_ls_extrapolate_eta(const axb_grid_t* g, const double* ball_phi, const uint8_t* inside_solid,
                           double* eta1, double* eta2, double extrap_zone, const double* gx,
                           const double* gy, void* work, int64_t work_bytes, int max_sweeps,
                           int* sweeps_host, axb_stream_t s);
/* The same without a host round trip (graph capturable): exactly `sweeps` (1..28) sweeps are enqueued on a fixed grid,
 * each reading its cell counts from the device; a sweep with nothing left to do costs an empty launch, a sweep that
 * finds no candidate ends the extrapolation like the reference's early return.  status_dev (device int[2], may be NULL)
 * <- {bit 0: pending list overflowed the workspace, bit 1: cells were still being added in the last sweep; number of
 * sweeps that added cells}.  Input and output reference maps may be different arrays (the output is written for every
 * cell), which lets a driver advect into a scratch pair and land the extrapolated maps back in the original one. */
int axb_ls_extrapolate_eta_device(const axb_grid_t* g, const double* ball_phi, const uint8_t* inside_solid,
                                  const double* eta1_in, const double* eta2_in, double* eta1_out, double* eta2_out,
                                  double extrap_zone, const double* gx, const double* gy, void* work,
                                  int64_t work_bytes, int sweeps, int32_t* status_dev, axb_stream_t s);
/* The device form in pieces, so that a driver can run independent work on a second stream while the (latency-bound,
 * nearly empty) sweep launches run: parts = 1 fill of the doubled work arrays and the first pending list, 2 the sweeps,
 * 4 the write-back into eta1_out / eta2_out; any sum of them in that order over one or more calls with the same
 * arguments and workspace is the call above (elasto_kernels/extrapolate_eta_using_least_squares_unb.py:7-30). */
int axb_ls_extrapolate_eta_device_parts(const axb_grid_t* g, const double* ball_phi, const uint8_t* inside_solid,
                                        const double* eta1_in, const double* eta2_in, double* eta1_out, double* eta2_out,
                                        double extrap_zone, const double* gx, const double* gy, void* work,
                                        int64_t work_bytes, int sweeps, int32_t* status_dev, int parts, axb_stream_t s);

/* ---- a21: core/src/particles_to_mesh.hpp:163-184 (periodic = 0) and the periodic twin
 *      particles_to_mesh_2D_mp4.  mesh is zeroed first, like the reference. ------------------ */
int axb_p2m_mp4_2d(int n0, int n1, const double* px, const double* py, const double* val, double* mesh,
                   double dx, double dy, int periodic, axb_stream_t s);
/* kernels/advect_particle.py:5-35 fused for lattice particles: push the mirrored lattice by
 * u*dt, remesh with MP4 on the doubled grid, return the physical half (w_out != w_in).
 * periodic = 0 runs the gather form: particles that move less than a cell are remeshed without atomics, every node
 * summed in the order of the reference's sequential particle loop (particles_to_mesh_2D.hpp:273-321) -- the result
 * is the reference's bit for bit; particles that move a cell or more are added by an atomic scatter pass.  The
 * _flagged form takes a caller-owned device int: the scatter pass returns at once when the gather pass saw no such
 * particle (the plain form always scans for them).  periodic = 1, or axb_set_p2m_atomic(1): the atomic scatter
 * (sums agree to rounding, order not fixed). ----- */
int axb_advect_vorticity_particles(const axb_grid_t* g, double* w_out, const double* w_in, const double* u_z,
                                   const double* u_r, const double* zl1d, const double* rl1d, double dt,
                                   const double* dt_dev, int periodic, axb_stream_t s);
int axb_advect_vorticity_particles_flagged(const axb_grid_t* g, double* w_out, const double* w_in, const double* u_z,
                                           const double* u_r, const double* zl1d, const double* rl1d, double dt,
                                           const double* dt_dev, int32_t* far_flag, axb_stream_t s);
int axb_set_p2m_atomic(int on);

/* ---- the rest of the particle <-> mesh family of the C++ core (core/src/instantiate.yml:1-27).
 *      kernel = one of AXB_PK_*: core/src/particle_kernels/LinearKernel.hpp, MP4.hpp, MP6.hpp,
 *      YangSmoothThreePointKernel.hpp.  periodic = 0: the "_unbounded" (edge-clipped) functions, 1: modulo wrap.
 *      Arrays are dense row-major, like the pybind11 bindings read them (core/src/*_bind.cpp).
 *   axb_m2p_2d   replaces mesh_to_particles_2D_{linear_kernel,mp4,mp6,yang_smooth_three_point_kernel} and the
 *                _unbounded_ twins (core/src/mesh_to_particles.hpp:61-253): two mesh fields (m0 x m1) sampled at
 *                p0 x p1 particles; sums in the reference's order, results bit-identical
 *   axb_p2m_2d   replaces particles_to_mesh_2D_* (core/src/particles_to_mesh.hpp:23-207): mesh zeroed, then the
 *                scatter (FP64 atomics: indices / weights exact, sum order free)
 *   axb_m2p_1d_mp4 / axb_p2m_1d_mp4   mesh_to_particles.hpp:25-37, particles_to_mesh.hpp:8-20 (periodic)
 *   axb_wrap_particles_2d   wrap_particles_around_2D_domain (mesh_to_particles.hpp:39-55): x wraps the first /
 *                last 10 entries of every row, y every entry of the first / last 10 rows; either pointer may be
 *                NULL; a 1-D array is n0 = 1 with px (wrap_particles_around_1D_domain) ------------------------- */
enum { AXB_PK_LINEAR = 0, AXB_PK_MP4 = 1, AXB_PK_MP6 = 2, AXB_PK_YANG = 3 };
int axb_m2p_2d(int kernel, int m0, int m1, const double* field_x, const double* field_y, int p0, int p1,
               const double* px, const double* py, double* out_x, double* out_y, double dx, double dy, int periodic,
               axb_stream_t s);
int axb_p2m_2d(int kernel, int m0, int m1, int p0, int p1, const double* px, const double* py, const double* val,
               double* mesh, double dx, double dy, int periodic, axb_stream_t s);
int axb_m2p_1d_mp4(int m, const double* field, int np, const double* pos, double* out, double dx, axb_stream_t s);
int axb_p2m_1d_mp4(int m, int np, const double* pos, const double* val, double* mesh, double dx, axb_stream_t s);
int axb_wrap_particles_2d(int n0, int n1, double* px, double* py, double x0, double x1, double y0, double y1,
                          axb_stream_t s);

/* ---- periodic-z transforms of G-FD (kernels/FastDiagonalisationStokesSolver.py:88-93, :137-156 with the periodic
 *      z operator): real FFT of every row in one pass over HBM, any even n = 2M whose half length has prime factors
 *      <= 64 only and fits shared memory (M <= 4266), e.g. the 4092 inner columns of the 1024 x 4096 periodic
 *      configuration.  Half-complex layout of a transformed row: [Re X_0 .. Re X_M | Im X_1 .. Im X_{M-1}].
 *      tables: n + 1 (re, im) pairs  [exp(-2 pi i k / M), k < M | exp(-2 pi i k / n), k <= M]  (host-rounded).
 *   axb_rfft_rows    dst = scale * rfft(src); columns n .. pad_to-1 of dst are zero-filled (pad_to <= ld_dst)
 *   axb_irfft_rows   dst = scale * (M * irfft(src)), i.e. pass scale = 1/M for the inverse of axb_rfft_rows(scale 1)
 *   axb_rfft_supported(n) -> 1 / 0 ------------------------------------------------------------------------------ */
int axb_rfft_rows(int rows, int n, const double* src, int64_t ld_src, double* dst, int64_t ld_dst, int pad_to,
                  const double* tables, double scale, axb_stream_t s);
int axb_irfft_rows(int rows, int n, const double* src, int64_t ld_src, double* dst, int64_t ld_dst,
                   const double* tables, double scale, axb_stream_t s);
int axb_rfft_supported(int n);

/* ---- static-PDE extrapolation (examples/PeriodicSoftSlab/bounded_static_PDE_extrapolation.py).
 *      Arrays are the reference's bounded arrays, dense (n0, n1) with n0, n1 >= 6; "interim" outputs are dense
 *      (n0-4, n1-4).  phi_b is the NEGATED level set (negative inside the solid), like `bounded_phi` (:68).
 *   axb_pde_extrap_setup   replaces _zones_setup (:142-147), _compute_normal_upwind (:149-171) and the set-up lines
 *                          :77-93 of extrapolate(): zone bit 0 = inside_solid, bit 1 = extrap_zone; the positive /
 *                          negative parts of the unit normal, denom = 3(|n_r| + |n_z|) + eps (interim arrays) and
 *                          grad_eta_n = inside * n . grad(eta) (bounded array, zero rim)
 *   axb_pde_extrap_jacobi  replaces _jacobi_iterate (:185-218): sweeps until ||update||_2 <= tol, terminated on the
 *                          device (the host reads a 16-byte state once per batch of 32 sweeps; the entry therefore
 *                          synchronises the stream, like the LS wavefront).  soln (bounded) is updated in place;
 *                          rhs (bounded) may be NULL = 0.  *sweeps_host = number of sweeps performed. ------------- */
int64_t axb_pde_extrap_workspace_bytes(int n0, int n1);
int axb_pde_extrap_setup(int n0, int n1, const double* phi_b, const double* eta_b, double dx, double offset, double band,
                         double eps, double* nr_pos, double* nr_neg, double* nz_pos, double* nz_neg, double* denom,
                         uint8_t* zone, double* grad_eta_n, axb_stream_t s);
int axb_pde_extrap_jacobi(int n0, int n1, double* soln, const double* rhs, const uint8_t* zone, const double* denom,
                          const double* nr_pos, const double* nr_neg, const double* nz_pos, const double* nz_neg,
                          double dx, double tol, int max_sweeps, void* work, int64_t work_bytes, int* sweeps_host,
                          axb_stream_t s);

/* ---- G-FD: kernels/FastDiagonalisationStokesSolver.py:130-156 (and the Potential /
 *      ImplicitEuler twins).  The plan holds device pointers to the caller-owned factors:
 *      Lr  = Vr^-1 (diag(r) folded in for the Stokes flavour)   (nr x nr, row-major)
 *      Rz  = Vz^-T                                               (nz x nz)
 *      Rzb = Vz^T                                                (nz x nz)
 *      Lrb = Vr                                                  (nr x nr)
 *      lam_r[nr], lam_z[nz]; spectral scaling 1 / (c0 + c1*(lam_z[n] + lam_r[m])).
 *      solve: psi = Lrb * (((Lr*rhs) * Rz) o scale) * Rzb, four FP64 tensor-core GEMMs with
 *      the scaling fused in the second one's epilogue.  work: 2*nr*nz doubles. ---------------- */
/*      Parity-split z-transform (n_leaves > 0): the cosine / sine eigenvectors of the uniform-grid
 *      z operator satisfy V[N-1-j, k] = (-1)^k V[j, k], so after folding a row into its even and
 *      odd parts (x[j] +- x[N-1-j], j < N/2) the even modes only see the even part and the odd
 *      modes the odd part: one N x N product becomes two N/2 x N/2 ones (half the flops).  For the
 *      Neumann (DCT-II) family the even branch is again a DCT-II and is split recursively.  The
 *      folded column layout is [E_L | O_L | ... | O_1]; leaf i multiplies columns
 *      [leaf_off[i], leaf_off[i] + leaf_n[i]) by leaf_fwd[i] (forward) / leaf_bwd[i] (backward),
 *      fold_len[] lists the segment lengths folded in order (N, N/2, ...), lam_z is in folded
 *      column order.  Rz / Rzb may be NULL in that case. */
#define AXB_FD_MAX_LEAVES 8
typedef struct axb_fd_plan {
  int32_t nr, nz;
  const double *Lr, *Rz, *Rzb, *Lrb;
  const double *lam_r, *lam_z;
  double c0, c1;
  double* work;
  int32_t n_leaves, n_folds;
  int32_t leaf_n[AXB_FD_MAX_LEAVES], leaf_off[AXB_FD_MAX_LEAVES], fold_len[AXB_FD_MAX_LEAVES];
  const double* leaf_fwd[AXB_FD_MAX_LEAVES];
  const double* leaf_bwd[AXB_FD_MAX_LEAVES];
  /* Optional direct r solve (r_tridiagonal != 0): the r operator A_r is tridiagonal, so after the
   * forward z transform every z-mode k is an independent tridiagonal system
   * (c0 I + c1 (A_r + lam_z[k] I)) x = rhs_k, solved by a batched Thomas sweep instead of the two
   * r-direction GEMMs (Lr / Lrb / lam_r are then unused and may be NULL).  r_sub / r_sup have nr-1
   * entries, r_diag nr; r_scale (nr entries or NULL) multiplies the right-hand side rows first (the
   * r o rhs of the Stokes flavour).  Same solution as the eigen-decomposition to rounding. */
  int32_t r_tridiagonal;
  const double *r_sub, *r_diag, *r_sup, *r_scale;
  /* Optional FFT z transforms (z_fft != 0; needs r_tridiagonal, the Neumann-z family and
   * nz = 2^p, 64 <= nz <= 16384): the Neumann eigenvectors are cos(pi k (2j+1) / (2 nz)), so the
   * forward / backward z transforms are a DCT-II / DCT-III of every row, done by axb_dct2_rows /
   * axb_dct3_rows in one pass over HBM each.  z_tables as described at axb_dct2_rows; lam_z in
   * natural mode order k = 0..nz-1.  Rz / Rzb / leaves are then unused. */
  int32_t z_fft;
  const double* z_tables;
  /* Optional reciprocal LU pivots of the nz tridiagonal systems (nr x nz, pitch nz) and per-row
   * coefficients (nr x 4), both filled once by axb_tridiag_factor_columns: the r solve is then two
   * division-free streaming sweeps (axb_tridiag_solve_factored).  NULL: the pivots are recomputed
   * inside every solve. */
  const double* r_inv_pivots;
  const double* r_row_coef;
  int32_t nz_spec;             /* z_fft == 2 (periodic real FFT, csrc/pfft.cu): pitch of the half-complex spectral rows,
                                  a multiple of 16 >= nz; lam_z, r_inv_pivots and work (2 * nr * nz_spec) follow it */
} axb_fd_plan_t;
/* Row-wise cosine transforms through a shared-memory FFT (n = 2^p, 64 <= n <= 16384):
 *   axb_dct2_rows: dst[m, k] = s_k * sum_j src[m, j] cos(pi k (2j+1) / (2n)), s_0 = scale0, s_k = scale
 *   axb_dct3_rows: dst[m, j] =       sum_k src[m, k] cos(pi k (2j+1) / (2n))
 * so dct3(dct2(x, 1/n, 2/n)) = x.  tables: 3n/2 + 2 complex numbers (re, im pairs of doubles),
 *   [exp(-2 pi i k / (n/2)), k < n/2 | exp(-2 pi i k / n), k <= n/2 | exp(-i pi k / (2n)), k <= n/2].
 * src and dst must not overlap. */
int axb_dct2_rows(int rows, int n, const double* src, int64_t ld_src, double* dst, int64_t ld_dst,
                  const double* tables, double scale0, double scale, axb_stream_t s);
int axb_dct3_rows(int rows, int n, const double* src, int64_t ld_src, double* dst, int64_t ld_dst,
                  const double* tables, axb_stream_t s);
/* in-place parity fold (inverse = 0: y[j] = x[j] + x[n-1-j], y[n/2+j] = x[j] - x[n-1-j]) or unfold
 * (inverse = 1: x[j] = y[j] + y[n/2+j], x[n-1-j] = y[j] - y[n/2+j]) of the first n columns of every
 * row of X (rows x >= n, pitch ld); n must be a multiple of 4. */
int axb_fd_fold(int rows, int n, double* X, int64_t ld, int inverse, axb_stream_t s);
/* same, out of place: the first n columns of src rows are folded into dst (src == dst allowed) */
int axb_fd_fold2(int rows, int n, const double* src, int64_t ld_src, double* dst, int64_t ld_dst, int inverse,
                 axb_stream_t s);
/* batched Thomas solve down the rows: for every column k < nz of X (nr x nz, pitch ld), in place,
 * (c0 I + c1 (tridiag(sub, diag, sup) + lam[k] I)) x = X[:, k] * scale.  scratch: nr*nz doubles. */
int axb_tridiag_solve_columns(int nr, int nz, double* X, int64_t ld, const double* sub, const double* diag,
                              const double* sup, const double* lam, const double* scale, double c0, double c1,
                              double* scratch, axb_stream_t s);
/* Factored form of the same solve.  axb_tridiag_factor_columns writes 1 / den[m, k] (the LU pivots of
 * column k's system) into inv_pivots (nr x nz, pitch nz) and the row coefficients
 * {c1 sub[m-1], scale[m] (1 if scale is NULL), c1 sup[m], 0} into row_coef (nr x 4);
 * axb_tridiag_solve_factored then solves in place with the right-hand sides in X (two TMA-streamed
 * sweeps).  Needs nz % 16 == 0, X / inv_pivots / row_coef 16-byte aligned, ld even. */
int axb_tridiag_factor_columns(int nr, int nz, const double* sub, const double* diag, const double* sup,
                               const double* lam, const double* scale, double c0, double c1, double* inv_pivots,
                               double* row_coef, axb_stream_t s);
int axb_tridiag_solve_factored(int nr, int nz, double* X, int64_t ld, const double* inv_pivots,
                               const double* row_coef, axb_stream_t s);
/* Partition (SPIKE) form of the r solve when the rows are split over ranks (multi-GPU r-slabs): every rank
 * solves its own diagonal block (axb_tridiag_solve_factored on its rows: g), the first / last rows of all
 * ranks' g are gathered into G (n_iface = 2 * ranks rows of nz), and
 *   X[m, k] = g[m, k] - V[m, k] xl[k] - W[m, k] xr[k],  xl = sum_i CL[i, k] G[i, k], xr = sum_i CR[i, k] G[i, k]
 * with the spikes V = T_p^-1 (a e_first), W = T_p^-1 (u e_last) and the rows CL / CR of the inverted reduced
 * interface system, all operator-only and prepared once (pyaxisymflow_b200/slab.py).  X, V, W: rows x nz. */
int axb_tridiag_partition_correct(int rows, int nz, double* X, int64_t ld, const double* V, const double* W,
                                  const double* G, const double* CL, const double* CR, int n_iface, axb_stream_t s);
/* The same with the decay of the spikes exploited: vcut_blk[b] / wcut_blk[b] (device, one entry per group of 256
 * columns, (nz / 2 + 128) / 128 groups) = the first row from which |V| is negligible / the first row |W| reaches in that
 * group; 32-row blocks in between are skipped (X is corrected in place).  NULL, NULL = correct everything. */
int axb_tridiag_partition_correct_banded(int rows, int nz, double* X, int64_t ld, const double* V, const double* W,
                                         const double* G, const double* CL, const double* CR, int n_iface,
                                         const int32_t* vcut_blk, const int32_t* wcut_blk, axb_stream_t s);
int axb_fd_solve(const axb_fd_plan_t* p, double* sol, int64_t ld_sol, const double* rhs, int64_t ld_rhs,
                 axb_stream_t s);
/* Plain row-major FP64 GEMM C = A*B (+ optional spectral scaling), the building block above:
 * A (M x K, lda), B (K x N, ldb), C (M x N, ldc).  scale_m / scale_n NULL = no scaling. */
int axb_dgemm(int M, int N, int K, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
              int64_t ldc, const double* scale_m, const double* scale_n, double c0, double c1,
              axb_stream_t s);

/* 0 (default): TMA + mbarrier pipeline when operands are 16-byte aligned; 1: force the LDGSTS
 * (cp.async) variant.  Both feed the same DMMA main loop; tests exercise both. */
int axb_dgemm_set_path(int force_ldgsts);

/* ---- transposes over peer memory: rank `me` stores, for every peer q < P, the block
 *      src[q * src_peer_stride + i * ld_src + c]  (i < rows, c < cols)  into
 *      peer_ptrs[q][dst_off + i * ld_dst + c],
 *      peer_ptrs[q] being rank q's destination buffer mapped into this process (NVLink peer memory).
 *      slab -> rows: src_peer_stride = (nr/P) * nzl, ld_src = nzl, dst_off = me * nzl, ld_dst = nz;
 *      rows -> slab: src_peer_stride = nzl, ld_src = nz, dst_off = me * (nr/P) * nzl, ld_dst = nzl.
 *      The caller synchronises the ranks (a device-side barrier) before the destinations are read. */
#define AXB_MAX_PEERS 16
int axb_peer_block_put(int P, int me, const uint64_t* peer_ptrs, int64_t dst_off, int64_t ld_dst, const double* src,
                       int64_t src_peer_stride, int64_t ld_src, int rows, int cols, axb_stream_t s);

/* ---- rank synchronisation through flags in peer-mapped memory: plain kernels (no collective library call), so a
 *      whole multi-GPU timestep can be captured in a CUDA graph.  Every rank owns a zero-initialised control block of
 *      AXB_CTL_WORDS 64-bit words mapped into all processes (ctl_ptrs[q] = rank q's block) and a LOCAL zero-initialised
 *      counter block of AXB_CTL_COUNTERS words holding the epochs (word 7 is raised if a wait gave up after ~2 s).
 *      axb_peer_sync: all-rank barrier.  axb_peer_allreduce_max: *value = max over the ranks (the CFL reduction of
 *      flow_past_sphere.py:150-153).  axb_row_halo_exchange: axb_row_halo_get fused with the neighbour handshake that
 *      must precede it (ctl_lower / ctl_upper = the neighbours' control blocks, NULL where there is none). ---------- */
#define AXB_CTL_WORDS 160
#define AXB_CTL_COUNTERS 8
int axb_peer_sync(int P, int me, const uint64_t* ctl_ptrs, uint64_t* counters, axb_stream_t s);
int axb_peer_allreduce_max(int P, int me, const uint64_t* ctl_ptrs, double* value, uint64_t* counters, axb_stream_t s);
int axb_row_halo_exchange(int nfields, const uint64_t* mine, const uint64_t* lower_peer, const uint64_t* upper_peer,
                          int64_t ld, int nz, int nrl, int halo, int width, uint64_t* ctl_mine, uint64_t* ctl_lower,
                          uint64_t* ctl_upper, uint64_t* counters, axb_stream_t s);

/* ---- z-slab plumbing (multi-GPU): pack / unpack `width` halo columns of a field ----------- */
int axb_halo_pack(const axb_grid_t* g, const double* f, double* buf_left, double* buf_right, int width,
                  axb_stream_t s);
int axb_halo_unpack(const axb_grid_t* g, double* f, const double* buf_left, const double* buf_right,
                    int width, double shift, axb_stream_t s);
/* the same exchange over peer memory: the first / last `width` owned columns of f are stored into the right
 * halo of left_peer_field (+ shift) / the left halo of right_peer_field (- shift), the neighbours' copies of
 * the same field mapped into this process (NULL = no neighbour on that side).  The caller synchronises the
 * ranks before the halos are read. */
int axb_halo_put(const axb_grid_t* g, const double* f, double* left_peer_field, double* right_peer_field, int width,
                 double shift, axb_stream_t s);
/* r-slab (row) decomposition: every rank stores (halo + nrl + halo) rows of pitch ld per field.  For up to 8 fields
 * at once (src[i] = this rank's block, lower_peer[i] / upper_peer[i] = the same field's block on rank-1 / rank+1
 * mapped into this process, 0 = no neighbour) the first `width` owned rows go into the lower neighbour's upper halo
 * rows and the last `width` owned rows into the upper neighbour's lower halo rows.  The pointer arrays are HOST
 * arrays.  The caller synchronises the ranks before the halos are read. */
int axb_row_halo_put(int nfields, const uint64_t* src, const uint64_t* lower_peer, const uint64_t* upper_peer, int64_t ld,
                     int nz, int nrl, int halo, int width, axb_stream_t s);
/* the pull form: this rank's halo rows are READ from the neighbours' owned edge rows.  The kernels of a step write
 * their own block's halo rows too (with values nobody uses), so with the pull form ONE rank barrier BEFORE the call
 * orders everything: the neighbour has finished producing its rows, and this rank's own stray halo writes are
 * stream-ordered before the pull (pyaxisymflow_b200/rowslab.py). */
int axb_row_halo_get(int nfields, const uint64_t* mine, const uint64_t* lower_peer, const uint64_t* upper_peer, int64_t ld,
                     int nz, int nrl, int halo, int width, axb_stream_t s);
/* local (nr x nz_local) slab  <->  P blocks of (nr/P x nz_local) for the all-to-all transpose */
int axb_slab_to_blocks(int nr, int nzl, int64_t ld, int P, const double* slab, double* blocks,
                       axb_stream_t s);
int axb_blocks_to_rows(int nrl, int nzl, int P, const double* blocks, double* rows, int64_t ld_rows,
                       axb_stream_t s);
int axb_rows_to_blocks(int nrl, int nzl, int P, const double* rows, int64_t ld_rows, double* blocks,
                       axb_stream_t s);
int axb_blocks_to_slab(int nr, int nzl, int64_t ld, int P, const double* blocks, double* slab,
                       axb_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* AXISYM_B200_H */
