// driver_glue.cu -- the per-step NumPy glue of the soft-sphere and particle drivers as kernels
// (SURVEY.md 8f rank 1): running averages, level-set pinning, Heaviside + inside mask, and the
// analytic bubble potential flow.  Same thread layout as stencils.cu, -fmad=false.
#include <math_constants.h>

#include <initializer_list>

#include "axb_common.cuh"

namespace {

__device__ __forceinline__ const double* rowp(const double* f, long long ld, int j) { return f + (long long)j * ld; }
__device__ __forceinline__ double* rowp(double* f, long long ld, int j) { return f + (long long)j * ld; }
inline bool vec_ok(const GridD& g, std::initializer_list<const void*> ptrs) {
  if ((g.ld & 1) || (g.bstride & 1)) return false;
  for (const void* p : ptrs)
    if (p && !axb_al16(p)) return false;
  return true;
}

// y += a * x          (soft_sphere_streaming.py:179-180, particle_in_bubble_oscillatory_flow.py:297-299)
__global__ void __launch_bounds__(TBX* TBY)
    k_axpy(GridD g, double* __restrict__ y, const double* __restrict__ x, double a, const double* __restrict__ a_dev,
           bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  if (a_dev) a = *a_dev;
  const double2 xv = ld_pair(rowp(x, g.ld, j), k, g.nz, vec);
  double2 yv = ld_pair(rowp(y, g.ld, j), k, g.nz, vec);
  yv.x = yv.x + xv.x * a;
  yv.y = yv.y + xv.y * a;
  st_pair(rowp(y, g.ld, j), k, g.ku0, g.ku1, vec, yv);
}

// phi_orig = -sqrt((eta1-zc)^2 + (eta2-rc)^2) + r_ball ; phi[phi > thresh] = phi_orig   (soft_sphere_streaming.py:191-193)
__global__ void __launch_bounds__(TBX* TBY)
    k_pin(GridD g, double* __restrict__ phi, double* __restrict__ phi_orig, const double* __restrict__ e1,
          const double* __restrict__ e2, double zc, double rc, double r_ball, double thresh, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  const double2 a = ld_pair(rowp(e1, g.ld, j), k, g.nz, vec), b = ld_pair(rowp(e2, g.ld, j), k, g.nz, vec);
  double2 p = ld_pair(rowp(phi, g.ld, j), k, g.nz, vec);
  double2 o;
  o.x = -sqrt((a.x - zc) * (a.x - zc) + (b.x - rc) * (b.x - rc)) + r_ball;
  o.y = -sqrt((a.y - zc) * (a.y - zc) + (b.y - rc) * (b.y - rc)) + r_ball;
  if (phi_orig) st_pair(rowp(phi_orig, g.ld, j), k, g.ku0, g.ku1, vec, o);
  if (p.x > thresh) p.x = o.x;
  if (p.y > thresh) p.y = o.y;
  st_pair(rowp(phi, g.ld, j), k, g.ku0, g.ku1, vec, p);
}

// same as stencils.cu: outside the band the reference's blend term is multiplied by an exact zero, so the division and
// the sine are only evaluated for band cells (same bits)
__device__ __forceinline__ double heav1(double phi, double w) {
  if (fabs(phi) < w) return 0.0 + 0.5 * (1 + phi / w + sin(CUDART_PI * phi / w) / CUDART_PI);
  return (phi >= w) ? 1.0 : 0.0;
}
// H = smooth_Heaviside(phi) and mask = (H > thresh) as a dense (nr, nz) uint8   (soft_sphere_streaming.py:205-206)
// Four columns per thread (two 128-bit accesses in flight per field, one 32-bit store of the mask) when the rows allow it.
__global__ void __launch_bounds__(128)
    k_heav_mask(GridD g, double* __restrict__ H, unsigned char* __restrict__ mask, const double* __restrict__ phi,
                double w, double thresh, int ge, bool vec4) {
  const int k = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int j = blockIdx.y;
  if (k >= g.nz) return;
  const double* p = phi + (long long)j * g.ld + k;
  double* o = H + (long long)j * g.ld + k;
  unsigned char* m = mask + (long long)j * g.nz + k;
  if (vec4 && k + 3 < g.nz) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    const double h0 = heav1(a.x, w), h1 = heav1(a.y, w), h2 = heav1(b.x, w), h3 = heav1(b.y, w);
    *reinterpret_cast<double2*>(o) = make_double2(h0, h1);
    *reinterpret_cast<double2*>(o + 2) = make_double2(h2, h3);
    uchar4 q;
    q.x = ge ? (h0 >= thresh) : (h0 > thresh);
    q.y = ge ? (h1 >= thresh) : (h1 > thresh);
    q.z = ge ? (h2 >= thresh) : (h2 > thresh);
    q.w = ge ? (h3 >= thresh) : (h3 > thresh);
    *reinterpret_cast<uchar4*>(m) = q;
    return;
  }
  for (int c = 0; c < 4 && k + c < g.nz; ++c) {
    const double h = heav1(p[c], w);
    o[c] = h;
    m[c] = ge ? (h >= thresh) : (h > thresh);
  }
}

// bubble breathing mode + exterior potential flow added to (u_z, u_r)   (particle_in_bubble_oscillatory_flow.py:273-294)
__global__ void __launch_bounds__(TBX* TBY)
    k_bubble(GridD g, double* __restrict__ u_z, double* __restrict__ u_r, const double* __restrict__ chi_b,
             const double* __restrict__ z1d, const double* __restrict__ r1d, double bz, double br, double r0, double U0,
             double s, const double* __restrict__ s_dev, const double* __restrict__ U0_dev, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  {
    const long long fo = member_field(g);
    u_z += fo; u_r += fo; chi_b += fo;
  }
  if (s_dev) s = s_dev[member_scalar(g)];
  if (U0_dev) U0 = U0_dev[member_scalar(g)];
  const double2 c = ld_pair(rowp(chi_b, g.ld, j), k, g.nz, vec);
  double2 uz = ld_pair(rowp(u_z, g.ld, j), k, g.nz, vec), ur = ld_pair(rowp(u_r, g.ld, j), k, g.nz, vec);
  const double dr = r1d[j] - br;
  const double dz[2] = {z1d[k] - bz, z1d[(k + 1 < g.nz) ? k + 1 : k] - bz};
  const double in[2] = {(c.x >= 0.5) ? 1.0 : 0.0, (c.y >= 0.5) ? 1.0 : 0.0};
  double vz[2] = {uz.x, uz.y}, vr[2] = {ur.x, ur.y};
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double zb = U0 * dz[i] * s / r0, rb = U0 * dr * s / r0;
    vz[i] = vz[i] + in[i] * zb;
    vr[i] = vr[i] + in[i] * rb;
    const double d15 = pow(dz[i] * dz[i] + dr * dr, 1.5);
    vz[i] = vz[i] + (1.0 - in[i]) * U0 * dz[i] * s * (r0 * r0) / d15;
    vr[i] = vr[i] + (1.0 - in[i]) * U0 * dr * s * (r0 * r0) / d15;
  }
  st_pair(rowp(u_z, g.ld, j), k, g.ku0, g.ku1, vec, make_double2(vz[0], vz[1]));
  st_pair(rowp(u_r, g.ld, j), k, g.ku0, g.ku1, vec, make_double2(vr[0], vr[1]));
}

// The same update from a precomputed geometry field (the bubble does not move): geom = -(d^2)^1.5 inside the bubble
// (bubble_char_func >= 0.5), +(d^2)^1.5 outside.  A cell inside takes only the breathing term, a cell outside only
// the potential-flow term -- the other term of particle_in_bubble_oscillatory_flow.py:273-294 is an exact zero -- so
// the result has the bits of k_bubble without its pow() and with two instead of four divisions per cell.  `geom` is
// shared by the members of an ensemble (not moved by the member index).
__global__ void __launch_bounds__(TBX* TBY)
    k_bubble_geometry(GridD g, double* __restrict__ geom, const double* __restrict__ chi_b, const double* __restrict__ z1d,
                      const double* __restrict__ r1d, double bz, double br) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.nr || k >= g.nz) return;
  const double dz = z1d[k] - bz, dr = r1d[j] - br;
  const double d15 = pow(dz * dz + dr * dr, 1.5);
  geom[(long long)j * g.ld + k] = (chi_b[(long long)j * g.ld + k] >= 0.5) ? -d15 : d15;
}
__global__ void __launch_bounds__(TBX* TBY)
    k_bubble_pre(GridD g, double* __restrict__ u_z, double* __restrict__ u_r, const double* __restrict__ geom, long long ld_geom,
                 const double* __restrict__ z1d, const double* __restrict__ r1d, double bz, double br, double r0, double U0,
                 const double* __restrict__ s_dev, const double* __restrict__ U0_dev, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  {
    const long long fo = member_field(g);
    u_z += fo; u_r += fo;
  }
  const double s = s_dev[member_scalar(g)];
  if (U0_dev) U0 = U0_dev[member_scalar(g)];
  const double2 gm = ld_pair(rowp(geom, ld_geom, j), k, g.nz, vec && !(ld_geom & 1));
  double2 uz = ld_pair(rowp(u_z, g.ld, j), k, g.nz, vec), ur = ld_pair(rowp(u_r, g.ld, j), k, g.nz, vec);
  const double dr = r1d[j] - br;
  const double dz[2] = {z1d[k] - bz, z1d[(k + 1 < g.nz) ? k + 1 : k] - bz};
  const double gg[2] = {gm.x, gm.y};
  double vz[2] = {uz.x, uz.y}, vr[2] = {ur.x, ur.y};
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (gg[i] < 0.0) {
      vz[i] = vz[i] + U0 * dz[i] * s / r0;
      vr[i] = vr[i] + U0 * dr * s / r0;
    } else {
      vz[i] = vz[i] + U0 * dz[i] * s * (r0 * r0) / gg[i];
      vr[i] = vr[i] + U0 * dr * s * (r0 * r0) / gg[i];
    }
  }
  st_pair(rowp(u_z, g.ld, j), k, g.ku0, g.ku1, vec, make_double2(vz[0], vz[1]));
  st_pair(rowp(u_r, g.ld, j), k, g.ku0, g.ku1, vec, make_double2(vr[0], vr[1]));
}

// running averages of three fields in one pass, restarted when the cycle timer wrapped (device flag):
//   avg_i = (wrap ? 0 : avg_i) + a x_i;  on a wrap the completed averages are kept in last_i (may be null)
struct Avg3 {
  double* avg[3];
  const double* x[3];
  double* last[3];
};
__global__ void __launch_bounds__(TBX* TBY)
    k_cycle_avg3(GridD g, Avg3 f, const double* __restrict__ a_dev, const double* __restrict__ wrap_dev, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  const double a = a_dev[member_scalar(g)];
  const bool wrap = wrap_dev[member_scalar(g)] != 0.0;
  const long long fo = member_field(g);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (!f.avg[i]) continue;                       // fewer than three fields
    const double2 xv = ld_pair(rowp(f.x[i] + fo, g.ld, j), k, g.nz, vec);
    double2 yv = ld_pair(rowp(f.avg[i] + fo, g.ld, j), k, g.nz, vec);
    if (wrap) {
      if (f.last[i]) st_pair(rowp(f.last[i] + fo, g.ld, j), k, g.ku0, g.ku1, vec, yv);
      yv = make_double2(0.0, 0.0);
    }
    yv.x = yv.x + xv.x * a;
    yv.y = yv.y + xv.y * a;
    st_pair(rowp(f.avg[i] + fo, g.ld, j), k, g.ku0, g.ku1, vec, yv);
  }
}

// The host decisions of one particle step (particle_in_bubble_oscillatory_flow.py:168-170, 255-270, 297-301,
// 323-355) on a device-resident scalar block, so that the step needs no host round trip:
//   st[0] t, [1] dt, [2] max|w| (reduction target), [3] penalisation sum (reduction target), [4] U_z_cm_part,
//   [5] 0 (U_r), [6] part_Z_cm, [7] F_total, [8] it, [9] sin(omega t), [10] dt / cycle, [11] freqTimer, [12] avg_Z_cm,
//   [13] avg_time, [14] cycles, [15] wrap flag of this step, [16] diff, [17] last avg_T, [18] last avg trajectory point
//   [19] omega  [20] cycle time  [21] U_0  [22] nu  [23] diffusive dt limit: the constants that differ between the
//   members of an ensemble; read from the block instead of the by-value parameters when `member_consts` is set.
// One block per ensemble member: member m's scalars are st + m * stride, its trace ring trace + m * 5 * trace_cap.
struct ParticleParams {
  double dt_diff, cfl, eps, cycle, omega, rho_lam, part_vol, part_mass, bubble_z, r0;
};
__global__ void k_particle_scalars(int phase, double* st, double* trace, int trace_cap, ParticleParams p, int stride,
                                   int member_consts) {
  if (threadIdx.x != 0) return;
  st += (long long)blockIdx.x * stride;
  if (trace) trace += (long long)blockIdx.x * 5 * trace_cap;
  if (member_consts) {
    p.omega = st[19];
    p.cycle = st[20];
    p.dt_diff = st[23];
  }
  if (phase == 1) {
    double wrap = 0.0;
    if (st[11] >= p.cycle) {
      wrap = 1.0;
      st[11] = 0.0;
      st[14] = st[14] + 1.0;
      st[17] = st[13] / p.cycle;
      st[18] = (st[12] / p.cycle - p.bubble_z) / p.r0;
      st[12] = 0.0;
      st[13] = 0.0;
    }
    st[15] = wrap;
    const double dt = fmin(fmin(p.dt_diff, p.cfl / (st[2] + p.eps)), 0.01 * p.cycle);
    st[1] = dt;
    st[9] = sin(p.omega * st[0]);
    st[10] = dt / p.cycle;
    st[12] = st[12] + st[6] * dt;
    st[13] = st[13] + st[0] * dt;
    st[11] = st[11] + dt;
    st[3] = 0.0;
  } else {
    const double dt = st[1];
    const double F_pen = p.rho_lam * st[3];
    const double F_un = (st[16] * p.part_vol) / dt;
    const double F = F_pen + F_un;
    st[7] = F;
    if (trace && trace_cap > 0) {
      double* row = trace + 5 * ((long long)st[8] % trace_cap);
      row[0] = st[0]; row[1] = dt; row[2] = st[4]; row[3] = st[6]; row[4] = F;
    }
    const double U_old = st[4];
    st[4] = st[4] + 0.5 * dt * (st[16] / dt + (F / p.part_mass));
    st[16] = dt * F / p.part_mass;
    st[6] = st[6] + (U_old * dt + (0.5 * dt * dt * F / p.part_mass));
    st[0] = st[0] + dt;
    st[8] = st[8] + 1.0;
    st[2] = 0.0;
  }
}


// The host decisions of one soft-sphere step (soft_sphere_streaming.py:139-176, 201-203, 236-241, 262-264) on a
// device block:  st[0] t  [1] dt  [2] max(|u_z| + |u_r|) (reduction target)  [3] freqTimer  [4] U_0 cos(omega t)
//   [5] 0 (U_r)  [6] Z_cm + e r_ball sin(omega t)  [7] cycles  [8] wrap flag: the previous step completed a cycle (the
//   running averages restart)  [9] it
struct SoftParams {
  double dt_wave, cfl_dx, eps, dt_diff, cycle, t_end, omega, U0, Z_cm, amp;
};
__global__ void k_soft_scalars(int phase, double* st, SoftParams p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (phase == 1) {
    double dt = fmin(fmin(p.dt_wave, p.cfl_dx / (st[2] + p.eps)), p.dt_diff);
    if (st[3] + dt > p.cycle) dt = p.cycle - st[3];
    if (st[0] + dt > p.t_end) dt = p.t_end - st[0];
    st[1] = dt;
    st[4] = p.U0 * cos(p.omega * st[0]);
    st[6] = p.Z_cm + p.amp * sin(p.omega * st[0]);
  } else {
    st[0] = st[0] + st[1];
    st[3] = st[3] + st[1];
    double wrap = 0.0;
    if (st[3] >= p.cycle) {
      st[3] = 0.0;
      st[7] = st[7] + 1.0;
      wrap = 1.0;
    }
    st[8] = wrap;
    st[2] = 0.0;
    st[9] = st[9] + 1.0;
  }
}

}  // namespace

extern "C" {

int axb_cycle_average3(const axb_grid_t* g, double* avg0, const double* x0, double* last0, double* avg1, const double* x1,
                       double* last1, double* avg2, const double* x2, double* last2, const double* a_dev,
                       const double* wrap_dev, axb_stream_t s) {
  if (!avg0 || !x0 || (!avg1 != !x1) || (!avg2 != !x2) || !a_dev || !wrap_dev) return AXB_EINVAL;
  int rc = axb_check_grid_batched(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  Avg3 f;
  f.avg[0] = avg0; f.avg[1] = avg1; f.avg[2] = avg2;
  f.x[0] = x0; f.x[1] = x1; f.x[2] = x2;
  f.last[0] = last0; f.last[1] = last1; f.last[2] = last2;
  k_cycle_avg3<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, f, a_dev, wrap_dev,
                                                   vec_ok(d, {avg0, x0, last0, avg1, x1, last1, avg2, x2, last2}));
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_soft_sphere_scalars(int phase, double* state, double dt_wave_limit, double cfl_dx, double eps, double dt_diff_limit,
                            double cycle, double t_end, double omega, double U_0, double Z_cm, double amplitude,
                            axb_stream_t s) {
  if (!state || (phase != 1 && phase != 2)) return AXB_EINVAL;
  const SoftParams p = {dt_wave_limit, cfl_dx, eps, dt_diff_limit, cycle, t_end, omega, U_0, Z_cm, amplitude};
  k_soft_scalars<<<1, 32, 0, s>>>(phase, state, p);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_particle_scalars(int phase, double* state, double* trace, int trace_cap, double dt_diff_limit, double cfl,
                         double eps, double cycle, double omega, double rho_lam, double part_vol, double part_mass,
                         double bubble_z_cm, double r0_bubble, axb_stream_t s) {
  if (!state || (phase != 1 && phase != 2) || trace_cap < 0) return AXB_EINVAL;
  const ParticleParams p = {dt_diff_limit, cfl, eps, cycle, omega, rho_lam, part_vol, part_mass, bubble_z_cm, r0_bubble};
  k_particle_scalars<<<1, 32, 0, s>>>(phase, state, trace, trace_cap, p, 0, 0);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_particle_scalars_batched(int phase, int batch, int scalar_stride, double* state, double* trace, int trace_cap,
                                 double cfl, double eps, double rho_lam, double part_vol, double part_mass,
                                 double bubble_z_cm, double r0_bubble, axb_stream_t s) {
  if (!state || (phase != 1 && phase != 2) || trace_cap < 0 || batch < 1 || scalar_stride < 24) return AXB_EINVAL;
  const ParticleParams p = {0.0, cfl, eps, 0.0, 0.0, rho_lam, part_vol, part_mass, bubble_z_cm, r0_bubble};
  k_particle_scalars<<<batch, 32, 0, s>>>(phase, state, trace, trace_cap, p, scalar_stride, 1);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_axpy(const axb_grid_t* g, double* y, const double* x, double a, const double* a_dev, axb_stream_t s) {
  if (!y || !x) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  k_axpy<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, y, x, a, a_dev, vec_ok(d, {y, x}));
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_pin_level_set(const axb_grid_t* g, double* phi, double* phi_orig, const double* eta1, const double* eta2,
                      double z_cm, double r_cm, double r_ball, double thresh, axb_stream_t s) {
  if (!phi || !eta1 || !eta2) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  k_pin<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, phi, phi_orig, eta1, eta2, z_cm, r_cm, r_ball, thresh,
                                           vec_ok(d, {phi, phi_orig, eta1, eta2}));
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_smooth_heaviside_mask(const axb_grid_t* g, double* H, uint8_t* mask, const double* phi, double blend_w,
                              double thresh, int greater_equal, axb_stream_t s) {
  if (!H || !mask || !phi) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  if (g->ku0 != 0 || g->ku1 != g->nz) return AXB_ENOSUP;
  const GridD d = to_dev(g);
  const bool vec4 = vec_ok(d, {H, phi}) && (d.nz % 4 == 0) && ((uintptr_t)mask % 4 == 0);
  k_heav_mask<<<dim3((d.nz + 511) / 512, d.nr), 128, 0, s>>>(d, H, mask, phi, blend_w, thresh, greater_equal, vec4);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

static int bubble_flow(const axb_grid_t* g, double* u_z, double* u_r, const double* bubble_char_func, const double* z1d,
                       const double* r1d, double bubble_z_cm, double bubble_r_cm, double r0_bubble, double U_0,
                       double sin_omega_t, const double* sin_dev, const double* U0_dev, axb_stream_t s) {
  if (!u_z || !u_r || !bubble_char_func || !z1d || !r1d) return AXB_EINVAL;
  int rc = sin_dev ? axb_check_grid_batched(g) : axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  k_bubble<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, u_z, u_r, bubble_char_func, z1d, r1d, bubble_z_cm, bubble_r_cm,
                                              r0_bubble, U_0, sin_omega_t, sin_dev, U0_dev,
                                              vec_ok(d, {u_z, u_r, bubble_char_func}));
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_add_bubble_flow(const axb_grid_t* g, double* u_z, double* u_r, const double* bubble_char_func,
                        const double* z1d, const double* r1d, double bubble_z_cm, double bubble_r_cm, double r0_bubble,
                        double U_0, double sin_omega_t, axb_stream_t s) {
  return bubble_flow(g, u_z, u_r, bubble_char_func, z1d, r1d, bubble_z_cm, bubble_r_cm, r0_bubble, U_0, sin_omega_t,
                     nullptr, nullptr, s);
}
int axb_bubble_flow_geometry(const axb_grid_t* g, double* geom, const double* bubble_char_func, const double* z1d,
                             const double* r1d, double bubble_z_cm, double bubble_r_cm, axb_stream_t s) {
  if (!geom || !bubble_char_func || !z1d || !r1d) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  k_bubble_geometry<<<dim3((d.nz + 31) / 32, (d.nr + 7) / 8), dim3(32, 8), 0, s>>>(d, geom, bubble_char_func, z1d, r1d,
                                                                                 bubble_z_cm, bubble_r_cm);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_add_bubble_flow_geom(const axb_grid_t* g, double* u_z, double* u_r, const double* geom, int64_t ld_geom,
                             const double* z1d, const double* r1d, double bubble_z_cm, double bubble_r_cm,
                             double r0_bubble, double U_0, const double* U_0_dev, const double* sin_omega_t_dev,
                             axb_stream_t s) {
  if (!u_z || !u_r || !geom || !z1d || !r1d || !sin_omega_t_dev) return AXB_EINVAL;
  int rc = axb_check_grid_batched(g);
  if (rc) return rc;
  if (ld_geom < g->nz) return AXB_EINVAL;
  const GridD d = to_dev(g);
  k_bubble_pre<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, u_z, u_r, geom, ld_geom, z1d, r1d, bubble_z_cm, bubble_r_cm,
                                                  r0_bubble, U_0, sin_omega_t_dev, U_0_dev, vec_ok(d, {u_z, u_r, geom}));
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_add_bubble_flow_dev(const axb_grid_t* g, double* u_z, double* u_r, const double* bubble_char_func,
                            const double* z1d, const double* r1d, double bubble_z_cm, double bubble_r_cm,
                            double r0_bubble, double U_0, const double* U_0_dev, const double* sin_omega_t_dev,
                            axb_stream_t s) {
  if (!sin_omega_t_dev) return AXB_EINVAL;
  return bubble_flow(g, u_z, u_r, bubble_char_func, z1d, r1d, bubble_z_cm, bubble_r_cm, r0_bubble, U_0, 0.0,
                     sin_omega_t_dev, U_0_dev, s);
}

}  // extern "C"
