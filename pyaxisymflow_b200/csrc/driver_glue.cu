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
  if (g.ld & 1) return false;
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
__global__ void k_heav_mask(GridD g, double* __restrict__ H, unsigned char* __restrict__ mask,
                            const double* __restrict__ phi, double w, double thresh, int ge) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (k >= g.nz) return;
  const double h = heav1(phi[(long long)j * g.ld + k], w);
  H[(long long)j * g.ld + k] = h;
  mask[(long long)j * g.nz + k] = ge ? (h >= thresh) : (h > thresh);
}

// bubble breathing mode + exterior potential flow added to (u_z, u_r)   (particle_in_bubble_oscillatory_flow.py:273-294)
__global__ void __launch_bounds__(TBX* TBY)
    k_bubble(GridD g, double* __restrict__ u_z, double* __restrict__ u_r, const double* __restrict__ chi_b,
             const double* __restrict__ z1d, const double* __restrict__ r1d, double bz, double br, double r0, double U0,
             double s, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
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

}  // namespace

extern "C" {

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
  k_heav_mask<<<dim3((d.nz + 127) / 128, d.nr), 128, 0, s>>>(d, H, mask, phi, blend_w, thresh, greater_equal);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_add_bubble_flow(const axb_grid_t* g, double* u_z, double* u_r, const double* bubble_char_func,
                        const double* z1d, const double* r1d, double bubble_z_cm, double bubble_r_cm, double r0_bubble,
                        double U_0, double sin_omega_t, axb_stream_t s) {
  if (!u_z || !u_r || !bubble_char_func || !z1d || !r1d) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  k_bubble<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, u_z, u_r, bubble_char_func, z1d, r1d, bubble_z_cm, bubble_r_cm,
                                              r0_bubble, U_0, sin_omega_t, vec_ok(d, {u_z, u_r, bubble_char_func}));
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

}  // extern "C"
