// elasto.cu -- hyperelastic solid stress passes (fusion groups G-SOL-1/2/3).
//   G-SOL-1  elasto_kernels/solid_sigma.py:18-29      gradients of (eta1, eta2) -> sigma
//   G-SOL-2  elasto_kernels/div_tau.py:8-28           tau = div(sigma) with the 1/r terms
//   G-SOL-3  elasto_kernels/div_tau.py:30-34          w[int] += dt * curl(tau)
// Same thread layout as stencils.cu (two z columns per thread, 64 x 8 tiles), -fmad=false.
#include <stdlib.h>

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

// One reference map: writes eta_z (cols 1..nz-2, all rows) and eta_r (interior + row 0);
// cells the reference leaves untouched keep their old contents, which the stress then reads.
__device__ __forceinline__ void grad_eta(const GridD& g, const double* __restrict__ eta, double* __restrict__ ez,
                                         double* __restrict__ er, int j, int k, bool vec, bool eager, double2& gz,
                                         double2& gr) {
  const int nz = g.nz;
  const double h = 2 * g.dx;
  const double* e = rowp(eta, g.ld, j);
  // Cells the reference does not recompute (first / last global column for d/dz; last row, and those columns of the
  // rows >= 1, for d/dr) keep their stale contents, which the stress then reads.  Only pairs that contain such a
  // cell load the old gradients; everywhere else they are overwritten (saves 32 B/pt of reads).
  bool stale = eager || (j == g.nr - 1);
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int kk = k + cc, kg = kk + g.kz0;
    if (kk >= g.ku0 && kk < g.ku1 && (kg < 1 || kg > g.nzg - 2)) stale = true;
  }
  gz = make_double2(0.0, 0.0);
  gr = make_double2(0.0, 0.0);
  if (stale) {
    gz = ld_pair(rowp(ez, g.ld, j), k, nz, vec);
    gr = ld_pair(rowp(er, g.ld, j), k, nz, vec);
  }
  const double2 c = ld_pair(e, k, nz, vec);
  double2 up = c, dn = c;
  if (j >= 1 && j < g.nr - 1) {
    up = ld_pair(rowp(eta, g.ld, j + 1), k, nz, vec);
    dn = ld_pair(rowp(eta, g.ld, j - 1), k, nz, vec);
  } else if (j == 0) {
    up = ld_pair(rowp(eta, g.ld, 1), k, nz, vec);   // eta[1]
    dn = ld_pair(rowp(eta, g.ld, 2), k, nz, vec);   // eta[2]
  }
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int kk = k + cc;
    if (kk < g.ku0 || kk >= g.ku1) continue;
    const int kg = kk + g.kz0;
    const bool zin = (kg >= 1 && kg <= g.nzg - 2);
    const double ev = cc ? c.y : c.x, uv = cc ? up.y : up.x, dv = cc ? dn.y : dn.x;
    if (zin) {
      const double dz = (e[kk + 1] - e[kk - 1]) / h;
      if (cc) gz.y = dz; else gz.x = dz;
    }
    if (j == 0) {
      const double d = (-dv + 4 * uv - 3 * ev) / h;
      if (cc) gr.y = d; else gr.x = d;
    } else if (j < g.nr - 1 && zin) {
      const double d = (uv - dv) / h;
      if (cc) gr.y = d; else gr.x = d;
    }
  }
  st_pair(rowp(ez, g.ld, j), k, g.ku0, g.ku1, vec, gz);
  st_pair(rowp(er, g.ld, j), k, g.ku0, g.ku1, vec, gr);
}

__global__ void __launch_bounds__(TBX* TBY)
    k_solid_sigma(GridD g, double* __restrict__ s11, double* __restrict__ s12, double* __restrict__ s22, double G,
                  const double* __restrict__ eta1, const double* __restrict__ eta2, double* __restrict__ e1z,
                  double* __restrict__ e1r, double* __restrict__ e2z, double* __restrict__ e2r,
                  const double* __restrict__ chi, bool vec, bool eager) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  double2 z1, r1, z2, r2;
  grad_eta(g, eta1, e1z, e1r, j, k, vec, eager, z1, r1);
  grad_eta(g, eta2, e2z, e2r, j, k, vec, eager, z2, r2);
  double2 a12, a11, a22;
  a12.x = -G * (z1.x * r1.x + z2.x * r2.x);
  a12.y = -G * (z1.y * r1.y + z2.y * r2.y);
  a11.x = (0.5 * G) * (r1.x * r1.x + r2.x * r2.x - z1.x * z1.x - z2.x * z2.x);
  a11.y = (0.5 * G) * (r1.y * r1.y + r2.y * r2.y - z1.y * z1.y - z2.y * z2.y);
  a22.x = -a11.x; a22.y = -a11.y;
  if (chi) {
    const double2 c = ld_pair(rowp(chi, g.ld, j), k, g.nz, vec);
    a11.x = c.x * a11.x; a11.y = c.y * a11.y;
    a12.x = c.x * a12.x; a12.y = c.y * a12.y;
    a22.x = c.x * a22.x; a22.y = c.y * a22.y;
  }
  st_pair(rowp(s11, g.ld, j), k, g.ku0, g.ku1, vec, a11);
  st_pair(rowp(s12, g.ld, j), k, g.ku0, g.ku1, vec, a12);
  st_pair(rowp(s22, g.ld, j), k, g.ku0, g.ku1, vec, a22);
}

// tau_z = d_z t11 + d_r t12 + t12/r ; tau_r = d_z t12 + d_r t22 + t22/r   (rows 0..nr-2, cols 1..nz-2)
__global__ void __launch_bounds__(TBX* TBY)
    k_solid_tau(GridD g, double* __restrict__ tau_z, double* __restrict__ tau_r, const double* __restrict__ t11,
                const double* __restrict__ t12, const double* __restrict__ t22, const double* __restrict__ r1d,
                bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr - 1 || k >= g.ku1 || k + 1 < g.ku0) return;
  const int nz = g.nz;
  const double h = 2 * g.dx;
  const double r = r1d[j];
  const double* a = rowp(t11, g.ld, j);
  const double* b = rowp(t12, g.ld, j);
  const double2 b0 = ld_pair(b, k, nz, vec);
  const double2 c0 = ld_pair(rowp(t22, g.ld, j), k, nz, vec);
  // r-direction samples: interior uses rows j+1 / j-1, row 0 uses rows 1 and 2 (one-sided)
  const int ja = (j == 0) ? 1 : j + 1, jb = (j == 0) ? 2 : j - 1;
  const double2 bu = ld_pair(rowp(t12, g.ld, ja), k, nz, vec), bd = ld_pair(rowp(t12, g.ld, jb), k, nz, vec);
  const double2 cu = ld_pair(rowp(t22, g.ld, ja), k, nz, vec), cd = ld_pair(rowp(t22, g.ld, jb), k, nz, vec);
  double* oz = rowp(tau_z, g.ld, j);
  double* orr = rowp(tau_r, g.ld, j);
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int kk = k + cc;
    if (kk < g.ku0 || kk >= g.ku1) continue;
    const int kg = kk + g.kz0;
    if (kg < 1 || kg > g.nzg - 2) continue;
    const double b_c = cc ? b0.y : b0.x, c_c = cc ? c0.y : c0.x;
    const double b_u = cc ? bu.y : bu.x, b_d = cc ? bd.y : bd.x;
    const double c_u = cc ? cu.y : cu.x, c_d = cc ? cd.y : cd.x;
    double tz, tr;
    if (j > 0) {
      tz = (a[kk + 1] - a[kk - 1] + b_u - b_d) / h + b_c / r;
      tr = (b[kk + 1] - b[kk - 1] + c_u - c_d) / h + c_c / r;
    } else {
      // b_u = row 1, b_d = row 2
      tz = (a[kk + 1] - a[kk - 1] - b_d + 4 * b_u - 3 * b_c) / h + b_c / r;
      tr = (b[kk + 1] - b[kk - 1] - c_d + 4 * c_u - 3 * c_c) / h + c_c / r;
    }
    oz[kk] = tz;
    orr[kk] = tr;
  }
}

__global__ void __launch_bounds__(TBX* TBY)
    k_solid_curl(GridD g, double* __restrict__ w, const double* __restrict__ tau_z, const double* __restrict__ tau_r,
                 double dt, const double* __restrict__ dt_dev, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j < 1 || j >= g.nr - 1 || k >= g.ku1 || k + 1 < g.ku0) return;
  if (dt_dev) dt = *dt_dev;
  const int nz = g.nz;
  const double h = 2 * g.dx;
  const double2 zu = ld_pair(rowp(tau_z, g.ld, j + 1), k, nz, vec), zd = ld_pair(rowp(tau_z, g.ld, j - 1), k, nz, vec);
  const double* tr = rowp(tau_r, g.ld, j);
  double* wr = rowp(w, g.ld, j);
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int kk = k + cc;
    if (kk < g.ku0 || kk >= g.ku1) continue;
    const int kg = kk + g.kz0;
    if (kg < 1 || kg > g.nzg - 2) continue;
    const double u = cc ? zu.y : zu.x, d = cc ? zd.y : zd.x;
    wr[kk] = wr[kk] + dt * (tr[kk + 1] - tr[kk - 1] - u + d) / h;
  }
}

}  // namespace

extern "C" {

int axb_solid_sigma(const axb_grid_t* g, double* s11, double* s12, double* s22, double G, const double* eta1,
                    const double* eta2, double* eta1z, double* eta1r, double* eta2z, double* eta2r,
                    const double* chi, axb_stream_t s) {
  if (!s11 || !s12 || !s22 || !eta1 || !eta2 || !eta1z || !eta1r || !eta2z || !eta2r) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  if (d.nr < 3 || d.nzg < 3) return AXB_EINVAL;
  const bool vec = vec_ok(d, {s11, s12, s22, eta1, eta2, eta1z, eta1r, eta2z, eta2r, chi});
  static const bool eager = getenv("AXB_SIGMA_EAGER") != nullptr;  // load the old gradients everywhere (A/B switch)
  k_solid_sigma<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, s11, s12, s22, G, eta1, eta2, eta1z, eta1r, eta2z, eta2r, chi,
                                                     vec, eager);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_solid_tau(const axb_grid_t* g, double* tau_z, double* tau_r, const double* t11, const double* t12,
                  const double* t22, const double* r1d, axb_stream_t s) {
  if (!tau_z || !tau_r || !t11 || !t12 || !t22 || !r1d) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  if (d.nr < 3 || d.nzg < 3) return AXB_EINVAL;
  const bool vec = vec_ok(d, {tau_z, tau_r, t11, t12, t22});
  k_solid_tau<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, tau_z, tau_r, t11, t12, t22, r1d, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_solid_vorticity_update(const axb_grid_t* g, double* w, const double* tau_z, const double* tau_r,
                               double dt, const double* dt_dev, axb_stream_t s) {
  if (!w || !tau_z || !tau_r) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  const bool vec = vec_ok(d, {w, tau_z, tau_r});
  k_solid_curl<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, w, tau_z, tau_r, dt, dt_dev, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

}  // extern "C"
