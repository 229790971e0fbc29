// elasto.cu -- hyperelastic solid stress passes (fusion groups G-SOL-1/2/3).
//   G-SOL-1  elasto_kernels/solid_sigma.py:18-29      gradients of (eta1, eta2) -> sigma
//   G-SOL-2  elasto_kernels/div_tau.py:8-28           tau = div(sigma) with the 1/r terms
//   G-SOL-3  elasto_kernels/div_tau.py:30-34          w[int] += dt * curl(tau)
// Same thread layout as stencils.cu (two z columns per thread, 64 x 8 tiles), -fmad=false.
#include <stdlib.h>

#include <initializer_list>

#include "axb_common.cuh"

extern int g_axb_legacy_stencils;  // capi.cu: 1 = 2-D tiled kernels only
extern int g_axb_solid_march;      // capi.cu: 1 = row-marching G-SOL-1/2 kernels (axb_set_solid_march)

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

__device__ __forceinline__ void sigma_pair(const GridD& g, double* __restrict__ s11, double* __restrict__ s12,
                                           double* __restrict__ s22, double G, const double* __restrict__ eta1,
                                           const double* __restrict__ eta2, double* __restrict__ e1z,
                                           double* __restrict__ e1r, double* __restrict__ e2z, double* __restrict__ e2r,
                                           const double* __restrict__ chi, bool vec, bool eager, int j, int k) {
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

__global__ void __launch_bounds__(TBX* TBY)
    k_solid_sigma(GridD g, double* __restrict__ s11, double* __restrict__ s12, double* __restrict__ s22, double G,
                  const double* __restrict__ eta1, const double* __restrict__ eta2, double* __restrict__ e1z,
                  double* __restrict__ e1r, double* __restrict__ e2z, double* __restrict__ e2r,
                  const double* __restrict__ chi, bool vec, bool eager) {
  sigma_pair(g, s11, s12, s22, G, eta1, eta2, e1z, e1r, e2z, e2r, chi, vec, eager, blockIdx.y * TBY + threadIdx.y,
             2 * (blockIdx.x * TBX + threadIdx.x));
}

// tau_z = d_z t11 + d_r t12 + t12/r ; tau_r = d_z t12 + d_r t22 + t22/r   (rows 0..nr-2, cols 1..nz-2)
__device__ __forceinline__ void tau_pair(const GridD& g, double* __restrict__ tau_z, double* __restrict__ tau_r,
                                         const double* __restrict__ t11, const double* __restrict__ t12,
                                         const double* __restrict__ t22, const double* __restrict__ r1d, bool vec, int j,
                                         int k) {
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
    k_solid_tau(GridD g, double* __restrict__ tau_z, double* __restrict__ tau_r, const double* __restrict__ t11,
                const double* __restrict__ t12, const double* __restrict__ t22, const double* __restrict__ r1d,
                bool vec) {
  tau_pair(g, tau_z, tau_r, t11, t12, t22, r1d, vec, blockIdx.y * TBY + threadIdx.y,
           2 * (blockIdx.x * TBX + threadIdx.x));
}

// -------------------------------------------------------------------------------------
// Row-marching forms of G-SOL-1 / G-SOL-2 (same idea as stencils_march.cu): a block owns 256 adjacent columns (two
// per thread, 128-bit accesses) and marches over RBM rows with the r-neighbourhood in a rolling register window;
// z neighbours come from warp shuffles.  Interior blocks only -- every cell there is recomputed by the reference,
// so no old gradient is read; edge blocks run the general per-pair code above on the same grid.  Same expressions
// and true divisions as the tiled kernels: a cell gets the same bits whichever kernel computes it.
// OPT-IN (axb_set_solid_march(1)): measured inside the 2048 x 8192 soft-sphere step the pair interior + edge launch
// is slower than the tiled kernels (3.24 vs 3.05 ms per step) -- the edge kernel runs after the interior one on the
// same stream instead of beside it (stencils_march.cu forks it to a side stream); kept for that follow-up.
// -------------------------------------------------------------------------------------
constexpr int MTM = 128, RBM = 16;

__device__ __forceinline__ bool march_interior(const GridD& g, int mbx, int j0, bool vec) {
  const int kb0 = 2 * mbx * MTM, kb1 = kb0 + 2 * MTM;
  return vec && (j0 >= 1) && (j0 + RBM + 1 <= g.nr) && (kb0 >= g.ku0) && (kb1 <= g.ku1) && (kb0 + g.kz0 >= 1) &&
         (kb1 - 1 + g.kz0 <= g.nzg - 2) && (kb0 >= 1) && (kb1 < g.nz);
}
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
// left = f[k-1], right = f[k+2] of the thread's pair c = (f[k], f[k+1]); `row` points at f[k]
__device__ __forceinline__ void z_nb(const double2 c, const double* __restrict__ row, int lane, double& left, double& right) {
  left = __shfl_up_sync(0xffffffffu, c.y, 1);
  right = __shfl_down_sync(0xffffffffu, c.x, 1);
  if (lane == 0) left = row[-1];
  if (lane == 31) right = row[2];
}

template <bool CHI>
__global__ void __launch_bounds__(MTM)
    km_solid_sigma(GridD g, double* __restrict__ s11, double* __restrict__ s12, double* __restrict__ s22, double G,
                   const double* __restrict__ eta1, const double* __restrict__ eta2, double* __restrict__ e1z,
                   double* __restrict__ e1r, double* __restrict__ e2z, double* __restrict__ e2r,
                   const double* __restrict__ chi) {
  const int j0 = blockIdx.y * RBM;
  if (!march_interior(g, blockIdx.x, j0, true)) return;
  const int k = 2 * (blockIdx.x * MTM + threadIdx.x), lane = threadIdx.x & 31;
  const long long ld = g.ld;
  const double h = 2 * g.dx;
  long long o = (long long)(j0 - 1) * ld + k;
  double2 m1 = ld2(eta1 + o), m2 = ld2(eta2 + o);
  o += ld;
  double2 c1 = ld2(eta1 + o), c2 = ld2(eta2 + o);
  for (int j = j0; j < j0 + RBM; ++j, o += ld) {
    const double2 n1 = ld2(eta1 + o + ld), n2 = ld2(eta2 + o + ld);
    double2 cx = make_double2(1.0, 1.0);
    if (CHI) cx = ld2(chi + o);
    double l1, r1n, l2, r2n;
    z_nb(c1, eta1 + o, lane, l1, r1n);
    z_nb(c2, eta2 + o, lane, l2, r2n);
    double2 z1, r1, z2, r2;
    z1.x = (c1.y - l1) / h;   z1.y = (r1n - c1.x) / h;
    z2.x = (c2.y - l2) / h;   z2.y = (r2n - c2.x) / h;
    r1.x = (n1.x - m1.x) / h; r1.y = (n1.y - m1.y) / h;
    r2.x = (n2.x - m2.x) / h; r2.y = (n2.y - m2.y) / h;
    double2 a12, a11, a22;
    a12.x = -G * (z1.x * r1.x + z2.x * r2.x);
    a12.y = -G * (z1.y * r1.y + z2.y * r2.y);
    a11.x = (0.5 * G) * (r1.x * r1.x + r2.x * r2.x - z1.x * z1.x - z2.x * z2.x);
    a11.y = (0.5 * G) * (r1.y * r1.y + r2.y * r2.y - z1.y * z1.y - z2.y * z2.y);
    a22.x = -a11.x; a22.y = -a11.y;
    if (CHI) {
      a11.x = cx.x * a11.x; a11.y = cx.y * a11.y;
      a12.x = cx.x * a12.x; a12.y = cx.y * a12.y;
      a22.x = cx.x * a22.x; a22.y = cx.y * a22.y;
    }
    st2(e1z + o, z1); st2(e1r + o, r1); st2(e2z + o, z2); st2(e2r + o, r2);
    st2(s11 + o, a11); st2(s12 + o, a12); st2(s22 + o, a22);
    m1 = c1; c1 = n1;
    m2 = c2; c2 = n2;
  }
}

__global__ void __launch_bounds__(MTM)
    km_solid_sigma_edge(GridD g, double* __restrict__ s11, double* __restrict__ s12, double* __restrict__ s22, double G,
                        const double* __restrict__ eta1, const double* __restrict__ eta2, double* __restrict__ e1z,
                        double* __restrict__ e1r, double* __restrict__ e2z, double* __restrict__ e2r,
                        const double* __restrict__ chi, bool eager) {
  const int j0 = blockIdx.y * RBM;
  if (march_interior(g, blockIdx.x, j0, true)) return;
  const int k = 2 * (blockIdx.x * MTM + threadIdx.x);
  if (k >= g.nz) return;
  for (int j = j0; j < min(j0 + RBM, g.nr); ++j)
    sigma_pair(g, s11, s12, s22, G, eta1, eta2, e1z, e1r, e2z, e2r, chi, true, eager, j, k);
}

__global__ void __launch_bounds__(MTM)
    km_solid_tau(GridD g, double* __restrict__ tau_z, double* __restrict__ tau_r, const double* __restrict__ t11,
                 const double* __restrict__ t12, const double* __restrict__ t22, const double* __restrict__ r1d) {
  const int j0 = blockIdx.y * RBM;
  if (!march_interior(g, blockIdx.x, j0, true)) return;
  const int k = 2 * (blockIdx.x * MTM + threadIdx.x), lane = threadIdx.x & 31;
  const long long ld = g.ld;
  const double h = 2 * g.dx;
  long long o = (long long)(j0 - 1) * ld + k;
  double2 bm = ld2(t12 + o), cm = ld2(t22 + o);
  o += ld;
  double2 bc = ld2(t12 + o), cc = ld2(t22 + o);
  for (int j = j0; j < j0 + RBM; ++j, o += ld) {
    const double2 bn = ld2(t12 + o + ld), cn = ld2(t22 + o + ld);
    const double2 ac = ld2(t11 + o);
    const double r = r1d[j];
    double al, ar, bl, br;
    z_nb(ac, t11 + o, lane, al, ar);
    z_nb(bc, t12 + o, lane, bl, br);
    double2 tz, tr;
    tz.x = (ac.y - al + bn.x - bm.x) / h + bc.x / r;
    tz.y = (ar - ac.x + bn.y - bm.y) / h + bc.y / r;
    tr.x = (bc.y - bl + cn.x - cm.x) / h + cc.x / r;
    tr.y = (br - bc.x + cn.y - cm.y) / h + cc.y / r;
    st2(tau_z + o, tz);
    st2(tau_r + o, tr);
    bm = bc; bc = bn;
    cm = cc; cc = cn;
  }
}

__global__ void __launch_bounds__(MTM)
    km_solid_tau_edge(GridD g, double* __restrict__ tau_z, double* __restrict__ tau_r, const double* __restrict__ t11,
                      const double* __restrict__ t12, const double* __restrict__ t22, const double* __restrict__ r1d) {
  const int j0 = blockIdx.y * RBM;
  if (march_interior(g, blockIdx.x, j0, true)) return;
  const int k = 2 * (blockIdx.x * MTM + threadIdx.x);
  if (k >= g.nz) return;
  for (int j = j0; j < min(j0 + RBM, g.nr); ++j) tau_pair(g, tau_z, tau_r, t11, t12, t22, r1d, true, j, k);
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


// -------------------------------------------------------------------------------------
// G-SOL fused (stepper form): eta -> sigma -> tau -> w += dt curl(tau) in ONE pass.
// The three reference calls move 152 B per cell through seven intermediate arrays the driver never looks at
// (soft_sphere_streaming.py:208-234); here a block stages the (16 + 6) x (64 + 6) neighbourhood of the two reference
// maps in shared memory, evaluates sigma on the (16 + 4) x (64 + 4) ring, tau on (16 + 2) x (64 + 2) and updates its
// 16 x 64 cells of w: eta1, eta2, chi read once (x 1.3-1.5 for the halos), w read and written -- about 50 B per cell.
// Every stage evaluates the expressions of sigma_pair / tau_pair / k_solid_curl (true divisions, -fmad=false), and
// cells the reference leaves untouched (first / last column, last row: "stale" above) read as zero, which is what
// they hold in a driver whose work arrays start zeroed -- so w gets the bits of the three separate calls.
// -------------------------------------------------------------------------------------
constexpr int FT_R = 16, FT_C = 64;
constexpr int FE_R = FT_R + 6, FE_C = FT_C + 6, FS_R = FT_R + 4, FS_C = FT_C + 4, FQ_R = FT_R + 2, FQ_C = FT_C + 2;

// EXACT: true divisions (the bits of the three calls; FP64-issue bound: ~9 divisions per cell, 0.76 ms at 2048 x 8192);
// otherwise divisions by 2 dx and r are multiplications by reciprocals computed once (<= 2 ulp per operation, like the
// row-marching stencils; the parity bar is 1e-10).
template <bool EXACT>
__device__ __forceinline__ double qdiv(double a, double b, double inv_b) { return EXACT ? a / b : a * inv_b; }

template <bool CHI, bool EXACT>
__global__ void __launch_bounds__(256, 3)
    k_solid_fused(GridD g, double* __restrict__ w, const double* __restrict__ eta1, const double* __restrict__ eta2,
                  const double* __restrict__ chi, const double* __restrict__ r1d, double G, double dt,
                  const double* __restrict__ dt_dev) {
  __shared__ double e1[FE_R * FE_C], e2[FE_R * FE_C], s11[FS_R * FS_C], s12[FS_R * FS_C];
  __shared__ double inv_r[FQ_R];
  double* tz = e1;     // tau re-uses the staging arrays once sigma is done
  double* tr = e2;
  const int J0 = blockIdx.y * FT_R, K0 = blockIdx.x * FT_C;
  const int nr = g.nr, nz = g.nz;
  const double h = 2 * g.dx, ih = 1.0 / h;
  if (dt_dev) dt = *dt_dev;
  if (!EXACT && threadIdx.x < FQ_R) {
    const int j = J0 - 1 + (int)threadIdx.x;
    inv_r[threadIdx.x] = (j >= 0 && j < nr) ? 1.0 / r1d[j] : 0.0;
  }
  // stage 1: reference maps on rows J0-3 .. J0+FT_R+2, columns K0-3 .. K0+FT_C+2 (zero outside the domain).  All global
  // loads of the block -- the two maps, chi for the sigma ring, the block's own w -- are issued before the first use.
  constexpr int N1 = (FE_R * FE_C + 255) / 256, N2 = (FS_R * FS_C + 255) / 256, N4 = (FT_R * FT_C) / 256;
  double va[N1], vb[N1], cx[N2], w0[N4];
#pragma unroll
  for (int it = 0; it < N1; ++it) {
    const int i = threadIdx.x + it * 256;
    const int rr = i / FE_C, cc = i - rr * FE_C;
    const int j = J0 - 3 + rr, k = K0 - 3 + cc;
    va[it] = 0.0;
    vb[it] = 0.0;
    if (i < FE_R * FE_C && j >= 0 && j < nr && k >= 0 && k < nz) {
      va[it] = eta1[(long long)j * g.ld + k];
      vb[it] = eta2[(long long)j * g.ld + k];
    }
  }
#pragma unroll
  for (int it = 0; it < N2; ++it) {
    const int i = threadIdx.x + it * 256;
    const int rr = i / FS_C, cc = i - rr * FS_C;
    const int j = J0 - 2 + rr, k = K0 - 2 + cc;
    cx[it] = 1.0;
    if (CHI && i < FS_R * FS_C && j >= 0 && j < nr && k >= 0 && k < nz) cx[it] = chi[(long long)j * g.ld + k];
  }
#pragma unroll
  for (int it = 0; it < N4; ++it) {
    const int i = threadIdx.x + it * 256;
    const int rr = i / FT_C, cc = i - rr * FT_C;
    const int j = J0 + rr, k = K0 + cc;
    w0[it] = (j >= 1 && j < nr - 1 && k >= 1 && k <= nz - 2) ? w[(long long)j * g.ld + k] : 0.0;
  }
#pragma unroll
  for (int it = 0; it < N1; ++it) {
    const int i = threadIdx.x + it * 256;
    if (i < FE_R * FE_C) {
      e1[i] = va[it];
      e2[i] = vb[it];
    }
  }
  __syncthreads();
  // stage 2: sigma on rows J0-2 .. , columns K0-2 ..
#pragma unroll
  for (int it = 0; it < N2; ++it) {
    const int i = threadIdx.x + it * 256;
    if (i >= FS_R * FS_C) break;
    const int rr = i / FS_C, cc = i - rr * FS_C;
    const int j = J0 - 2 + rr, k = K0 - 2 + cc;
    double a11 = 0.0, a12 = 0.0;
    if (j >= 0 && j < nr && k >= 0 && k < nz) {
      const int c = (rr + 1) * FE_C + (cc + 1);                 // this cell in the eta staging
      const bool zin = (k >= 1 && k <= nz - 2);
      double z1 = 0.0, r1 = 0.0, z2 = 0.0, r2 = 0.0;
      if (zin) {
        z1 = qdiv<EXACT>(e1[c + 1] - e1[c - 1], h, ih);
        z2 = qdiv<EXACT>(e2[c + 1] - e2[c - 1], h, ih);
      }
      if (j == 0) {
        // rows 1 and 2 (one-sided); row 0 of the domain is staging row 3 of the first row block
        const int c1 = c + FE_C, c2 = c + 2 * FE_C;
        r1 = qdiv<EXACT>(-e1[c2] + 4 * e1[c1] - 3 * e1[c], h, ih);
        r2 = qdiv<EXACT>(-e2[c2] + 4 * e2[c1] - 3 * e2[c], h, ih);
      } else if (j < nr - 1 && zin) {
        r1 = qdiv<EXACT>(e1[c + FE_C] - e1[c - FE_C], h, ih);
        r2 = qdiv<EXACT>(e2[c + FE_C] - e2[c - FE_C], h, ih);
      }
      a12 = -G * (z1 * r1 + z2 * r2);
      a11 = (0.5 * G) * (r1 * r1 + r2 * r2 - z1 * z1 - z2 * z2);
      if (CHI) {
        a11 = cx[it] * a11;
        a12 = cx[it] * a12;
      }
    }
    s11[i] = a11;
    s12[i] = a12;
  }
  __syncthreads();
  // stage 3: tau on rows J0-1 .. , columns K0-1 ..  (rows 0 .. nr-2, columns 1 .. nz-2; zero elsewhere); t22 = -t11
  for (int i = threadIdx.x; i < FQ_R * FQ_C; i += 256) {
    const int rr = i / FQ_C, cc = i - rr * FQ_C;
    const int j = J0 - 1 + rr, k = K0 - 1 + cc;
    double vz = 0.0, vr = 0.0;
    if (j >= 0 && j < nr - 1 && k >= 1 && k <= nz - 2) {
      const int c = (rr + 1) * FS_C + (cc + 1);                 // this cell in the sigma arrays
      const double r = EXACT ? r1d[j] : 0.0, ir = EXACT ? 0.0 : inv_r[rr];
      const double b_c = s12[c], c_c = -s11[c];
      if (j > 0) {
        const double b_u = s12[c + FS_C], b_d = s12[c - FS_C], c_u = -s11[c + FS_C], c_d = -s11[c - FS_C];
        vz = qdiv<EXACT>(s11[c + 1] - s11[c - 1] + b_u - b_d, h, ih) + qdiv<EXACT>(b_c, r, ir);
        vr = qdiv<EXACT>(s12[c + 1] - s12[c - 1] + c_u - c_d, h, ih) + qdiv<EXACT>(c_c, r, ir);
      } else {
        const double b_u = s12[c + FS_C], b_d = s12[c + 2 * FS_C], c_u = -s11[c + FS_C], c_d = -s11[c + 2 * FS_C];
        vz = qdiv<EXACT>(s11[c + 1] - s11[c - 1] - b_d + 4 * b_u - 3 * b_c, h, ih) + qdiv<EXACT>(b_c, r, ir);
        vr = qdiv<EXACT>(s12[c + 1] - s12[c - 1] - c_d + 4 * c_u - 3 * c_c, h, ih) + qdiv<EXACT>(c_c, r, ir);
      }
    }
    tz[i] = vz;
    tr[i] = vr;
  }
  __syncthreads();
  // stage 4: the block's own cells
#pragma unroll
  for (int it = 0; it < N4; ++it) {
    const int i = threadIdx.x + it * 256;
    const int rr = i / FT_C, cc = i - rr * FT_C;
    const int j = J0 + rr, k = K0 + cc;
    if (j < 1 || j >= nr - 1 || k < 1 || k > nz - 2) continue;
    const int c = (rr + 1) * FQ_C + (cc + 1);
    w[(long long)j * g.ld + k] = w0[it] + qdiv<EXACT>(dt * (tr[c + 1] - tr[c - 1] - tz[c + FQ_C] + tz[c - FQ_C]), h, ih);
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
  const bool march = vec && g_axb_solid_march && !g_axb_legacy_stencils && d.nr >= RBM + 2 && d.nz >= 2 * MTM + 2;
  if (march) {
    const dim3 mg((d.nz + 2 * MTM - 1) / (2 * MTM), (d.nr + RBM - 1) / RBM);
    if (chi)
      km_solid_sigma<true><<<mg, MTM, 0, s>>>(d, s11, s12, s22, G, eta1, eta2, eta1z, eta1r, eta2z, eta2r, chi);
    else
      km_solid_sigma<false><<<mg, MTM, 0, s>>>(d, s11, s12, s22, G, eta1, eta2, eta1z, eta1r, eta2z, eta2r, nullptr);
    AXB_LAUNCHED();
    km_solid_sigma_edge<<<mg, MTM, 0, s>>>(d, s11, s12, s22, G, eta1, eta2, eta1z, eta1r, eta2z, eta2r, chi, eager);
  } else {
    k_solid_sigma<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, s11, s12, s22, G, eta1, eta2, eta1z, eta1r, eta2z, eta2r,
                                                       chi, vec, eager);
  }
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
  const bool march = vec && g_axb_solid_march && !g_axb_legacy_stencils && d.nr >= RBM + 2 && d.nz >= 2 * MTM + 2;
  if (march) {
    const dim3 mg((d.nz + 2 * MTM - 1) / (2 * MTM), (d.nr + RBM - 1) / RBM);
    km_solid_tau<<<mg, MTM, 0, s>>>(d, tau_z, tau_r, t11, t12, t22, r1d);
    AXB_LAUNCHED();
    km_solid_tau_edge<<<mg, MTM, 0, s>>>(d, tau_z, tau_r, t11, t12, t22, r1d);
  } else {
    k_solid_tau<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, tau_z, tau_r, t11, t12, t22, r1d, vec);
  }
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_solid_stress_vorticity_update(const axb_grid_t* g, double* w, const double* eta1, const double* eta2,
                                      const double* chi, const double* r1d, double G, double dt, const double* dt_dev,
                                      int exact_divisions, axb_stream_t s) {
  if (!w || !eta1 || !eta2 || !r1d || w == eta1 || w == eta2) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  if (g->ku0 != 0 || g->ku1 != g->nz || g->nz_global != g->nz || g->nr < 3 || g->nz < 3) return AXB_ENOSUP;
  const GridD d = to_dev(g);
  const dim3 grd((d.nz + FT_C - 1) / FT_C, (d.nr + FT_R - 1) / FT_R);
#define FUSED(C, E) k_solid_fused<C, E><<<grd, 256, 0, s>>>(d, w, eta1, eta2, chi, r1d, G, dt, dt_dev)
  if (chi) { if (exact_divisions) FUSED(true, true); else FUSED(true, false); }
  else { if (exact_divisions) FUSED(false, true); else FUSED(false, false); }
#undef FUSED
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
