// stencils.cu -- fused (r, z) stencil passes of the rigid-flow timestep (sm_100a).
//
// Fusion groups of SURVEY.md section 8a: G-BND, G-VEL, G-PEN, G-DIF, G-HEAV, the periodic
// ghost copy, the diagnostic reductions and the device-side scalar glue.  All kernels are
// HBM-bound FP64 streaming passes: every thread owns two adjacent z-columns (128-bit
// loads/stores on the contiguous axis), a block covers 64 columns x 8 rows so that the
// r-neighbours of a row are served by L1, and reductions go warp shuffle -> shared -> one
// atomic per block.  Compiled with -fmad=false so the arithmetic is the plain IEEE sequence
// the reference's NumPy/numba expressions perform.
#include <math_constants.h>

#include <initializer_list>

#include "axb_common.cuh"
#include "axb_march.cuh"

namespace {

// row pointer helper
__device__ __forceinline__ const double* rowp(const double* f, long long ld, int j) { return f + (long long)j * ld; }
__device__ __forceinline__ double* rowp(double* f, long long ld, int j) { return f + (long long)j * ld; }

// -------------------------------------------------------------------------------------
// G-BND  kernels/kill_boundary_vorticity_sine.py:9-14 (z) and :23-27 (r)
// -------------------------------------------------------------------------------------
__global__ void k_kill_z(GridD g, double* __restrict__ w, const double* __restrict__ z1d, int width) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.nr) return;
  w += member_field(g);
  double* row = rowp(w, g.ld, j);
  const double den = (double)(width - 1);
  // left end: global columns [0, width)
  {
    const int ksrc = (width - 1) - g.kz0;  // local index of global column width-1
    if (ksrc >= g.ku0 && ksrc < g.ku1) {
      const double src = row[ksrc];
      for (int kg = 0; kg < width; ++kg) {
        const int k = kg - g.kz0;
        if (k < g.ku0 || k >= g.ku1) continue;
        const double ramp = sin(CUDART_PI * (z1d[k] - 0.5 * g.dx) / 2 / den / g.dx);
        row[k] = ramp * src;
      }
    }
  }
  // right end: global columns [nzg - width, nzg)
  {
    const int ksrc = (g.nzg - width) - g.kz0;
    if (ksrc >= g.ku0 && ksrc < g.ku1) {
      const double src = row[ksrc];
      for (int kg = g.nzg - width; kg < g.nzg; ++kg) {
        const int k = kg - g.kz0;
        if (k < g.ku0 || k >= g.ku1) continue;
        const double ramp = sin(CUDART_PI * (1 - z1d[k] - 0.5 * g.dx) / 2 / den / g.dx);
        row[k] = ramp * src;
      }
    }
  }
}

__global__ void k_kill_r(GridD g, double* __restrict__ w, const double* __restrict__ r1d, int width, int parts) {
  const int k = g.ku0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.ku1) return;
  w += member_field(g);
  if (parts & 1) {
    const double den = (double)(width - 1);
    const double src = rowp(w, g.ld, g.nr - width)[k];
    for (int j = g.nr - width; j < g.nr; ++j) {
      const double ramp = sin(CUDART_PI * (1 - r1d[j] - 0.5 * g.dx) / 2 / den / g.dx);
      rowp(w, g.ld, j)[k] = ramp * src;
    }
  }
  if (parts & 2) rowp(w, g.ld, 0)[k] = 0.0;
}

// -------------------------------------------------------------------------------------
// a12  kernels/periodic_boundary_ghost_comm.py:12-13 / :26-31
// -------------------------------------------------------------------------------------
__global__ void k_ghost(GridD g, double* __restrict__ f, int ghost, double z_max, double two_g_dx) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.nr * 2 * ghost) return;
  const int j = idx / (2 * ghost), c = idx % (2 * ghost);
  double* row = rowp(f, g.ld, j);
  if (c < ghost) {
    row[c] = (row[g.nz - 2 * ghost + c] - z_max) + two_g_dx;
  } else {
    const int i = c - ghost;
    row[g.nz - ghost + i] = (row[ghost + i] + z_max) - two_g_dx;
  }
}

// -------------------------------------------------------------------------------------
// G-VEL  kernels/compute_velocity_from_psi.py:9-16 (+ free stream, + max reduction)
// -------------------------------------------------------------------------------------
template <bool REDUCE>
__global__ void __launch_bounds__(TBX* TBY)
    k_velocity(GridD g, double* __restrict__ u_z, double* __restrict__ u_r, const double* __restrict__ psi,
               const double* __restrict__ r1d, double uz_add, double ur_add, const double* __restrict__ add_dev,
               double* umax_out, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  double local_max = 0.0;
  if (j < g.nr && k < g.ku1 && k + 1 >= g.ku0) {
    if (add_dev) { uz_add = add_dev[0]; ur_add = add_dev[1]; }
    const double h = 2 * g.dx;
    const double r = r1d[j];
    const int nz = g.nz;
    const double* pc = rowp(psi, g.ld, j);
    double2 uz, ur;
    // ---- u_z = d(psi)/dr / r : rows
    if (j > 0 && j < g.nr - 1) {
      const double2 up = ld_pair(rowp(psi, g.ld, j + 1), k, nz, vec);
      const double2 dn = ld_pair(rowp(psi, g.ld, j - 1), k, nz, vec);
      uz.x = (up.x - dn.x) / h / r;
      uz.y = (up.y - dn.y) / h / r;
    } else if (j == 0) {
      const double2 p0 = ld_pair(pc, k, nz, vec);
      const double2 p1 = ld_pair(rowp(psi, g.ld, 1), k, nz, vec);
      const double2 p2 = ld_pair(rowp(psi, g.ld, 2), k, nz, vec);
      uz.x = (-p2.x + 4 * p1.x - 3 * p0.x) / h / r;
      uz.y = (-p2.y + 4 * p1.y - 3 * p0.y) / h / r;
    } else {
      const double2 p0 = ld_pair(pc, k, nz, vec);
      const double2 p1 = ld_pair(rowp(psi, g.ld, j - 1), k, nz, vec);
      const double2 p2 = ld_pair(rowp(psi, g.ld, j - 2), k, nz, vec);
      uz.x = (p2.x - 4 * p1.x + 3 * p0.x) / h / r;
      uz.y = (p2.y - 4 * p1.y + 3 * p0.y) / h / r;
    }
    // ---- u_r = -d(psi)/dz / r : columns (global index decides the one-sided ends)
    double v[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int kk = k + c;
      if (kk < g.ku0 || kk >= g.ku1) { v[c] = 0.0; continue; }
      const int kg = kk + g.kz0;
      if (kg > 0 && kg < g.nzg - 1) {
        v[c] = -(pc[kk + 1] - pc[kk - 1]) / h / r;
      } else if (kg == 0) {
        v[c] = -(-pc[kk + 2] + 4 * pc[kk + 1] - 3 * pc[kk]) / h / r;
      } else if (kg == g.nzg - 1) {
        v[c] = -(pc[kk - 2] - 4 * pc[kk - 1] + 3 * pc[kk]) / h / r;
      } else {
        v[c] = 0.0;  // halo column outside the global domain: never owned
      }
    }
    ur.x = v[0]; ur.y = v[1];
    uz.x += uz_add; uz.y += uz_add;
    ur.x += ur_add; ur.y += ur_add;
    st_pair(rowp(u_z, g.ld, j), k, g.ku0, g.ku1, vec, uz);
    st_pair(rowp(u_r, g.ld, j), k, g.ku0, g.ku1, vec, ur);
    if (REDUCE && j >= g.ju0 && j < g.ju1) {
      if (k >= g.ku0 && k < g.ku1) local_max = fmax(local_max, fabs(uz.x) + fabs(ur.x));
      if (k + 1 >= g.ku0 && k + 1 < g.ku1) local_max = fmax(local_max, fabs(uz.y) + fabs(ur.y));
    }
  }
  if (REDUCE) {
    const double m = block_max(local_max);
    if (threadIdx.x == 0 && threadIdx.y == 0 && m > 0.0) atomic_max_nonneg(umax_out, m);
  }
}

// -------------------------------------------------------------------------------------
// a9  kernels/brinkmann_penalize.py:11-16
// -------------------------------------------------------------------------------------
__device__ __forceinline__ double pen1(double u, double lamdt, double chi, double U) {
  return (u + lamdt * chi * U) / (1 + lamdt * chi);
}

__global__ void __launch_bounds__(TBX* TBY)
    k_brinkmann(GridD g, double lamdt, const double* __restrict__ chi, double U_z, double U_r,
                const double* __restrict__ Uzf, const double* __restrict__ Urf, const double* __restrict__ gz,
                const double* __restrict__ gr, double* __restrict__ pz, double* __restrict__ pr, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  const int nz = g.nz;
  const double2 c = ld_pair(rowp(chi, g.ld, j), k, nz, vec);
  const double2 a = ld_pair(rowp(gz, g.ld, j), k, nz, vec);
  const double2 b = ld_pair(rowp(gr, g.ld, j), k, nz, vec);
  double2 Uz = make_double2(U_z, U_z), Ur = make_double2(U_r, U_r);
  if (Uzf) Uz = ld_pair(rowp(Uzf, g.ld, j), k, nz, vec);
  if (Urf) Ur = ld_pair(rowp(Urf, g.ld, j), k, nz, vec);
  st_pair(rowp(pz, g.ld, j), k, g.ku0, g.ku1, vec,
          make_double2(pen1(a.x, lamdt, c.x, Uz.x), pen1(a.y, lamdt, c.y, Uz.y)));
  st_pair(rowp(pr, g.ld, j), k, g.ku0, g.ku1, vec,
          make_double2(pen1(b.x, lamdt, c.x, Ur.x), pen1(b.y, lamdt, c.y, Ur.y)));
}

// -------------------------------------------------------------------------------------
// a10  kernels/compute_vorticity_from_velocity.py:11-13 (optionally of a difference, +=)
// -------------------------------------------------------------------------------------
template <bool SUB, bool ACC>
__global__ void __launch_bounds__(TBX* TBY)
    k_curl(GridD g, double* __restrict__ vort, const double* __restrict__ u_z, const double* __restrict__ u_r,
           const double* __restrict__ s_z, const double* __restrict__ s_r, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j < 1 || j >= g.nr - 1 || k >= g.ku1 || k + 1 < g.ku0) return;
  const int nz = g.nz;
  const double h = 2 * g.dx;
  double2 up = ld_pair(rowp(u_z, g.ld, j + 1), k, nz, vec);
  double2 dn = ld_pair(rowp(u_z, g.ld, j - 1), k, nz, vec);
  if (SUB) {
    const double2 a = ld_pair(rowp(s_z, g.ld, j + 1), k, nz, vec);
    const double2 b = ld_pair(rowp(s_z, g.ld, j - 1), k, nz, vec);
    up.x -= a.x; up.y -= a.y; dn.x -= b.x; dn.y -= b.y;
  }
  const double* ur = rowp(u_r, g.ld, j);
  const double* sr = SUB ? rowp(s_r, g.ld, j) : nullptr;
  double* out = rowp(vort, g.ld, j);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int kk = k + c;
    if (kk < g.ku0 || kk >= g.ku1) continue;
    const int kg = kk + g.kz0;
    if (kg < 1 || kg > g.nzg - 2) continue;
    double rp = ur[kk + 1], rm = ur[kk - 1];
    if (SUB) { rp -= sr[kk + 1]; rm -= sr[kk - 1]; }
    const double dzu = c ? (up.y - dn.y) : (up.x - dn.x);
    const double curl = (rp - rm) / h - dzu / h;
    out[kk] = ACC ? (out[kk] + curl) : curl;
  }
}

// -------------------------------------------------------------------------------------
// G-PEN  flow_past_sphere.py:155-175 in one pass (see header)
// -------------------------------------------------------------------------------------
template <bool REDUCE>
__global__ void __launch_bounds__(TBX* TBY)
    k_penalise(GridD g, double* __restrict__ u_z, double* __restrict__ u_r, double* __restrict__ w,
               const double* __restrict__ uzu, const double* __restrict__ uru, const double* __restrict__ chi,
               double lam, double dt, const double* __restrict__ dt_dev, double U_z, double U_r,
               const double* __restrict__ U_dev, const double* __restrict__ r1d, double* sum_out, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  double local = 0.0;
  if (j < g.nr && k < g.ku1 && k + 1 >= g.ku0) {
    if (dt_dev) dt = *dt_dev;
    if (U_dev) { U_z = U_dev[0]; U_r = U_dev[1]; }
    const double lamdt = lam * dt;
    const int nz = g.nz;
    const double h = 2 * g.dx;
    const double2 c0 = ld_pair(rowp(chi, g.ld, j), k, nz, vec);
    const double2 z0 = ld_pair(rowp(uzu, g.ld, j), k, nz, vec);
    const double2 r0 = ld_pair(rowp(uru, g.ld, j), k, nz, vec);
    const double2 pz = make_double2(pen1(z0.x, lamdt, c0.x, U_z), pen1(z0.y, lamdt, c0.y, U_z));
    const double2 pr = make_double2(pen1(r0.x, lamdt, c0.x, U_r), pen1(r0.y, lamdt, c0.y, U_r));
    st_pair(rowp(u_z, g.ld, j), k, g.ku0, g.ku1, vec, pz);
    st_pair(rowp(u_r, g.ld, j), k, g.ku0, g.ku1, vec, pr);
    if (REDUCE && j >= g.ju0 && j < g.ju1) {
      const double r = r1d[j];
      if (k >= g.ku0 && k < g.ku1) local += r * c0.x * (pz.x - U_z);
      if (k + 1 >= g.ku0 && k + 1 < g.ku1) local += r * c0.y * (pz.y - U_z);
    }
    if (j >= 1 && j < g.nr - 1) {
      // d(u_z - u_z_upen)/dr from the rows above and below
      const double2 cu = ld_pair(rowp(chi, g.ld, j + 1), k, nz, vec);
      const double2 cd = ld_pair(rowp(chi, g.ld, j - 1), k, nz, vec);
      const double2 zu = ld_pair(rowp(uzu, g.ld, j + 1), k, nz, vec);
      const double2 zd = ld_pair(rowp(uzu, g.ld, j - 1), k, nz, vec);
      const double dzx = (pen1(zu.x, lamdt, cu.x, U_z) - zu.x) - (pen1(zd.x, lamdt, cd.x, U_z) - zd.x);
      const double dzy = (pen1(zu.y, lamdt, cu.y, U_z) - zu.y) - (pen1(zd.y, lamdt, cd.y, U_z) - zd.y);
      const double* cr = rowp(chi, g.ld, j);
      const double* rr = rowp(uru, g.ld, j);
      double* wr = rowp(w, g.ld, j);
      // neighbours in z: k-1 and k+2 come from L1, the pair itself is in registers
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int kk = k + c;
        if (kk < g.ku0 || kk >= g.ku1) continue;
        const int kg = kk + g.kz0;
        if (kg < 1 || kg > g.nzg - 2) continue;
        double dl, dr;  // (pen - upen) of u_r at kk-1 and kk+1
        if (c == 0) {
          const double ul = rr[kk - 1];
          dl = pen1(ul, lamdt, cr[kk - 1], U_r) - ul;
          dr = pr.y - r0.y;
        } else {
          const double un = rr[kk + 1];
          dl = pr.x - r0.x;
          dr = pen1(un, lamdt, cr[kk + 1], U_r) - un;
        }
        const double curl = (dr - dl) / h - (c ? dzy : dzx) / h;
        wr[kk] = wr[kk] + curl;
      }
    }
  }
  if (REDUCE) {
    const double s = block_sum(local);
    if (threadIdx.x == 0 && threadIdx.y == 0 && s != 0.0) atomicAdd(sum_out, s);
  }
}

// -------------------------------------------------------------------------------------
// G-DIF  kernels/diffusion_RK2.py:9-44
// -------------------------------------------------------------------------------------
__device__ __forceinline__ double diff_op(double up, double dn, double rt, double lf, double c, double dx, double r) {
  return (up + dn + rt + lf - 4 * c) / (dx * dx) + (up - dn) / (2 * dx) / r - c * (1.0 / (r * r));
}

// STAGE 1: out = in (everywhere) ; out[int] += coef * L(in)        (out = tmp, in = w)
// STAGE 2: out[int] += coef * L(in)                                 (out = w,   in = tmp)
template <int STAGE>
__global__ void __launch_bounds__(TBX* TBY)
    k_diffusion(GridD g, double* out, const double* __restrict__ in, const double* src2,
                const double* __restrict__ r1d, double nu, double dt, const double* __restrict__ dt_dev, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  if (dt_dev) dt = *dt_dev;
  const double coef = (STAGE == 1) ? (0.5 * nu * dt) : (nu * dt);
  const int nz = g.nz;
  const double* ic = rowp(in, g.ld, j);
  const double2 c = ld_pair(ic, k, nz, vec);
  double2 res;
  if (STAGE == 1) res = c;
  else res = ld_pair(rowp(src2, g.ld, j), k, nz, vec);
  if (j >= 1 && j < g.nr - 1) {
    const double r = r1d[j];
    const double2 up = ld_pair(rowp(in, g.ld, j + 1), k, nz, vec);
    const double2 dn = ld_pair(rowp(in, g.ld, j - 1), k, nz, vec);
    {
      const int kg = k + g.kz0;
      if (kg >= 1 && kg <= g.nzg - 2 && k >= g.ku0)
        res.x += coef * diff_op(up.x, dn.x, c.y, ic[k - 1], c.x, g.dx, r);
    }
    {
      const int kk = k + 1, kg = kk + g.kz0;
      if (kk < g.ku1 && kg >= 1 && kg <= g.nzg - 2)
        res.y += coef * diff_op(up.y, dn.y, ic[kk + 1], c.x, c.y, g.dx, r);
    }
  }
  st_pair(rowp(out, g.ld, j), k, g.ku0, g.ku1, vec, res);
}

// -------------------------------------------------------------------------------------
// G-HEAV  kernels/smooth_Heaviside.py:10-14
// -------------------------------------------------------------------------------------
// The reference evaluates the blend expression everywhere and multiplies it by the 0/1 band indicator; outside the
// band that product is an exact zero, so only band cells need the division and the sine (whose argument is in the
// thousands far from the interface: the slow range-reduction path).  Same bits, ~5x less FP64 work per field.
__device__ __forceinline__ double heav1(double phi, double w) {
  if (fabs(phi) < w) return 0.0 + 0.5 * (1 + phi / w + sin(CUDART_PI * phi / w) / CUDART_PI);
  return (phi >= w) ? 1.0 : 0.0;
}

template <bool SPHERE>
__global__ void __launch_bounds__(TBX* TBY)
    k_heaviside(GridD g, double* __restrict__ H, double* __restrict__ phi_out, const double* __restrict__ phi,
                const double* __restrict__ z1d, const double* __restrict__ r1d, double z_cm, double r_cm,
                double radius, double w, bool vec, const double* __restrict__ z_cm_dev = nullptr) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  H += member_field(g);
  phi_out = moved(phi_out, member_field(g));
  phi = moved(phi, member_field(g));
  double2 p;
  if (SPHERE) {
    if (z_cm_dev) z_cm = z_cm_dev[member_scalar(g)];
    const double dr = r1d[j] - r_cm;
    const double za = z1d[k] - z_cm, zb = z1d[(k + 1 < g.nz) ? k + 1 : k] - z_cm;
    const double da = za * za + dr * dr, db = zb * zb + dr * dr;
    if (!phi_out) {
      // phi itself is not wanted: cells safely inside / outside the blend shell (1e-9 relative margin on the squared
      // distance, far above the rounding of sqrt and of the subtraction) are 1 / 0 without the square root
      const double lo = radius - w, hi = radius + w;
      const double lo2 = (lo > 1e-3 * radius) ? lo * lo * (1 - 1e-9) : -1.0, hi2 = hi * hi * (1 + 1e-9);
      const bool in = da < lo2 && db < lo2, out = da > hi2 && db > hi2;
      if (in || out) {
        const double v = in ? 1.0 : 0.0;
        st_pair(rowp(H, g.ld, j), k, g.ku0, g.ku1, vec, make_double2(v, v));
        return;
      }
    }
    p.x = -sqrt(da) + radius;
    p.y = -sqrt(db) + radius;
    if (phi_out) st_pair(rowp(phi_out, g.ld, j), k, g.ku0, g.ku1, vec, p);
  } else {
    p = ld_pair(rowp(phi, g.ld, j), k, g.nz, vec);
  }
  st_pair(rowp(H, g.ld, j), k, g.ku0, g.ku1, vec, make_double2(heav1(p.x, w), heav1(p.y, w)));
}

// kernels/vortex_stretching.py:8-11
__global__ void __launch_bounds__(TBX* TBY)
    k_stretch(GridD g, double* __restrict__ w, const double* __restrict__ u_r, const double* __restrict__ r1d,
              double dt, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j < 1 || j >= g.nr - 1 || k >= g.ku1 || k + 1 < g.ku0) return;
  double* wr = rowp(w, g.ld, j);
  const double* ur = rowp(u_r, g.ld, j);
  const double r = r1d[j];
  for (int c = 0; c < 2; ++c) {
    const int kk = k + c, kg = kk + g.kz0;
    if (kk < g.ku0 || kk >= g.ku1 || kg < 1 || kg > g.nzg - 2) continue;
    wr[kk] = wr[kk] + dt * ur[kk] * wr[kk] / r;
  }
}

// pyst_kernels/elementwise_ops.py:28-33 and :73-77
__global__ void __launch_bounds__(TBX* TBY)
    k_sum(GridD g, double* __restrict__ s, const double* __restrict__ a, const double* __restrict__ b, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  const double2 x = ld_pair(rowp(a, g.ld, j), k, g.nz, vec), y = ld_pair(rowp(b, g.ld, j), k, g.nz, vec);
  st_pair(rowp(s, g.ld, j), k, g.ku0, g.ku1, vec, make_double2(x.x + y.x, x.y + y.y));
}
__global__ void __launch_bounds__(TBX* TBY) k_fill(GridD g, double* __restrict__ f, double v, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  st_pair(rowp(f, g.ld, j), k, g.ku0, g.ku1, vec, make_double2(v, v));
}

// -------------------------------------------------------------------------------------
// a15 diagnostics
// -------------------------------------------------------------------------------------
// MODE 0: max(|a|+|b|) (b optional)  MODE 1: max(a)  MODE 2: sum(r*c*(a-off))
template <int MODE>
__global__ void __launch_bounds__(TBX* TBY)
    k_reduce(GridD g, const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ r1d,
             double off, double* out, bool vec) {
  a += member_field(g);
  b = moved(b, member_field(g));
  out += member_scalar(g);
  double acc = (MODE == 1) ? -INFINITY : 0.0;
  // grid-stride over row blocks so the number of atomics stays small
  for (int j = blockIdx.y * TBY + threadIdx.y; j < g.nr; j += gridDim.y * TBY) {
    for (int k = 2 * (blockIdx.x * TBX + threadIdx.x); k < g.ku1; k += 2 * gridDim.x * TBX) {
      if (k + 1 < g.ku0) continue;
      const double2 x = ld_pair(rowp(a, g.ld, j), k, g.nz, vec);
      double2 y = make_double2(0.0, 0.0);
      if (MODE != 1 && b) y = ld_pair(rowp(b, g.ld, j), k, g.nz, vec);
      const bool in0 = (k >= g.ku0), in1 = (k + 1 < g.ku1);
      if (MODE == 0) {
        if (in0) acc = fmax(acc, fabs(x.x) + fabs(y.x));
        if (in1) acc = fmax(acc, fabs(x.y) + fabs(y.y));
      } else if (MODE == 1) {
        if (in0) acc = fmax(acc, x.x);
        if (in1) acc = fmax(acc, x.y);
      } else {
        const double r = r1d[j];
        if (in0) acc += r * y.x * (x.x - off);
        if (in1) acc += r * y.y * (x.y - off);
      }
    }
  }
  if (MODE == 2) {
    const double s = block_sum(acc);
    if (threadIdx.x == 0 && threadIdx.y == 0) atomicAdd(out, s);
  } else {
    const double m = block_max(acc);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      if (MODE == 0) atomic_max_nonneg(out, m);
      else atomic_max_any(out, m);
    }
  }
}

__global__ void k_fill_scalars(double* dst, int n, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

// flow_past_sphere.py:117-121, :150-153, :186-188 on device scalars (see header)
__global__ void k_rigid_scalars(int phase, double* st, double U0, double T_ramp, double ur_ramp, double dt_lim,
                                double cfl_dx) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (phase == 0) {
    const double t = st[0];
    double pre = 1.0, pre_r = 0.0;
    if (t < T_ramp) {
      pre = sin(0.5 * CUDART_PI * t / T_ramp);
      pre_r = ur_ramp * sin(CUDART_PI * t / T_ramp);
    }
    st[4] = U0 * pre;
    st[5] = U0 * pre_r;
    st[2] = 0.0;
    st[3] = 0.0;
  } else if (phase == 1) {
    const double eps = 2.220446049250313e-16;
    st[1] = fmin(dt_lim, cfl_dx / (st[2] + eps));
  } else {
    st[0] = st[0] + st[1];
    st[6] = st[6] + 1.0;
    st[7] = st[3];
  }
}

inline bool vec_ok(const GridD& g, std::initializer_list<const void*> ptrs) {
  if ((g.ld & 1) || (g.bstride & 1)) return false;
  for (const void* p : ptrs)
    if (p && !axb_al16(p)) return false;
  return true;
}
inline int al_check(std::initializer_list<const void*> ptrs) {
  for (const void* p : ptrs)
    if (p && !axb_al8(p)) return AXB_EALIGN;
  return AXB_OK;
}

}  // namespace

#define GRID_PROLOGUE_(CHECK, ...)               \
  int rc__ = CHECK(g);                          \
  if (rc__) return rc__;                        \
  rc__ = al_check({__VA_ARGS__});               \
  if (rc__) return rc__;                        \
  const GridD d = to_dev(g);                    \
  const bool vec = vec_ok(d, {__VA_ARGS__});    \
  const dim3 blk(TBX, TBY), grd = grid2d(d);    \
  (void)vec; (void)blk; (void)grd;
#define GRID_PROLOGUE(...) GRID_PROLOGUE_(axb_check_grid, __VA_ARGS__)
// entries that serve an ensemble with one launch: the grid's z dimension is the member (axb_grid_t.batch)
#define GRID_PROLOGUE_BATCHED(...) GRID_PROLOGUE_(axb_check_grid_batched, __VA_ARGS__)

extern "C" {

int axb_kill_boundary_vorticity_sine_z(const axb_grid_t* g, double* w, const double* z1d, int width,
                                       axb_stream_t s) {
  if (!w || !z1d || width < 2) return AXB_EINVAL;
  GRID_PROLOGUE_BATCHED(w, z1d)
  if (2 * width > d.nzg) return AXB_EINVAL;
  k_kill_z<<<dim3((d.nr + 127) / 128, 1, d.batch), 128, 0, s>>>(d, w, z1d, width);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_kill_boundary_vorticity_sine_r_parts(const axb_grid_t* g, double* w, const double* r1d, int width,
                                             int parts, axb_stream_t s) {
  if (!w || !r1d || width < 2 || parts < 0 || parts > 3) return AXB_EINVAL;
  GRID_PROLOGUE_BATCHED(w, r1d)
  if (width > d.nr) return AXB_EINVAL;
  const int n = d.ku1 - d.ku0;
  if (n <= 0 || !parts) return AXB_OK;
  k_kill_r<<<dim3((n + 127) / 128, 1, d.batch), 128, 0, s>>>(d, w, r1d, width, parts);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_kill_boundary_vorticity_sine_r(const axb_grid_t* g, double* w, const double* r1d, int width,
                                       axb_stream_t s) {
  return axb_kill_boundary_vorticity_sine_r_parts(g, w, r1d, width, 3, s);
}

int axb_periodic_ghost_comm(const axb_grid_t* g, double* f, int ghost, double z_max, double two_g_dx,
                            axb_stream_t s) {
  if (!f || ghost < 1) return AXB_EINVAL;
  GRID_PROLOGUE(f)
  if (4 * ghost > d.nz) return AXB_EINVAL;
  const int n = d.nr * 2 * ghost;
  k_ghost<<<(n + 255) / 256, 256, 0, s>>>(d, f, ghost, z_max, two_g_dx);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_velocity_from_psi(const axb_grid_t* g, double* u_z, double* u_r, const double* psi,
                          const double* r1d, double uz_add, double ur_add, const double* add_dev,
                          double* umax_out, axb_stream_t s) {
  if (!u_z || !u_r || !psi || !r1d) return AXB_EINVAL;
  GRID_PROLOGUE_BATCHED(u_z, u_r, psi)
  if (d.nr < 3 || d.nzg < 3) return AXB_EINVAL;
  if (d.batch > 1 && g_axb_legacy_stencils) return AXB_ENOSUP;     // ensembles: row-marching kernels only
  if (!g_axb_legacy_stencils) {
    const int rc = march_velocity(d, u_z, u_r, psi, r1d, uz_add, ur_add, add_dev, umax_out, vec, s);
    AXB_LAUNCHED();
    return rc;
  }
  if (umax_out)
    k_velocity<true><<<grd, blk, 0, s>>>(d, u_z, u_r, psi, r1d, uz_add, ur_add, add_dev, umax_out, vec);
  else
    k_velocity<false><<<grd, blk, 0, s>>>(d, u_z, u_r, psi, r1d, uz_add, ur_add, add_dev, nullptr, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_brinkmann_penalize(const axb_grid_t* g, double lam, double dt, const double* chi, double U_z,
                           double U_r, const double* U_z_field, const double* U_r_field,
                           const double* grid_u_z, const double* grid_u_r, double* pen_u_z,
                           double* pen_u_r, axb_stream_t s) {
  if (!chi || !grid_u_z || !grid_u_r || !pen_u_z || !pen_u_r) return AXB_EINVAL;
  GRID_PROLOGUE(chi, U_z_field, U_r_field, grid_u_z, grid_u_r, pen_u_z, pen_u_r)
  k_brinkmann<<<grd, blk, 0, s>>>(d, lam * dt, chi, U_z, U_r, U_z_field, U_r_field, grid_u_z, grid_u_r,
                                  pen_u_z, pen_u_r, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_vorticity_from_velocity(const axb_grid_t* g, double* vort, const double* u_z, const double* u_r,
                                const double* u_z_sub, const double* u_r_sub, int accumulate,
                                axb_stream_t s) {
  if (!vort || !u_z || !u_r || ((u_z_sub == nullptr) != (u_r_sub == nullptr))) return AXB_EINVAL;
  GRID_PROLOGUE(vort, u_z, u_r, u_z_sub, u_r_sub)
  if (u_z_sub) {
    if (accumulate) k_curl<true, true><<<grd, blk, 0, s>>>(d, vort, u_z, u_r, u_z_sub, u_r_sub, vec);
    else k_curl<true, false><<<grd, blk, 0, s>>>(d, vort, u_z, u_r, u_z_sub, u_r_sub, vec);
  } else {
    if (accumulate) k_curl<false, true><<<grd, blk, 0, s>>>(d, vort, u_z, u_r, nullptr, nullptr, vec);
    else k_curl<false, false><<<grd, blk, 0, s>>>(d, vort, u_z, u_r, nullptr, nullptr, vec);
  }
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_penalise_update_vorticity(const axb_grid_t* g, double* u_z, double* u_r, double* w,
                                  const double* u_z_upen, const double* u_r_upen, const double* chi,
                                  double lam, double dt, const double* dt_dev, double U_z, double U_r,
                                  const double* U_dev, const double* r1d, double* sum_out,
                                  axb_stream_t s) {
  if (!u_z || !u_r || !w || !u_z_upen || !u_r_upen || !chi) return AXB_EINVAL;
  if (u_z == u_z_upen || u_r == u_r_upen) return AXB_EINVAL;  // neighbours are read un-penalised
  if (sum_out && !r1d) return AXB_EINVAL;
  GRID_PROLOGUE_BATCHED(u_z, u_r, w, u_z_upen, u_r_upen, chi)
  if (d.batch > 1 && g_axb_legacy_stencils) return AXB_ENOSUP;
  if (!g_axb_legacy_stencils) {
    const int rc = march_penalise(d, u_z, u_r, w, u_z_upen, u_r_upen, chi, lam, dt, dt_dev, U_z, U_r, U_dev, r1d,
                                  sum_out, vec, s);
    AXB_LAUNCHED();
    return rc;
  }
  if (sum_out)
    k_penalise<true><<<grd, blk, 0, s>>>(d, u_z, u_r, w, u_z_upen, u_r_upen, chi, lam, dt, dt_dev, U_z, U_r,
                                         U_dev, r1d, sum_out, vec);
  else
    k_penalise<false><<<grd, blk, 0, s>>>(d, u_z, u_r, w, u_z_upen, u_r_upen, chi, lam, dt, dt_dev, U_z, U_r,
                                          U_dev, r1d, nullptr, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

static int diffusion_stage(int stage, const axb_grid_t* g, double* out, const double* in, const double* src2,
                           const double* r1d, double nu, const double* nu_dev, double dt, const double* dt_dev,
                           axb_stream_t s) {
  GRID_PROLOGUE_BATCHED(out, in, src2)
  if (g_axb_legacy_stencils && (d.batch > 1 || nu_dev)) return AXB_ENOSUP;
  if (!g_axb_legacy_stencils) {
    const int rc = march_diffusion(stage, d, out, in, src2, r1d, nu, nu_dev, dt, dt_dev, vec, s);
    AXB_LAUNCHED();
    return rc;
  }
  if (stage == 1) k_diffusion<1><<<grd, blk, 0, s>>>(d, out, in, nullptr, r1d, nu, dt, dt_dev, vec);
  else k_diffusion<2><<<grd, blk, 0, s>>>(d, out, in, src2, r1d, nu, dt, dt_dev, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_diffusion_rk2_stage1(const axb_grid_t* g, double* tmp, const double* w, const double* r1d,
                             double nu, double dt, const double* dt_dev, axb_stream_t s) {
  if (!tmp || !w || !r1d || tmp == w) return AXB_EINVAL;
  return diffusion_stage(1, g, tmp, w, nullptr, r1d, nu, nullptr, dt, dt_dev, s);
}
int axb_diffusion_rk2_stage2(const axb_grid_t* g, double* w, const double* w_src, const double* tmp,
                             const double* r1d, double nu, double dt, const double* dt_dev,
                             axb_stream_t s) {
  if (!tmp || !w || !w_src || !r1d || tmp == w) return AXB_EINVAL;
  return diffusion_stage(2, g, w, tmp, w_src, r1d, nu, nullptr, dt, dt_dev, s);
}
int axb_diffusion_rk2_stage1_dev(const axb_grid_t* g, double* tmp, const double* w, const double* r1d,
                                 const double* nu_dev, const double* dt_dev, axb_stream_t s) {
  if (!tmp || !w || !r1d || tmp == w || !nu_dev || !dt_dev) return AXB_EINVAL;
  return diffusion_stage(1, g, tmp, w, nullptr, r1d, 0.0, nu_dev, 0.0, dt_dev, s);
}
int axb_diffusion_rk2_stage2_dev(const axb_grid_t* g, double* w, const double* w_src, const double* tmp,
                                 const double* r1d, const double* nu_dev, const double* dt_dev, axb_stream_t s) {
  if (!tmp || !w || !w_src || !r1d || tmp == w || !nu_dev || !dt_dev) return AXB_EINVAL;
  return diffusion_stage(2, g, w, tmp, w_src, r1d, 0.0, nu_dev, 0.0, dt_dev, s);
}

int axb_diffusion_rk2_fused(const axb_grid_t* g, double* w, const double* w_src, double* tmp, const double* r1d,
                            double nu, double dt, const double* dt_dev, axb_stream_t s) {
  if (!w || !w_src || !r1d || w == w_src) return AXB_EINVAL;
  GRID_PROLOGUE(w, w_src)
  if (!g_axb_legacy_stencils) {
    const int rc = march_diffusion_fused(d, w, w_src, r1d, nu, dt, dt_dev, vec, s);
    AXB_LAUNCHED();
    return rc;
  }
  // 2-D tiled path: the two stages through the caller's scratch field
  if (!tmp || tmp == w || tmp == w_src) return AXB_EINVAL;
  k_diffusion<1><<<grd, blk, 0, s>>>(d, tmp, w_src, nullptr, r1d, nu, dt, dt_dev, vec);
  AXB_LAUNCHED();
  k_diffusion<2><<<grd, blk, 0, s>>>(d, w, tmp, w_src, r1d, nu, dt, dt_dev, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_smooth_heaviside(const axb_grid_t* g, double* H, const double* phi, double blend_w,
                         axb_stream_t s) {
  if (!H || !phi) return AXB_EINVAL;
  GRID_PROLOGUE(H, phi)
  k_heaviside<false><<<grd, blk, 0, s>>>(d, H, nullptr, phi, nullptr, nullptr, 0, 0, 0, blend_w, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_smooth_heaviside_sphere(const axb_grid_t* g, double* H, double* phi_out, const double* z1d,
                                const double* r1d, double z_cm, double r_cm, double radius,
                                double blend_w, axb_stream_t s) {
  if (!H || !z1d || !r1d) return AXB_EINVAL;
  GRID_PROLOGUE(H, phi_out)
  k_heaviside<true><<<grd, blk, 0, s>>>(d, H, phi_out, nullptr, z1d, r1d, z_cm, r_cm, radius, blend_w, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_smooth_heaviside_sphere_dev(const axb_grid_t* g, double* H, double* phi_out, const double* z1d,
                                    const double* r1d, const double* z_cm_dev, double r_cm, double radius,
                                    double blend_w, axb_stream_t s) {
  if (!H || !z1d || !r1d || !z_cm_dev) return AXB_EINVAL;
  GRID_PROLOGUE_BATCHED(H, phi_out)
  k_heaviside<true><<<grd, blk, 0, s>>>(d, H, phi_out, nullptr, z1d, r1d, 0.0, r_cm, radius, blend_w, vec, z_cm_dev);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_vortex_stretching(const axb_grid_t* g, double* w, const double* u_r, const double* r1d, double dt,
                          axb_stream_t s) {
  if (!w || !u_r || !r1d) return AXB_EINVAL;
  GRID_PROLOGUE(w, u_r)
  k_stretch<<<grd, blk, 0, s>>>(d, w, u_r, r1d, dt, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_elementwise_sum(const axb_grid_t* g, double* sum, const double* f1, const double* f2,
                        axb_stream_t s) {
  if (!sum || !f1 || !f2) return AXB_EINVAL;
  GRID_PROLOGUE(sum, f1, f2)
  k_sum<<<grd, blk, 0, s>>>(d, sum, f1, f2, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_set_fixed_val(const axb_grid_t* g, double* f, double val, axb_stream_t s) {
  if (!f) return AXB_EINVAL;
  GRID_PROLOGUE(f)
  k_fill<<<grd, blk, 0, s>>>(d, f, val, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

static dim3 reduce_grid(const GridD& d) {
  dim3 full = grid2d(d);
  // ~4 blocks per SM is plenty to saturate HBM; fewer blocks = fewer atomics
  const unsigned maxb = 148 * 4;
  unsigned gx = full.x, gy = full.y;
  while ((unsigned long long)gx * gy > maxb && gy > 1) gy = (gy + 1) / 2;
  while ((unsigned long long)gx * gy > maxb && gx > 1) gx = (gx + 1) / 2;
  return dim3(gx, gy, d.batch);
}

int axb_reduce_max_abs_sum(const axb_grid_t* g, const double* a, const double* b, double* out,
                           axb_stream_t s) {
  if (!a || !out) return AXB_EINVAL;
  GRID_PROLOGUE_BATCHED(a, b)
  k_reduce<0><<<reduce_grid(d), blk, 0, s>>>(d, a, b, nullptr, 0.0, out, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_reduce_max(const axb_grid_t* g, const double* a, double* out, axb_stream_t s) {
  if (!a || !out) return AXB_EINVAL;
  GRID_PROLOGUE(a)
  k_reduce<1><<<reduce_grid(d), blk, 0, s>>>(d, a, nullptr, nullptr, 0.0, out, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_reduce_weighted_sum(const axb_grid_t* g, const double* r1d, const double* c, const double* a,
                            double off, double* out, axb_stream_t s) {
  if (!a || !c || !r1d || !out) return AXB_EINVAL;
  GRID_PROLOGUE(a, c)
  k_reduce<2><<<reduce_grid(d), blk, 0, s>>>(d, a, c, r1d, off, out, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_fill_scalars(double* dst, int n, double val, axb_stream_t s) {
  if (!dst || n < 1) return AXB_EINVAL;
  k_fill_scalars<<<(n + 127) / 128, 128, 0, s>>>(dst, n, val);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_rigid_flow_scalars(int phase, double* state, double U0, double T_ramp, double ur_ramp,
                           double dt_diff_limit, double cfl_dx, axb_stream_t s) {
  if (!state || phase < 0 || phase > 2) return AXB_EINVAL;
  k_rigid_scalars<<<1, 32, 0, s>>>(phase, state, U0, T_ramp, ur_ramp, dt_diff_limit, cfl_dx);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

}  // extern "C"
