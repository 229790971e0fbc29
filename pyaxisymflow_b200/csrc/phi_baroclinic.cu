// phi_baroclinic.cu -- kernels of the SURVEY.md section 8f "next" rows (sm_100a).
//
//   8f-2  kernels/compute_velocity_from_phi.py:4-17        -> axb_velocity_from_phi
//   8f-4  kernels/update_baroclinic_vorticity.py:4-127     -> axb_baroclinic_vorticity_update
//
// Same conventions as stencils.cu: a thread owns two adjacent z columns, a block covers
// 64 columns x 8 rows (r-neighbours through L1), global z index decides the one-sided ends so
// the kernels also run on z-slabs.  Compiled with -fmad=false; the operation order is the
// reference's NumPy/numba expression order, so results are bit-identical to the oracle.
#include <initializer_list>

#include "axb_common.cuh"

extern int g_axb_legacy_stencils;  // capi.cu: 1 = 2-D tiled kernels only

namespace {

__device__ __forceinline__ const double* rowp(const double* f, long long ld, int j) { return f + (long long)j * ld; }
__device__ __forceinline__ double* rowp(double* f, long long ld, int j) { return f + (long long)j * ld; }

// -------------------------------------------------------------------------------------
// Interior fast path (same idea as stencils_march.cu): a block owns 256 adjacent columns (two per thread, 128-bit
// accesses) and marches over RBW rows with the r-neighbourhood in a rolling register window, so every input row
// is loaded once; z neighbours come from warp shuffles.  Blocks that touch a domain edge, a slab halo or a
// ragged end run the general per-row form on the same grid; small or unaligned grids take the 2-D tiled kernels.  Both evaluate the same
// expressions (true divisions), so a cell gets the same bits whichever kernel computes it.
// -------------------------------------------------------------------------------------
constexpr int MTW = 128, RBW = 16, URW = 4;

__device__ __forceinline__ bool march_interior(const GridD& g, int mbx, int j0, bool vec) {
  const int kb0 = 2 * mbx * MTW, kb1 = kb0 + 2 * MTW;
  return vec && (j0 >= 1) && (j0 + RBW + 1 <= g.nr) && (kb0 >= g.ku0) && (kb1 <= g.ku1) && (kb0 + g.kz0 >= 1) &&
         (kb1 - 1 + g.kz0 <= g.nzg - 2) && (kb0 >= 1) && (kb1 < g.nz);
}
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

__global__ void __launch_bounds__(MTW)
    km_velocity_phi(GridD g, double* __restrict__ u_z, double* __restrict__ u_r, const double* __restrict__ phi) {
  const int j0 = blockIdx.y * RBW;
  if (!march_interior(g, blockIdx.x, j0, true)) return;
  const int k = 2 * (blockIdx.x * MTW + threadIdx.x), lane = threadIdx.x & 31;
  const double h = 2 * g.dx;
  const long long ld = g.ld;
  const double* p = phi + (long long)(j0 - 1) * ld + k;
  double2 pm = ld2(p), pc = ld2(p + ld);
  p += 2 * ld;  // row j + 1
  double* oz = u_z + (long long)j0 * ld + k;
  double* orr = u_r + (long long)j0 * ld + k;
  for (int jb = 0; jb < RBW; jb += URW) {
    double2 pn[URW];
    double le[URW], re[URW];
#pragma unroll
    for (int u = 0; u < URW; ++u) pn[u] = ld2(p + u * ld);
#pragma unroll
    for (int u = 0; u < URW; ++u) {  // warp-edge z neighbours of rows j .. j+3
      const double* rr = p + (u - 1) * ld;
      le[u] = (lane == 0) ? rr[-1] : 0.0;
      re[u] = (lane == 31) ? rr[2] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < URW; ++u) {
      double left = __shfl_up_sync(0xffffffffu, pc.y, 1), right = __shfl_down_sync(0xffffffffu, pc.x, 1);
      if (lane == 0) left = le[u];
      if (lane == 31) right = re[u];
      st2(oz + u * ld, make_double2((pc.y - left) / h, (right - pc.x) / h));
      st2(orr + u * ld, make_double2((pn[u].x - pm.x) / h, (pn[u].y - pm.y) / h));
      pm = pc;
      pc = pn[u];
    }
    p += URW * ld;
    oz += URW * ld;
    orr += URW * ld;
  }
}

// -------------------------------------------------------------------------------------
// 8f-2  u_z = d(phi)/dz, u_r = d(phi)/dr, centred inside, second-order one-sided at the ends
// -------------------------------------------------------------------------------------
// general form for one column pair of one row: any row, global z ends by global index, slab ownership, any pitch
__device__ __forceinline__ void phi_pair(const GridD& g, double* __restrict__ u_z, double* __restrict__ u_r,
                                         const double* __restrict__ phi, int j, int k, bool vec) {
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  const double h = 2 * g.dx;
  const int nz = g.nz;
  const double* pc = rowp(phi, g.ld, j);
  double2 ur;
  if (j > 0 && j < g.nr - 1) {
    const double2 up = ld_pair(rowp(phi, g.ld, j + 1), k, nz, vec);
    const double2 dn = ld_pair(rowp(phi, g.ld, j - 1), k, nz, vec);
    ur.x = (up.x - dn.x) / h;
    ur.y = (up.y - dn.y) / h;
  } else if (j == 0) {
    const double2 p0 = ld_pair(pc, k, nz, vec);
    const double2 p1 = ld_pair(rowp(phi, g.ld, 1), k, nz, vec);
    const double2 p2 = ld_pair(rowp(phi, g.ld, 2), k, nz, vec);
    ur.x = (-p2.x + 4 * p1.x - 3 * p0.x) / h;
    ur.y = (-p2.y + 4 * p1.y - 3 * p0.y) / h;
  } else {
    const double2 p0 = ld_pair(pc, k, nz, vec);
    const double2 p1 = ld_pair(rowp(phi, g.ld, j - 1), k, nz, vec);
    const double2 p2 = ld_pair(rowp(phi, g.ld, j - 2), k, nz, vec);
    ur.x = (p2.x - 4 * p1.x + 3 * p0.x) / h;
    ur.y = (p2.y - 4 * p1.y + 3 * p0.y) / h;
  }
  double v[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int kk = k + c;
    v[c] = 0.0;
    if (kk < g.ku0 || kk >= g.ku1) continue;
    const int kg = kk + g.kz0;
    if (kg > 0 && kg < g.nzg - 1) {
      v[c] = (pc[kk + 1] - pc[kk - 1]) / h;
    } else if (kg == 0) {
      v[c] = (-pc[kk + 2] + 4 * pc[kk + 1] - 3 * pc[kk]) / h;
    } else if (kg == g.nzg - 1) {
      v[c] = (pc[kk - 2] - 4 * pc[kk - 1] + 3 * pc[kk]) / h;
    }
  }
  st_pair(rowp(u_z, g.ld, j), k, g.ku0, g.ku1, vec, make_double2(v[0], v[1]));
  st_pair(rowp(u_r, g.ld, j), k, g.ku0, g.ku1, vec, ur);
}

// 2-D tiled kernel: small grids, unaligned views, axb_set_stencil_path(1)
__global__ void __launch_bounds__(TBX* TBY)
    k_velocity_phi(GridD g, double* __restrict__ u_z, double* __restrict__ u_r, const double* __restrict__ phi,
                   bool vec) {
  phi_pair(g, u_z, u_r, phi, blockIdx.y * TBY + threadIdx.y, 2 * (blockIdx.x * TBX + threadIdx.x), vec);
}

// the blocks km_velocity_phi leaves out (same grid): domain edges, slab halos, ragged ends
__global__ void __launch_bounds__(MTW)
    km_velocity_phi_edge(GridD g, double* __restrict__ u_z, double* __restrict__ u_r, const double* __restrict__ phi) {
  const int j0 = blockIdx.y * RBW;
  if (march_interior(g, blockIdx.x, j0, true)) return;
  const int k = 2 * (blockIdx.x * MTW + threadIdx.x);
  for (int j = j0; j < min(j0 + RBW, g.nr); ++j) phi_pair(g, u_z, u_r, phi, j, k, true);
}

// -------------------------------------------------------------------------------------
// 8f-4  baroclinic vorticity source.  MODE 0: update_baroclinic_vorticity (:5-35),
//       MODE 1: ..._penal (:38-67), MODE 2: ..._diff_penal (:70-127).
// -------------------------------------------------------------------------------------
// exact form (the reference's operation order, true divisions) for one column pair of one row
template <int MODE>
__device__ __forceinline__ void baro_pair(const GridD& g, double* __restrict__ w, const double* __restrict__ u_z,
                                          const double* __restrict__ u_r, const double* __restrict__ o_z,
                                          const double* __restrict__ o_r, const double* __restrict__ rho,
                                          const double* __restrict__ p_z, const double* __restrict__ p_r,
                                          const double* __restrict__ r1d, double nu, double dt, int j, int k0) {
  if (j < 1 || j >= g.nr - 1) return;
  const double h = 2 * g.dx;
  const double* zc = rowp(u_z, g.ld, j);
  const double* zu = rowp(u_z, g.ld, j + 1);
  const double* zd = rowp(u_z, g.ld, j - 1);
  const double* rc = rowp(u_r, g.ld, j);
  const double* ru = rowp(u_r, g.ld, j + 1);
  const double* rd = rowp(u_r, g.ld, j - 1);
  const double* dc = rowp(rho, g.ld, j);
  const double* du = rowp(rho, g.ld, j + 1);
  const double* dd = rowp(rho, g.ld, j - 1);
  const double* ozc = rowp(o_z, g.ld, j);
  const double* orc = rowp(o_r, g.ld, j);
  double* out = rowp(w, g.ld, j);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int k = k0 + c;
    if (k < g.ku0 || k >= g.ku1) continue;
    const int kg = k + g.kz0;
    if (kg < 1 || kg > g.nzg - 2) continue;
    const double uz = zc[k], ur = rc[k];
    double Dz = (uz - ozc[k]) / dt + uz * (zc[k + 1] - zc[k - 1]) / h + ur * (zu[k] - zd[k]) / h;
    double Dr = (ur - orc[k]) / dt + uz * (rc[k + 1] - rc[k - 1]) / h + ur * (ru[k] - rd[k]) / h;
    if (MODE >= 1) {
      Dz = Dz - rowp(p_z, g.ld, j)[k];
      Dr = Dr - rowp(p_r, g.ld, j)[k];
    }
    if (MODE == 2) {
      const double r = r1d[j];
      const double dx2 = g.dx * g.dx;
      const double lz = (zu[k] + zd[k] + zc[k + 1] + zc[k - 1] - 4 * uz) / dx2 + (zu[k] - zd[k]) / h / r;
      const double lr =
          (ru[k] + rd[k] + rc[k + 1] + rc[k - 1] - 4 * ur) / dx2 + (ru[k] - rd[k]) / h / r - ur * (1.0 / (r * r));
      Dz = Dz - nu * lz;
      Dr = Dr - nu * lr;
    }
    const double src = dt * (Dz * (du[k] - dd[k]) / h - Dr * (dc[k + 1] - dc[k - 1]) / h) / dc[k];
    out[k] = out[k] + src;
  }
}

// 2-D tiled kernel: the reference's divisions bit for bit (small grids, unaligned views, axb_set_stencil_path(1))
template <int MODE>
__global__ void __launch_bounds__(TBX* TBY)
    k_baroclinic(GridD g, double* __restrict__ w, const double* __restrict__ u_z, const double* __restrict__ u_r,
                 const double* __restrict__ o_z, const double* __restrict__ o_r, const double* __restrict__ rho,
                 const double* __restrict__ p_z, const double* __restrict__ p_r, const double* __restrict__ r1d,
                 double nu, double dt) {
  baro_pair<MODE>(g, w, u_z, u_r, o_z, o_r, rho, p_z, p_r, r1d, nu, dt, blockIdx.y * TBY + threadIdx.y,
                  2 * (blockIdx.x * TBX + threadIdx.x));
}

// edge blocks of the march grid: exact form
template <int MODE>
__global__ void __launch_bounds__(MTW)
    km_baroclinic_edge(GridD g, double* __restrict__ w, const double* __restrict__ u_z, const double* __restrict__ u_r,
                       const double* __restrict__ o_z, const double* __restrict__ o_r, const double* __restrict__ rho,
                       const double* __restrict__ p_z, const double* __restrict__ p_r, const double* __restrict__ r1d,
                       double nu, double dt) {
  const int j0 = blockIdx.y * RBW;
  if (march_interior(g, blockIdx.x, j0, true)) return;
  const int k = 2 * (blockIdx.x * MTW + threadIdx.x);
  if (k >= g.nz) return;
  for (int j = j0; j < min(j0 + RBW, g.nr); ++j) baro_pair<MODE>(g, w, u_z, u_r, o_z, o_r, rho, p_z, p_r, r1d, nu, dt, j, k);
}

// Interior blocks, row marching.  The reference's expression has 9 (13 with the viscous terms) divisions per cell,
// which makes the exact form FP64-issue bound at a third to a half of the HBM rate; here the divisions by 2 dx, dt,
// dx^2 and r become multiplications by reciprocals computed once (<= a few ulp per term from the reference's
// sequence, like the marching kernels of stencils_march.cu) and only the division by the density stays.
struct Win3 {  // rows j-1, j, j+1 of one field for the thread's column pair
  double2 m, c, n;
};
__device__ __forceinline__ void z_nb(const double2 c, const double* __restrict__ row, int lane, double& left, double& right) {
  left = __shfl_up_sync(0xffffffffu, c.y, 1);
  right = __shfl_down_sync(0xffffffffu, c.x, 1);
  if (lane == 0) left = row[-1];
  if (lane == 31) right = row[2];
}

template <int MODE>
__global__ void __launch_bounds__(MTW)
    km_baroclinic(GridD g, double* __restrict__ w, const double* __restrict__ u_z, const double* __restrict__ u_r,
                  const double* __restrict__ o_z, const double* __restrict__ o_r, const double* __restrict__ rho,
                  const double* __restrict__ p_z, const double* __restrict__ p_r, const double* __restrict__ r1d,
                  double nu, double dt) {
  const int j0 = blockIdx.y * RBW;
  if (!march_interior(g, blockIdx.x, j0, true)) return;
  const int k = 2 * (blockIdx.x * MTW + threadIdx.x), lane = threadIdx.x & 31;
  const long long ld = g.ld;
  const double inv_h = 1.0 / (2 * g.dx), inv_dt = 1.0 / dt, inv_dx2 = 1.0 / (g.dx * g.dx);
  long long o = (long long)(j0 - 1) * ld + k;
  Win3 Z, R, D;
  Z.m = ld2(u_z + o); R.m = ld2(u_r + o); D.m = ld2(rho + o);
  o += ld;
  Z.c = ld2(u_z + o); R.c = ld2(u_r + o); D.c = ld2(rho + o);
  for (int j = j0; j < j0 + RBW; ++j, o += ld) {
    Z.n = ld2(u_z + o + ld); R.n = ld2(u_r + o + ld); D.n = ld2(rho + o + ld);
    const double2 oz = ld2(o_z + o), orr = ld2(o_r + o), wc = ld2(w + o);
    double2 pz = make_double2(0, 0), pr = make_double2(0, 0);
    if (MODE >= 1) { pz = ld2(p_z + o); pr = ld2(p_r + o); }
    double zl, zr, rl, rr, dl, dr;
    z_nb(Z.c, u_z + o, lane, zl, zr);
    z_nb(R.c, u_r + o, lane, rl, rr);
    z_nb(D.c, rho + o, lane, dl, dr);
    double inv_r = 0.0, inv_r2 = 0.0;
    if (MODE == 2) {
      const double r = r1d[j];
      inv_r = 1.0 / r;
      inv_r2 = 1.0 / (r * r);
    }
    double out[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double uz = c ? Z.c.y : Z.c.x, ur = c ? R.c.y : R.c.x;
      const double zL = c ? Z.c.x : zl, zR = c ? zr : Z.c.y, zU = c ? Z.n.y : Z.n.x, zD = c ? Z.m.y : Z.m.x;
      const double rL = c ? R.c.x : rl, rR = c ? rr : R.c.y, rU = c ? R.n.y : R.n.x, rD = c ? R.m.y : R.m.x;
      const double dL = c ? D.c.x : dl, dR = c ? dr : D.c.y, dU = c ? D.n.y : D.n.x, dD = c ? D.m.y : D.m.x;
      const double dC = c ? D.c.y : D.c.x;
      double Dz = (uz - (c ? oz.y : oz.x)) * inv_dt + uz * (zR - zL) * inv_h + ur * (zU - zD) * inv_h;
      double Dr = (ur - (c ? orr.y : orr.x)) * inv_dt + uz * (rR - rL) * inv_h + ur * (rU - rD) * inv_h;
      if (MODE >= 1) {
        Dz = Dz - (c ? pz.y : pz.x);
        Dr = Dr - (c ? pr.y : pr.x);
      }
      if (MODE == 2) {
        const double lz = (zU + zD + zR + zL - 4 * uz) * inv_dx2 + (zU - zD) * inv_h * inv_r;
        const double lr = (rU + rD + rR + rL - 4 * ur) * inv_dx2 + (rU - rD) * inv_h * inv_r - ur * inv_r2;
        Dz = Dz - nu * lz;
        Dr = Dr - nu * lr;
      }
      const double src = dt * (Dz * (dU - dD) * inv_h - Dr * (dR - dL) * inv_h) / dC;
      out[c] = (c ? wc.y : wc.x) + src;
    }
    st2(w + o, make_double2(out[0], out[1]));
    Z.m = Z.c; Z.c = Z.n;
    R.m = R.c; R.c = R.n;
    D.m = D.c; D.c = D.n;
  }
}

inline int al_check(std::initializer_list<const void*> ptrs) {
  for (const void* p : ptrs)
    if (p && !axb_al8(p)) return AXB_EALIGN;
  return AXB_OK;
}
inline bool vec_ok(const GridD& g, std::initializer_list<const void*> ptrs) {
  if (g.ld & 1) return false;
  for (const void* p : ptrs)
    if (p && !axb_al16(p)) return false;
  return true;
}

}  // namespace

extern "C" {

int axb_velocity_from_phi(const axb_grid_t* g, double* u_z, double* u_r, const double* phi, axb_stream_t s) {
  if (!u_z || !u_r || !phi) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  rc = al_check({u_z, u_r, phi});
  if (rc) return rc;
  const GridD d = to_dev(g);
  if (d.nr < 3 || d.nzg < 3) return AXB_EINVAL;
  const bool vec = vec_ok(d, {u_z, u_r, phi});
  const bool march = vec && !g_axb_legacy_stencils && d.nr >= RBW + 2 && d.nz >= 2 * MTW + 2;
  if (march) {
    const dim3 mg((d.nz + 2 * MTW - 1) / (2 * MTW), (d.nr + RBW - 1) / RBW);
    km_velocity_phi<<<mg, MTW, 0, s>>>(d, u_z, u_r, phi);
    AXB_LAUNCHED();
    km_velocity_phi_edge<<<mg, MTW, 0, s>>>(d, u_z, u_r, phi);
  } else {
    k_velocity_phi<<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, u_z, u_r, phi, vec);
  }
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_baroclinic_vorticity_update(const axb_grid_t* g, double* w, const double* u_z, const double* u_r,
                                    const double* old_u_z, const double* old_u_r, const double* density,
                                    const double* penal_z, const double* penal_r, const double* r1d, double nu,
                                    double dt, int mode, axb_stream_t s) {
  if (!w || !u_z || !u_r || !old_u_z || !old_u_r || !density) return AXB_EINVAL;
  if (mode < 0 || mode > 2) return AXB_EINVAL;
  if (mode >= 1 && (!penal_z || !penal_r)) return AXB_EINVAL;
  if (mode == 2 && !r1d) return AXB_EINVAL;
  if (w == u_z || w == u_r || w == density) return AXB_EINVAL;  // neighbours are read while w is written
  int rc = axb_check_grid(g);
  if (rc) return rc;
  rc = al_check({w, u_z, u_r, old_u_z, old_u_r, density, penal_z, penal_r, r1d});
  if (rc) return rc;
  const GridD d = to_dev(g);
  if (d.nr < 3 || d.nzg < 3) return AXB_OK;  // no interior: the reference's slices are empty
  const bool vec = vec_ok(d, {w, u_z, u_r, old_u_z, old_u_r, density, penal_z, penal_r});
  const bool march = vec && !g_axb_legacy_stencils && d.nr >= RBW + 2 && d.nz >= 2 * MTW + 2;
  const dim3 blk(TBX, TBY), grd = grid2d(d);
  const dim3 mg((d.nz + 2 * MTW - 1) / (2 * MTW), (d.nr + RBW - 1) / RBW);
#define BARO(M, PZ, PR, R1, NU)                                                                                        \
  if (march) {                                                                                                         \
    km_baroclinic<M><<<mg, MTW, 0, s>>>(d, w, u_z, u_r, old_u_z, old_u_r, density, PZ, PR, R1, NU, dt);              \
    AXB_LAUNCHED();                                                                                                    \
    km_baroclinic_edge<M><<<mg, MTW, 0, s>>>(d, w, u_z, u_r, old_u_z, old_u_r, density, PZ, PR, R1, NU, dt);         \
  } else {                                                                                                             \
    k_baroclinic<M><<<grd, blk, 0, s>>>(d, w, u_z, u_r, old_u_z, old_u_r, density, PZ, PR, R1, NU, dt);              \
  }
  if (mode == 0) {
    BARO(0, nullptr, nullptr, nullptr, 0.0)
  } else if (mode == 1) {
    BARO(1, penal_z, penal_r, nullptr, 0.0)
  } else {
    BARO(2, penal_z, penal_r, r1d, nu)
  }
#undef BARO
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

}  // extern "C"
