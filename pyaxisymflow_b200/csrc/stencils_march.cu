// stencils_march.cu -- row-marching versions of the four hot stencil passes of the rigid-flow
// step (G-VEL, G-PEN, G-ADV/G-REF, G-DIF).
//
// Thread layout: a block is 128 threads = 256 adjacent z columns (two per thread, 128-bit
// accesses) and marches over RB consecutive rows.  The r-direction neighbourhood lives in a
// rolling register window, so every input row is loaded once per thread (the 2-D tiled
// kernels of stencils.cu / eno3.cu re-load it for each of the 3-5 rows that use it, through
// L1), z-neighbours of a row come from warp shuffles (lane +-1) with the two warp-edge lanes
// falling back to L1 loads.  Per-row reciprocals 1/r are computed once per block into shared
// memory, and FP64 divisions by the constants 2dx, dx^2 and r are multiplications by
// reciprocals (<= 2 ulp from the reference's sequence of divisions; the parity bar is 1e-10).
// In G-PEN each cell is penalised once and its velocity defect (u_pen - u) is kept in the
// rolling window for the three rows that need it.  In G-ADV the r-direction face flux of row
// j is re-used as the back face of row j+1.  Compiled with -fmad=false like stencils.cu.
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include <initializer_list>

#include "axb_common.cuh"
#include "axb_march.cuh"

namespace {

constexpr int MT = 128;  // threads per block: 256 columns

__device__ __forceinline__ const double* rowp(const double* f, long long ld, int j) { return f + (long long)j * ld; }
__device__ __forceinline__ double* rowp(double* f, long long ld, int j) { return f + (long long)j * ld; }

__device__ __forceinline__ double shfl_up_d(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn_d(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// z-neighbours of a register pair: left = f[k-1], right = f[k+2].  Interior lanes take them from
// the adjacent lanes' pairs, the warp-edge lanes (and lanes next to an inactive lane) load them.
// `row` may be null when the caller guarantees interior lanes only.
__device__ __forceinline__ void z_neighbours(const double2 c, const double* __restrict__ row, int k, int nz, int lane,
                                             double& left, double& right) {
  left = shfl_up_d(c.y);
  right = shfl_dn_d(c.x);
  if (lane == 0) left = (k >= 1) ? row[k - 1] : 0.0;
  if (lane == 31 || k + 2 >= nz) right = (k + 2 < nz) ? row[k + 2] : 0.0;
}

struct Cols {
  int k;
  bool own0, own1;     // column k / k+1 is owned (inside [ku0, ku1))
  bool int0, int1;     // ... and has both z neighbours inside the global domain (kg in [1, nzg-2])
  bool valid;          // the pair starts inside the stored row
  int ks;              // column used for stores: k, or an out-of-range column for parked lanes
  int kg;
};
// Threads whose pair lies beyond the stored row stay alive (the kernels use full-mask warp shuffles):
// they are parked on column 0 with every ownership flag false, so they load valid memory and never store.
__device__ __forceinline__ Cols make_cols(const GridD& g, int bx) {
  Cols c;
  c.k = 2 * (bx * MT + threadIdx.x);
  c.valid = c.k < g.nz;
  c.ks = c.valid ? c.k : (g.nz + 2);
  if (!c.valid) c.k = 0;
  c.kg = c.k + g.kz0;
  c.own0 = c.valid && (c.k >= g.ku0) && (c.k < g.ku1);
  c.own1 = c.valid && (c.k + 1 >= g.ku0) && (c.k + 1 < g.ku1);
  c.int0 = c.own0 && (c.kg >= 1) && (c.kg <= g.nzg - 2);
  c.int1 = c.own1 && (c.kg + 1 >= 1) && (c.kg + 1 <= g.nzg - 2);
  return c;
}

// Block-uniform test for the branch-free fast paths: all 256 columns of the block are owned and at
// least `hz` columns away from the global z ends, rows [j0, j1) are a full chunk at least `hr` rows
// away from both r ends, and 128-bit accesses are legal.  The fast and the general code evaluate the
// same floating-point expressions, so a cell gets the same bits whichever path computes it.
// (Separable in the column block and the row chunk; the launchers evaluate the same two predicates on the host to
// enumerate the edge blocks.)
__host__ __device__ inline bool cols_interior(const GridD& g, int bx, int hz) {
  const int kb0 = 2 * bx * MT, kb1 = kb0 + 2 * MT;
  return (kb0 >= g.ku0) && (kb1 <= g.ku1) && (kb0 + g.kz0 >= hz) && (kb1 - 1 + g.kz0 <= g.nzg - 1 - hz) && (kb0 - hz >= 0) &&
         (kb1 - 1 + hz < g.nz);
}
__host__ __device__ inline bool rows_interior(const GridD& g, int p0, int p1, int RB, int hr, bool owned_window) {
  return (p1 - p0 == RB) && (p0 >= hr) && (p1 + hr <= g.nr) && (!owned_window || (p0 >= g.ju0 && p1 <= g.ju1));
}
// Compact enumeration of the edge blocks of one launch (built by edge_map() on the host): the non-interior column blocks
// over all fine row chunks, then the remaining column blocks over the fine chunks whose parent chunk is not
// row-interior.  on = 0: plain 2-D grid (x = column block, y = fine chunk).
struct EdgeMap {
  int on, total, nfy;
  int n_bc, bc[4];
  int n_rr, r0[4], r1[4], rows_r;
  int n_col_part;
};
// Row chunk of this block.  Interior kernels (PATH 1) walk chunks of RB rows.  The edge kernels (PATH 2) walk FINE chunks
// of RE rows (RE divides RB) and take the interior test from the parent RB-chunk: an edge block is latency bound
// (a row's loads are consumed before the next row's are issued, few blocks per SM), so its duration is its row count --
// 32-row edge blocks took 50-100 us whatever the grid, as long as the interior kernel of a small grid or of one rank's
// row slab.  Returns false when the block belongs to the other kernel of the pair.
// launched block (ex, ey) of an edge kernel -> column block bx, fine row chunk fy; false: nothing to do
__host__ __device__ inline bool edge_decode(const EdgeMap& em, int ex, int ey, int& bx, int& fy) {
  if (!em.on) { bx = ex; fy = ey; return true; }
  int e = ex;
  if (e >= em.total) return false;
  if (e < em.n_col_part) { bx = em.bc[e / em.nfy]; fy = e % em.nfy; return true; }
  e -= em.n_col_part;
  bx = e / em.rows_r;
  int ri = e % em.rows_r;
  for (int i = 0; i < em.n_bc; ++i)
    if (em.bc[i] <= bx) ++bx;
  fy = 0;
  for (int i = 0; i < em.n_rr; ++i) {
    const int len = em.r1[i] - em.r0[i];
    if (ri < len) { fy = em.r0[i] + ri; break; }
    ri -= len;
  }
  return true;
}
template <int PATH>
__device__ __forceinline__ bool row_chunk(const GridD& g, int RB, int RE, const EdgeMap& em, int hr, int hz, bool vec,
                                          bool owned_window, int& bx, int& j0, int& j1) {
  int p0, p1;
  if (PATH == 1) {
    bx = blockIdx.x;
    j0 = blockIdx.y * RB; j1 = min(j0 + RB, g.nr);
    p0 = j0; p1 = j1;
  } else {
    int fy;
    if (!edge_decode(em, (int)blockIdx.x, (int)blockIdx.y, bx, fy)) return false;
    j0 = fy * RE; j1 = min(j0 + RE, g.nr);
    p0 = (j0 / RB) * RB; p1 = min(p0 + RB, g.nr);
  }
  const bool interior = vec && rows_interior(g, p0, p1, RB, hr, owned_window) && cols_interior(g, bx, hz);
  return (PATH == 1) == interior;
}
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
constexpr int UR = 4;   // rows per unrolled step of the fast paths (RB is a multiple of it)

// -------------------------------------------------------------------------------------
// G-VEL
// -------------------------------------------------------------------------------------
// PATH: 1 = branch-free interior blocks only, 2 = the general (edge) blocks only.  Every operation is
// launched as the pair <1>, <2>: the two kernels get their own register allocation (the interior one
// is the hot one), blocks that belong to the other kernel leave at once.
template <bool REDUCE, int PATH>
__global__ void __launch_bounds__(MT)
    km_velocity(GridD g, int RB, int RE, EdgeMap em, double* __restrict__ u_z, double* __restrict__ u_r, const double* __restrict__ psi,
                const double* __restrict__ r1d, double uz_add, double ur_add, const double* __restrict__ add_dev,
                double* umax_out, bool vec) {
  extern __shared__ double s_inv[];
  int bx, j0, j1;
  // a block that straddles the owned-row window [ju0, ju1) reduces row by row on the general path
  if (!row_chunk<PATH>(g, RB, RE, em, 1, 1, vec, REDUCE, bx, j0, j1)) return;
  for (int i = threadIdx.x; i < j1 - j0; i += MT) s_inv[i] = 1.0 / r1d[j0 + i];
  __syncthreads();
  const Cols c = make_cols(g, bx);
  const int lane = threadIdx.x & 31, nz = g.nz, k = c.k;
  double local_max = 0.0;
  {
    const long long fo = member_field(g);
    u_z += fo; u_r += fo; psi += fo;
    add_dev = moved(add_dev, member_scalar(g));
    umax_out = moved(umax_out, member_scalar(g));
  }
  if (add_dev) { uz_add = add_dev[0]; ur_add = add_dev[1]; }
  if (PATH == 1) {
    const double inv_h = 1.0 / (2 * g.dx);
    const long long ld = g.ld;
    const double* p = psi + (long long)(j0 - 1) * ld + k;
    double2 pm = ld2(p), pc = ld2(p + ld);
    p += 2 * ld;                                   // row j + 1
    double* oz = u_z + (long long)j0 * ld + k;
    double* orr = u_r + (long long)j0 * ld + k;
    for (int jb = 0; jb < RB; jb += UR) {
      double2 pn[UR];
      double le[UR], re[UR];
#pragma unroll
      for (int u = 0; u < UR; ++u) pn[u] = ld2(p + u * ld);
#pragma unroll
      for (int u = 0; u < UR; ++u) {                 // warp-edge z neighbours of rows j .. j+3
        const double* rr = p + (u - 1) * ld;
        le[u] = (lane == 0) ? rr[-1] : 0.0;
        re[u] = (lane == 31) ? rr[2] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const double ir = s_inv[jb + u];
        double left = shfl_up_d(pc.y), right = shfl_dn_d(pc.x);
        if (lane == 0) left = le[u];
        if (lane == 31) right = re[u];
        double2 uz, ur;
        uz.x = (pn[u].x - pm.x) * inv_h * ir;
        uz.y = (pn[u].y - pm.y) * inv_h * ir;
        ur.x = -(pc.y - left) * inv_h * ir;
        ur.y = -(right - pc.x) * inv_h * ir;
        uz.x += uz_add; uz.y += uz_add;
        ur.x += ur_add; ur.y += ur_add;
        st2(oz + u * ld, uz);
        st2(orr + u * ld, ur);
        if (REDUCE) local_max = fmax(local_max, fmax(fabs(uz.x) + fabs(ur.x), fabs(uz.y) + fabs(ur.y)));
        pm = pc; pc = pn[u];
      }
      p += UR * ld; oz += UR * ld; orr += UR * ld;
    }
  } else if (PATH == 2) {
    const double inv_h = 1.0 / (2 * g.dx);
    double2 pm = make_double2(0, 0), pc = ld_pair(rowp(psi, g.ld, j0), k, nz, vec), pp = pc;
    if (j0 > 0) pm = ld_pair(rowp(psi, g.ld, j0 - 1), k, nz, vec);
    for (int j = j0; j < j1; ++j) {
      const double* prow = rowp(psi, g.ld, j);
      if (j + 1 < g.nr) pp = ld_pair(rowp(psi, g.ld, j + 1), k, nz, vec);
      const double ir = s_inv[j - j0];
      double2 uz;
      if (j > 0 && j < g.nr - 1) {
        uz.x = (pp.x - pm.x) * inv_h * ir;
        uz.y = (pp.y - pm.y) * inv_h * ir;
      } else if (j == 0) {
        const double2 p2 = ld_pair(rowp(psi, g.ld, 2), k, nz, vec);
        uz.x = (-p2.x + 4 * pp.x - 3 * pc.x) * inv_h * ir;
        uz.y = (-p2.y + 4 * pp.y - 3 * pc.y) * inv_h * ir;
      } else {
        const double2 p2 = ld_pair(rowp(psi, g.ld, j - 2), k, nz, vec);
        uz.x = (p2.x - 4 * pm.x + 3 * pc.x) * inv_h * ir;
        uz.y = (p2.y - 4 * pm.y + 3 * pc.y) * inv_h * ir;
      }
      double left, right;
      z_neighbours(pc, prow, k, nz, lane, left, right);
      double2 ur = make_double2(0, 0);
      if (c.int0) ur.x = -(pc.y - left) * inv_h * ir;
      else if (c.own0 && c.kg == 0) ur.x = -(-right + 4 * pc.y - 3 * pc.x) * inv_h * ir;
      else if (c.own0) ur.x = -(prow[k - 2] - 4 * left + 3 * pc.x) * inv_h * ir;         // kg == nzg-1
      if (c.int1) ur.y = -(right - pc.x) * inv_h * ir;
      else if (c.own1 && c.kg + 1 == g.nzg - 1) ur.y = -(left - 4 * pc.x + 3 * pc.y) * inv_h * ir;
      else if (c.own1) ur.y = -(-prow[k + 3] + 4 * right - 3 * pc.y) * inv_h * ir;        // kg+1 == 0
      uz.x += uz_add; uz.y += uz_add;
      ur.x += ur_add; ur.y += ur_add;
      st_pair(rowp(u_z, g.ld, j), c.ks, g.ku0, g.ku1, vec, uz);
      st_pair(rowp(u_r, g.ld, j), c.ks, g.ku0, g.ku1, vec, ur);
      if (REDUCE && j >= g.ju0 && j < g.ju1) {
        if (c.own0) local_max = fmax(local_max, fabs(uz.x) + fabs(ur.x));
        if (c.own1) local_max = fmax(local_max, fabs(uz.y) + fabs(ur.y));
      }
      pm = pc; pc = pp;
    }
  }
  if (REDUCE) {
    const double m = block_max(local_max);
    if (threadIdx.x == 0 && m > 0.0) atomic_max_nonneg(umax_out, m);
  }
}

// -------------------------------------------------------------------------------------
// G-PEN
// -------------------------------------------------------------------------------------
struct PenRow {      // one penalised row pair
  double2 dz;        // pen(u_z) - u_z
  double2 pz, pr;    // penalised components
  double2 dr;        // pen(u_r) - u_r
  double2 chi;
};
__device__ __forceinline__ double pen_val(double u, double lamdt_chi, double U, double inv_den) {
  return (u + lamdt_chi * U) * inv_den;
}
__device__ __forceinline__ PenRow pen_row(const double* __restrict__ uzu, const double* __restrict__ uru,
                                         const double* __restrict__ chi, long long ld, int j, int k, int nz, bool vec,
                                         double lamdt, double U_z, double U_r) {
  PenRow r;
  const double2 c = ld_pair(rowp(chi, ld, j), k, nz, vec);
  const double2 z = ld_pair(rowp(uzu, ld, j), k, nz, vec);
  const double2 q = ld_pair(rowp(uru, ld, j), k, nz, vec);
  const double lx = lamdt * c.x, ly = lamdt * c.y;
  const double ix = 1.0 / (1 + lx), iy = 1.0 / (1 + ly);
  r.pz = make_double2(pen_val(z.x, lx, U_z, ix), pen_val(z.y, ly, U_z, iy));
  r.pr = make_double2(pen_val(q.x, lx, U_r, ix), pen_val(q.y, ly, U_r, iy));
  r.dz = make_double2(r.pz.x - z.x, r.pz.y - z.y);
  r.dr = make_double2(r.pr.x - q.x, r.pr.y - q.y);
  r.chi = c;
  return r;
}

__device__ __forceinline__ PenRow pen_from(double2 c, double2 z, double2 q, double lamdt, double U_z, double U_r) {
  PenRow r;
  const double lx = lamdt * c.x, ly = lamdt * c.y;
  const double ix = 1.0 / (1 + lx), iy = 1.0 / (1 + ly);
  r.pz = make_double2(pen_val(z.x, lx, U_z, ix), pen_val(z.y, ly, U_z, iy));
  r.pr = make_double2(pen_val(q.x, lx, U_r, ix), pen_val(q.y, ly, U_r, iy));
  r.dz = make_double2(r.pz.x - z.x, r.pz.y - z.y);
  r.dr = make_double2(r.pr.x - q.x, r.pr.y - q.y);
  r.chi = c;
  return r;
}
__device__ __forceinline__ double pen_defect(double cc, double uu, double lamdt, double U) {
  const double l = lamdt * cc;
  return pen_val(uu, l, U, 1.0 / (1 + l)) - uu;
}

// (interior kernel capped at 128 registers = 4 blocks per SM: with the fused reduction it took 162 and ran at 0.87 of HBM
// where the 128-register form without it runs at 0.92; ptxas spills 76 bytes.  Measured with the same cap on the one-pass
// RK2 -- 168 -> 128 registers, 4 bytes spilled --: C4 2.850 -> 2.828 / 2.831 with either, 2.806 with both; the same
// treatment of ENO3 and G-VEL, 94 -> 80 registers for 6 blocks per SM, LOST 2.5 % on C4 and 7 % on C3 and is not used; the
// form without the reduction (150 registers, config C3) measured the same capped or not, 1.832 vs 1.834 ms, and stays free.)
template <bool REDUCE, int PATH>
__global__ void __launch_bounds__(MT, (PATH == 1 && REDUCE) ? 4 : 1)
    km_penalise(GridD g, int RB, int RE, EdgeMap em, double* __restrict__ u_z, double* __restrict__ u_r, double* __restrict__ w,
                const double* __restrict__ uzu, const double* __restrict__ uru, const double* __restrict__ chi,
                double lam, double dt, const double* __restrict__ dt_dev, double U_z, double U_r,
                const double* __restrict__ U_dev, const double* __restrict__ r1d, double* sum_out, bool vec) {
  int bx, j0, j1;
  if (!row_chunk<PATH>(g, RB, RE, em, 1, 1, vec, REDUCE, bx, j0, j1)) return;
  const Cols c = make_cols(g, bx);
  const int lane = threadIdx.x & 31, nz = g.nz, k = c.k;
  double local = 0.0;
  {
    const long long fo = member_field(g);
    u_z += fo; u_r += fo; w += fo; uzu += fo; uru += fo; chi += fo;
    dt_dev = moved(dt_dev, member_scalar(g));
    U_dev = moved(U_dev, member_scalar(g));
    sum_out = moved(sum_out, member_scalar(g));
  }
  if (dt_dev) dt = *dt_dev;
  if (U_dev) { U_z = U_dev[0]; U_r = U_dev[1]; }
  if (PATH == 1) {
    const double lamdt = lam * dt;
    const double inv_h = 1.0 / (2 * g.dx);
    const long long ld = g.ld;
    const long long off = (long long)(j0 - 1) * ld + k;
    const double *pc_ = chi + off, *pz_ = uzu + off, *pr_ = uru + off;
    double2 dz_m = pen_from(ld2(pc_), ld2(pz_), ld2(pr_), lamdt, U_z, U_r).dz;
    PenRow cur = pen_from(ld2(pc_ + ld), ld2(pz_ + ld), ld2(pr_ + ld), lamdt, U_z, U_r);
    pc_ += 2 * ld; pz_ += 2 * ld; pr_ += 2 * ld;           // row j + 1
    double* oz = u_z + off + ld;
    double* orr = u_r + off + ld;
    double* wr = w + off + ld;
    for (int jb = 0; jb < RB; jb += 2) {
      // all loads of two rows first: rows j+1, j+2 of (chi, u_z, u_r), rows j, j+1 of w, warp-edge columns
      const double2 c1 = ld2(pc_), z1 = ld2(pz_), q1 = ld2(pr_);
      const double2 c2 = ld2(pc_ + ld), z2 = ld2(pz_ + ld), q2 = ld2(pr_ + ld);
      double2 w0 = ld2(wr), w1 = ld2(wr + ld);
      double ec[2] = {0, 0}, eu[2] = {0, 0};
      if (lane == 0) {
        ec[0] = pc_[-ld - 1]; eu[0] = pr_[-ld - 1]; ec[1] = pc_[-1]; eu[1] = pr_[-1];
      } else if (lane == 31) {
        ec[0] = pc_[-ld + 2]; eu[0] = pr_[-ld + 2]; ec[1] = pc_[2]; eu[1] = pr_[2];
      }
      const PenRow n1 = pen_from(c1, z1, q1, lamdt, U_z, U_r), n2 = pen_from(c2, z2, q2, lamdt, U_z, U_r);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const PenRow& nx = u ? n2 : n1;
        st2(oz + u * ld, cur.pz);
        st2(orr + u * ld, cur.pr);
        if (REDUCE) {
          const double r = r1d[j0 + jb + u];
          local += r * cur.chi.x * (cur.pz.x - U_z);
          local += r * cur.chi.y * (cur.pz.y - U_z);
        }
        double dl = shfl_up_d(cur.dr.y), dr = shfl_dn_d(cur.dr.x);
        if (lane == 0) dl = pen_defect(ec[u], eu[u], lamdt, U_r);
        if (lane == 31) dr = pen_defect(ec[u], eu[u], lamdt, U_r);
        double2 wv = u ? w1 : w0;
        const double dzx = nx.dz.x - dz_m.x, dzy = nx.dz.y - dz_m.y;
        wv.x = wv.x + ((cur.dr.y - dl) * inv_h - dzx * inv_h);
        wv.y = wv.y + ((dr - cur.dr.x) * inv_h - dzy * inv_h);
        st2(wr + u * ld, wv);
        dz_m = cur.dz;
        cur = nx;
      }
      pc_ += 2 * ld; pz_ += 2 * ld; pr_ += 2 * ld; oz += 2 * ld; orr += 2 * ld; wr += 2 * ld;
    }
  } else if (PATH == 2) {
    const double lamdt = lam * dt;
    const double inv_h = 1.0 / (2 * g.dx);
    PenRow cur = pen_row(uzu, uru, chi, g.ld, j0, k, nz, vec, lamdt, U_z, U_r), nxt = cur;
    double2 dz_m = make_double2(0, 0);
    if (j0 > 0) dz_m = pen_row(uzu, uru, chi, g.ld, j0 - 1, k, nz, vec, lamdt, U_z, U_r).dz;
    for (int j = j0; j < j1; ++j) {
      if (j + 1 < g.nr) nxt = pen_row(uzu, uru, chi, g.ld, j + 1, k, nz, vec, lamdt, U_z, U_r);
      st_pair(rowp(u_z, g.ld, j), c.ks, g.ku0, g.ku1, vec, cur.pz);
      st_pair(rowp(u_r, g.ld, j), c.ks, g.ku0, g.ku1, vec, cur.pr);
      if (REDUCE && j >= g.ju0 && j < g.ju1) {
        const double r = r1d[j];
        if (c.own0) local += r * cur.chi.x * (cur.pz.x - U_z);
        if (c.own1) local += r * cur.chi.y * (cur.pz.y - U_z);
      }
      // z neighbours of the u_r defect: lanes +-1, warp-edge lanes recompute from memory
      double dl = shfl_up_d(cur.dr.y), dr = shfl_dn_d(cur.dr.x);
      if (j >= 1 && j < g.nr - 1) {
        if (lane == 0 && c.int0) {
          const double cc = rowp(chi, g.ld, j)[k - 1], uu = rowp(uru, g.ld, j)[k - 1];
          const double l = lamdt * cc;
          dl = pen_val(uu, l, U_r, 1.0 / (1 + l)) - uu;
        }
        if ((lane == 31 || k + 2 >= nz) && c.int1) {
          const double cc = rowp(chi, g.ld, j)[k + 2], uu = rowp(uru, g.ld, j)[k + 2];
          const double l = lamdt * cc;
          dr = pen_val(uu, l, U_r, 1.0 / (1 + l)) - uu;
        }
        if (c.int0 || c.int1) {
          double* wr = rowp(w, g.ld, j);
          double2 wv = ld_pair(wr, k, nz, vec);
          const double dzx = nxt.dz.x - dz_m.x, dzy = nxt.dz.y - dz_m.y;
          if (c.int0) wv.x = wv.x + ((cur.dr.y - dl) * inv_h - dzx * inv_h);
          if (c.int1) wv.y = wv.y + ((dr - cur.dr.x) * inv_h - dzy * inv_h);
          st_pair(wr, k, c.int0 ? k : k + 1, c.int1 ? k + 2 : k + 1, vec, wv);
        }
      }
      dz_m = cur.dz;
      cur = nxt;
    }
  }
  if (REDUCE) {
    const double s = block_sum(local);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(sum_out, s);
  }
}

// -------------------------------------------------------------------------------------
// G-DIF
// -------------------------------------------------------------------------------------
template <int STAGE, int PATH>
__global__ void __launch_bounds__(MT)
    km_diffusion(GridD g, int RB, int RE, EdgeMap em, double* out, const double* __restrict__ in, const double* src2,
                 const double* __restrict__ r1d, double nu, const double* __restrict__ nu_dev, double dt,
                 const double* __restrict__ dt_dev, bool vec) {
  extern __shared__ double s_inv[];
  int bx, j0, j1;
  if (!row_chunk<PATH>(g, RB, RE, em, 1, 1, vec, false, bx, j0, j1)) return;
  for (int i = threadIdx.x; i < j1 - j0; i += MT) s_inv[i] = 1.0 / r1d[j0 + i];
  __syncthreads();
  const Cols c = make_cols(g, bx);
  const int lane = threadIdx.x & 31, nz = g.nz, k = c.k;
  {
    const long long fo = member_field(g);
    out += fo; in += fo; src2 = moved(src2, fo);
  }
  if (dt_dev) dt = dt_dev[member_scalar(g)];
  if (nu_dev) nu = nu_dev[member_scalar(g)];
  const double coef = (STAGE == 1) ? (0.5 * nu * dt) : (nu * dt);
  const double inv_dx2 = 1.0 / (g.dx * g.dx), inv_h = 1.0 / (2 * g.dx);
  if (PATH == 1) {
    const long long ld = g.ld;
    const double* p = in + (long long)(j0 - 1) * ld + k;
    double2 pm = ld2(p), pc = ld2(p + ld);
    p += 2 * ld;
    double* o = out + (long long)j0 * ld + k;
    const double* s2 = (STAGE == 1) ? nullptr : src2 + (long long)j0 * ld + k;
    for (int jb = 0; jb < RB; jb += UR) {
      double2 pn[UR], sv[UR];
      double le[UR], re[UR];
#pragma unroll
      for (int u = 0; u < UR; ++u) pn[u] = ld2(p + u * ld);
      if (STAGE == 2) {
#pragma unroll
        for (int u = 0; u < UR; ++u) sv[u] = ld2(s2 + u * ld);
      }
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const double* rr = p + (u - 1) * ld;
        le[u] = (lane == 0) ? rr[-1] : 0.0;
        re[u] = (lane == 31) ? rr[2] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const double ir = s_inv[jb + u];
        const double ir2 = ir * ir;
        double left = shfl_up_d(pc.y), right = shfl_dn_d(pc.x);
        if (lane == 0) left = le[u];
        if (lane == 31) right = re[u];
        double2 res = (STAGE == 1) ? pc : sv[u];
        const double2 pp = pn[u];
        res.x += coef * ((pp.x + pm.x + pc.y + left - 4 * pc.x) * inv_dx2 + (pp.x - pm.x) * inv_h * ir - pc.x * ir2);
        res.y += coef * ((pp.y + pm.y + right + pc.x - 4 * pc.y) * inv_dx2 + (pp.y - pm.y) * inv_h * ir - pc.y * ir2);
        st2(o + u * ld, res);
        pm = pc; pc = pp;
      }
      p += UR * ld; o += UR * ld;
      if (STAGE == 2) s2 += UR * ld;
    }
    return;
  }
  if (PATH != 2) return;
  double2 pm = make_double2(0, 0), pc = ld_pair(rowp(in, g.ld, j0), k, nz, vec), pp = pc;
  if (j0 > 0) pm = ld_pair(rowp(in, g.ld, j0 - 1), k, nz, vec);
  for (int j = j0; j < j1; ++j) {
    const double* irow = rowp(in, g.ld, j);
    if (j + 1 < g.nr) pp = ld_pair(rowp(in, g.ld, j + 1), k, nz, vec);
    double2 res;
    if (STAGE == 1) res = pc;
    else res = ld_pair(rowp(src2, g.ld, j), k, nz, vec);
    double left, right;
    z_neighbours(pc, irow, k, nz, lane, left, right);
    if (j >= 1 && j < g.nr - 1) {
      const double ir = s_inv[j - j0];
      const double ir2 = ir * ir;
      if (c.int0)
        res.x += coef * ((pp.x + pm.x + pc.y + left - 4 * pc.x) * inv_dx2 + (pp.x - pm.x) * inv_h * ir - pc.x * ir2);
      if (c.int1)
        res.y += coef * ((pp.y + pm.y + right + pc.x - 4 * pc.y) * inv_dx2 + (pp.y - pm.y) * inv_h * ir - pc.y * ir2);
    }
    st_pair(rowp(out, g.ld, j), c.ks, g.ku0, g.ku1, vec, res);
    pm = pc; pc = pp;
  }
}

// -------------------------------------------------------------------------------------
// G-DIF, both RK2 stages in one pass (kernels/diffusion_RK2.py:4-45): out = w + nu dt L(tmp),
// tmp = w + nu dt / 2 L(w) on interior cells and tmp = w elsewhere.  tmp never goes to memory: a
// block keeps a rolling window of three w rows and three tmp rows; the z-neighbours of tmp come from the
// adjacent lanes, the two warp-edge lanes evaluate tmp on their ghost column themselves (one more
// rolling w column + one load per row).  Same expressions, evaluated in the same order, as the two
// km_diffusion stages, so the result has the same bits.  16 B/pt instead of 40.
// -------------------------------------------------------------------------------------
struct DifK {
  double c1, c2, inv_dx2, inv_h;
};
// w + c * L(w) at one cell: up / down = r-neighbours, zp / zm = the z-neighbours k+1 / k-1
__device__ __forceinline__ double dif_cell(double base, double c, double up, double dn, double zp, double zm, double ctr,
                                           double ir, double ir2, const DifK& K) {
  return base + c * ((up + dn + zp + zm - 4 * ctr) * K.inv_dx2 + (up - dn) * K.inv_h * ir - ctr * ir2);
}

template <int PATH>
__global__ void __launch_bounds__(MT, PATH == 1 ? 4 : 1)
    km_diffusion_fused(GridD g, int RB, int RE, EdgeMap em, double* __restrict__ out, const double* __restrict__ in,
                       const double* __restrict__ r1d, double nu, double dt, const double* __restrict__ dt_dev, bool vec) {
  extern __shared__ double s_inv[];          // 1 / r for rows j0 - 1 .. j1
  int bx, j0, j1;
  if (!row_chunk<PATH>(g, RB, RE, em, 2, 2, vec, false, bx, j0, j1)) return;
  for (int i = threadIdx.x; i < j1 - j0 + 2; i += MT) {
    const int j = j0 - 1 + i;
    s_inv[i] = (j >= 0 && j < g.nr) ? 1.0 / r1d[j] : 0.0;
  }
  __syncthreads();
  const Cols c = make_cols(g, bx);
  const int lane = threadIdx.x & 31, k = c.k;
  if (dt_dev) dt = *dt_dev;
  DifK K;
  K.c1 = 0.5 * nu * dt;
  K.c2 = nu * dt;
  K.inv_dx2 = 1.0 / (g.dx * g.dx);
  K.inv_h = 1.0 / (2 * g.dx);
  if (PATH == 1) {
    const long long ld = g.ld;
    // ghost column of the warp-edge lanes: k - 1 (lane 0) or k + 2 (lane 31); gz = its outer z-neighbour
    const bool edge = (lane == 0) || (lane == 31);
    const int kg = (lane == 0) ? k - 1 : k + 2;
    const int go = (lane == 0) ? k - 2 : k + 3;
    // tmp row t from w rows t-1, t, t+1 (wm, wc, wp); returns the pair and the ghost-column value
    auto tmp_row = [&](int t, const double2& wm, const double2& wc, const double2& wp, double gm, double gc, double gp,
                       double gout, double2& tv, double& tg) {
      const double ir = s_inv[t - j0 + 1], ir2 = ir * ir;
      double left = shfl_up_d(wc.y), right = shfl_dn_d(wc.x);
      if (lane == 0) left = gc;
      if (lane == 31) right = gc;
      tv.x = dif_cell(wc.x, K.c1, wp.x, wm.x, wc.y, left, wc.x, ir, ir2, K);
      tv.y = dif_cell(wc.y, K.c1, wp.y, wm.y, right, wc.x, wc.y, ir, ir2, K);
      // ghost column: its z-neighbours are (own pair element, gout) in k+1 / k-1 order
      tg = (lane == 0) ? dif_cell(gc, K.c1, gp, gm, wc.x, gout, gc, ir, ir2, K)
                       : dif_cell(gc, K.c1, gp, gm, gout, wc.y, gc, ir, ir2, K);
    };
    const double* p = in + (long long)(j0 - 2) * ld;
    double2 w0 = ld2(p + k), w1 = ld2(p + ld + k), w2 = ld2(p + 2 * ld + k), w3 = ld2(p + 3 * ld + k);   // rows j0-2 .. j0+1
    double g0 = 0, g1 = 0, g2 = 0, g3 = 0, o1 = 0, o2 = 0;
    if (edge) {
      g0 = p[kg]; g1 = p[ld + kg]; g2 = p[2 * ld + kg]; g3 = p[3 * ld + kg];
      o1 = p[ld + go]; o2 = p[2 * ld + go];
    }
    double2 tm, tc;
    double tgm, tgc;
    tmp_row(j0 - 1, w0, w1, w2, g0, g1, g2, o1, tm, tgm);
    tmp_row(j0, w1, w2, w3, g1, g2, g3, o2, tc, tgc);
    // window now: wa = w[j-1] (unused), wb = w[j], wc_ = w[j+1]
    double2 wb = w2, wcn = w3;
    double gb = g2, gcn = g3;
    p += 4 * ld;                                 // row j + 2
    double* o = out + (long long)j0 * ld + k;
    (void)tgm;
    // software pipeline: the loads of the next UR rows are issued before the current UR rows are computed
    // (this pass has a single input field, so one batch of loads in flight does not cover the HBM latency)
    double2 pn[UR], pq[UR];
    double gn[UR], on[UR], gq[UR], oq[UR];
    auto fetch = [&](double2 (&a)[UR], double (&gg)[UR], double (&oo)[UR]) {
#pragma unroll
      for (int u = 0; u < UR; ++u) a[u] = ld2(p + u * ld + k);
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        gg[u] = edge ? p[u * ld + kg] : 0.0;                 // w[j+2][ghost]
        oo[u] = edge ? p[(u - 1) * ld + go] : 0.0;           // w[j+1][outer]
      }
    };
    fetch(pn, gn, on);
    for (int jb = 0; jb < RB; jb += UR) {
      p += UR * ld;
      if (jb + UR < RB) fetch(pq, gq, oq);
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const int j = j0 + jb + u;
        double2 tn;
        double tgn;
        tmp_row(j + 1, wb, wcn, pn[u], gb, gcn, gn[u], on[u], tn, tgn);
        const double ir = s_inv[j - j0 + 1], ir2 = ir * ir;
        double left = shfl_up_d(tc.y), right = shfl_dn_d(tc.x);
        if (lane == 0) left = tgc;
        if (lane == 31) right = tgc;
        double2 res;
        res.x = dif_cell(wb.x, K.c2, tn.x, tm.x, tc.y, left, tc.x, ir, ir2, K);
        res.y = dif_cell(wb.y, K.c2, tn.y, tm.y, right, tc.x, tc.y, ir, ir2, K);
        st2(o + u * ld, res);
        tm = tc; tc = tn; tgc = tgn;
        wb = wcn; wcn = pn[u];
        gb = gcn; gcn = gn[u];
      }
#pragma unroll
      for (int u = 0; u < UR; ++u) { pn[u] = pq[u]; gn[u] = gq[u]; on[u] = oq[u]; }
      o += UR * ld;
    }
    return;
  }
  if (PATH != 2) return;
  // general blocks: every cell from memory, boundary rules applied cell by cell
  const int nz = g.nz;
  auto cell_interior = [&](int j, int kk) {
    const int kgl = kk + g.kz0;
    return j >= 1 && j < g.nr - 1 && kgl >= 1 && kgl <= g.nzg - 2 && kk >= 1 && kk + 1 < nz;
  };
  auto tmp_at = [&](int j, int kk) -> double {
    const double* r0 = rowp(in, g.ld, j);
    const double v = r0[kk];
    if (!cell_interior(j, kk)) return v;
    const double ir = 1.0 / r1d[j], ir2 = ir * ir;
    return dif_cell(v, K.c1, rowp(in, g.ld, j + 1)[kk], rowp(in, g.ld, j - 1)[kk], r0[kk + 1], r0[kk - 1], v, ir, ir2, K);
  };
  for (int j = j0; j < j1; ++j) {
    const double ir = s_inv[j - j0 + 1], ir2 = ir * ir;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kk = k + e;
      const bool own = e ? c.own1 : c.own0;
      if (!own) continue;
      const double w = rowp(in, g.ld, j)[kk];
      double res = w;
      if (cell_interior(j, kk)) {
        const double tcv = tmp_at(j, kk);
        res = dif_cell(w, K.c2, tmp_at(j + 1, kk), tmp_at(j - 1, kk), tmp_at(j, kk + 1), tmp_at(j, kk - 1), tcv, ir, ir2, K);
      }
      rowp(out, g.ld, j)[kk] = res;
    }
  }
}

// -------------------------------------------------------------------------------------
// G-ADV / G-REF (and the plain-array pystencils forms)
// -------------------------------------------------------------------------------------
#define C13 (1.0 / 3.0)
#define C56 (5.0 / 6.0)
#define C16 (1.0 / 6.0)
__device__ __forceinline__ double eno_face(double qm1, double q0, double qp1, double qp2, double v0, double vp1) {
  const bool up = v0 > -vp1;
  const double a = up ? qp1 : q0, b = up ? q0 : qp1, cc = up ? qm1 : qp2;
  return C13 * a + C56 * b - C16 * cc;
}
template <bool MIRROR>
__device__ __forceinline__ double2 ld_row_m(const double* f, long long ld, int jj, int k, int nz, bool vec, double sign) {
  if (MIRROR && jj < 0) {
    const double2 v = ld_pair(rowp(f, ld, -jj - 1), k, nz, vec);
    return make_double2(sign * v.x, sign * v.y);
  }
  return ld_pair(rowp(f, ld, jj), k, nz, vec);
}

// NF fields share the velocity.  Row window per field: samples q[0..3] = rows j-1 .. j+2 (products
// with u_r when CONS), centre values w0 = f[j], w1 = f[j+1], w2 = f[j+2]; per-thread rolling back face.
template <int NF, bool CONS, bool MIRROR, bool FLUXONLY, int PATH>
__global__ void __launch_bounds__(MT)
    km_eno3(GridD g, int RB, int RE, EdgeMap em, double* __restrict__ out0, double* __restrict__ out1, const double* __restrict__ in0,
            const double* __restrict__ in1, const double* __restrict__ u_z, const double* __restrict__ u_r,
            double inv_dx, double dt, const double* __restrict__ dt_dev, double sign0, double sign1, bool vec) {
  const int j_lo = MIRROR ? 0 : 2, j_hi = g.nr - 3;   // rows that are advected
  int bx, j0, j1;
  if (!row_chunk<PATH>(g, RB, RE, em, 2, 2, vec, false, bx, j0, j1)) return;
  const Cols c = make_cols(g, bx);
  const int nz = g.nz, k = c.k;
  if (!c.own0 && !c.own1) return;   // no shuffles in this kernel: idle threads may leave
  if (dt_dev) dt = *dt_dev;
  if (!FLUXONLY) inv_dx = -(dt / g.dx);
  const bool ok0c = c.own0 && (c.kg >= 2) && (c.kg <= g.nzg - 3);
  const bool ok1c = c.own1 && (c.kg + 1 >= 2) && (c.kg + 1 <= g.nzg - 3);
  double* outs[2] = {out0, out1};
  const double* ins[2] = {in0, in1};
  const double signs[2] = {sign0, sign1};
  const int jmin = MIRROR ? -2 : 0;

  if (PATH == 1) {
    // branch-free interior path: every load of a row is issued before any of it is consumed
    const long long ld = g.ld;
    const long long off = (long long)j0 * ld + k;
    const double* pv = u_r + off - 2 * ld;              // row j - 2, advanced to j + 2 below
    const double* pu = u_z + off;                       // row j
    const double* pw[2] = {in0 + off - 2 * ld, (NF > 1 ? in1 : in0) + off - 2 * ld};
    double* po[2] = {out0 + off, (NF > 1 ? out1 : out0) + off};
    double2 v[4], qq[2][4], wcen[2][3], fb[2];
    {
      const double2 vm2 = ld2(pv);
      v[0] = ld2(pv + ld); v[1] = ld2(pv + 2 * ld); v[2] = ld2(pv + 3 * ld);
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const double2 a = ld2(pw[f]), b = ld2(pw[f] + ld), c2 = ld2(pw[f] + 2 * ld), d2 = ld2(pw[f] + 3 * ld);
        const double2 qm2 = CONS ? make_double2(a.x * vm2.x, a.y * vm2.y) : a;
        qq[f][0] = CONS ? make_double2(b.x * v[0].x, b.y * v[0].y) : b;
        qq[f][1] = CONS ? make_double2(c2.x * v[1].x, c2.y * v[1].y) : c2;
        qq[f][2] = CONS ? make_double2(d2.x * v[2].x, d2.y * v[2].y) : d2;
        wcen[f][0] = c2; wcen[f][1] = d2;
        fb[f].x = eno_face(qm2.x, qq[f][0].x, qq[f][1].x, qq[f][2].x, v[0].x, v[1].x);
        fb[f].y = eno_face(qm2.y, qq[f][0].y, qq[f][1].y, qq[f][2].y, v[0].y, v[1].y);
      }
    }
    pv += 4 * ld;                                        // row j + 2
    pw[0] += 4 * ld; pw[1] += 4 * ld;
    for (int jb = 0; jb < RB; ++jb) {
      // ---- loads
      v[3] = ld2(pv);
      const double2 uzc = ld2(pu), uzl = ld2(pu - 2), uzr = ld2(pu + 2);
      double2 w3[2], wl[2], wrr[2], oo[2];
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        w3[f] = ld2(pw[f]);
        wl[f] = ld2(pw[f] - 2 * ld - 2);
        wrr[f] = ld2(pw[f] - 2 * ld + 2);
        if (FLUXONLY) oo[f] = ld2(po[f]);
      }
      const double vz[6] = {uzl.x, uzl.y, uzc.x, uzc.y, uzr.x, uzr.y};
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        qq[f][3] = CONS ? make_double2(w3[f].x * v[3].x, w3[f].y * v[3].y) : w3[f];
        wcen[f][2] = w3[f];
        double2 Ff;
        Ff.x = eno_face(qq[f][0].x, qq[f][1].x, qq[f][2].x, qq[f][3].x, v[1].x, v[2].x);
        Ff.y = eno_face(qq[f][0].y, qq[f][1].y, qq[f][2].y, qq[f][3].y, v[1].y, v[2].y);
        const double2 cen = wcen[f][0];
        double qz[6] = {wl[f].x, wl[f].y, cen.x, cen.y, wrr[f].x, wrr[f].y};
        if (CONS) {
#pragma unroll
          for (int i = 0; i < 6; ++i) qz[i] *= vz[i];
        }
        const double Fz0 = eno_face(qz[0], qz[1], qz[2], qz[3], vz[1], vz[2]);
        const double Fz1 = eno_face(qz[1], qz[2], qz[3], qz[4], vz[2], vz[3]);
        const double Fz2 = eno_face(qz[2], qz[3], qz[4], qz[5], vz[3], vz[4]);
        double a0 = FLUXONLY ? oo[f].x : 0.0, a1 = FLUXONLY ? oo[f].y : 0.0;
        if (CONS) {
          a0 = a0 + inv_dx * Fz1; a0 = a0 - inv_dx * Fz0; a0 = a0 + inv_dx * Ff.x; a0 = a0 - inv_dx * fb[f].x;
          a1 = a1 + inv_dx * Fz2; a1 = a1 - inv_dx * Fz1; a1 = a1 + inv_dx * Ff.y; a1 = a1 - inv_dx * fb[f].y;
        } else {
          const double z0 = vz[2], z1 = vz[3], r0 = v[1].x, r1 = v[1].y;
          a0 = a0 + inv_dx * Fz1 * z0; a0 = a0 - inv_dx * Fz0 * z0; a0 = a0 + inv_dx * Ff.x * r0; a0 = a0 - inv_dx * fb[f].x * r0;
          a1 = a1 + inv_dx * Fz2 * z1; a1 = a1 - inv_dx * Fz1 * z1; a1 = a1 + inv_dx * Ff.y * r1; a1 = a1 - inv_dx * fb[f].y * r1;
        }
        st2(po[f], FLUXONLY ? make_double2(a0, a1) : make_double2(cen.x + a0, cen.y + a1));
        fb[f] = Ff;
        qq[f][0] = qq[f][1]; qq[f][1] = qq[f][2]; qq[f][2] = qq[f][3];
        wcen[f][0] = wcen[f][1]; wcen[f][1] = wcen[f][2];
        pw[f] += ld; po[f] += ld;
      }
      v[0] = v[1]; v[1] = v[2]; v[2] = v[3];
      pv += ld; pu += ld;
    }
    return;
  }
  if (PATH != 2) return;
  // rolling state: velocity rows j-1..j+2 (u_r), per field q rows j-1..j+2 and the centres
  double2 vr[4];
  double2 q[2][4], wc[2][3], Fb[2];
  auto load_row = [&](int jj, double2& vrow, double2 qrow[2], double2 wrow[2]) {
    const int jc = clampi(jj, jmin, g.nr - 1);
    vrow = ld_row_m<MIRROR>(u_r, g.ld, jc, k, nz, vec, -1.0);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const double2 v = ld_row_m<MIRROR>(ins[f], g.ld, jc, k, nz, vec, signs[f]);
      wrow[f] = v;
      qrow[f] = CONS ? make_double2(v.x * vrow.x, v.y * vrow.y) : v;
    }
  };
  // prime the window for row j0: rows j0-2 (only for the first back face), j0-1, j0, j0+1
  double2 vm2, qm2[2], wtmp[2];
  load_row(j0 - 2, vm2, qm2, wtmp);
  {
    double2 qq[2], ww[2];
    load_row(j0 - 1, vr[0], qq, ww);
#pragma unroll
    for (int f = 0; f < NF; ++f) q[f][0] = qq[f];
    load_row(j0, vr[1], qq, ww);
#pragma unroll
    for (int f = 0; f < NF; ++f) { q[f][1] = qq[f]; wc[f][0] = ww[f]; }
    load_row(j0 + 1, vr[2], qq, ww);
#pragma unroll
    for (int f = 0; f < NF; ++f) { q[f][2] = qq[f]; wc[f][1] = ww[f]; }
  }
  // first back face (j0-1 | j0)
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    Fb[f].x = eno_face(qm2[f].x, q[f][0].x, q[f][1].x, q[f][2].x, vr[0].x, vr[1].x);
    Fb[f].y = eno_face(qm2[f].y, q[f][0].y, q[f][1].y, q[f][2].y, vr[0].y, vr[1].y);
  }

  for (int j = j0; j < j1; ++j) {
    {
      double2 qq[2], ww[2];
      load_row(j + 2, vr[3], qq, ww);
#pragma unroll
      for (int f = 0; f < NF; ++f) { q[f][3] = qq[f]; wc[f][2] = ww[f]; }
    }
    const bool row_ok = (j >= j_lo) && (j <= j_hi);
    const bool ok0 = row_ok && ok0c, ok1 = row_ok && ok1c;
    // z direction: u_z at columns k-2 .. k+3 of row j
    double vz[6];
    if (ok0 || ok1) {
      const double* uzr = rowp(u_z, g.ld, j);
      const double2 cz = ld_pair(uzr, k, nz, vec);
      const double2 lz = ld_pair(uzr, (k >= 2) ? k - 2 : k, nz, vec), rz = ld_pair(uzr, (k + 2 < nz) ? k + 2 : k, nz, vec);
      vz[0] = lz.x; vz[1] = lz.y; vz[2] = cz.x; vz[3] = cz.y; vz[4] = rz.x; vz[5] = rz.y;
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      // front face (j | j+1) from rows j-1 .. j+2
      double2 Ff;
      Ff.x = eno_face(q[f][0].x, q[f][1].x, q[f][2].x, q[f][3].x, vr[1].x, vr[2].x);
      Ff.y = eno_face(q[f][0].y, q[f][1].y, q[f][2].y, q[f][3].y, vr[1].y, vr[2].y);
      const double2 cen = wc[f][0];
      double2 o;
      if (FLUXONLY) o = ld_pair(rowp(outs[f], g.ld, j), k, nz, vec);
      else o = cen;
      if (ok0 || ok1) {
        const double* fr = rowp(ins[f], g.ld, j);
        double qz[6];
        {
          const double2 l = ld_pair(fr, (k >= 2) ? k - 2 : k, nz, vec), r = ld_pair(fr, (k + 2 < nz) ? k + 2 : k, nz, vec);
          qz[0] = l.x; qz[1] = l.y; qz[2] = cen.x; qz[3] = cen.y; qz[4] = r.x; qz[5] = r.y;
        }
        if (CONS) {
#pragma unroll
          for (int i = 0; i < 6; ++i) qz[i] *= vz[i];
        }
        const double Fz0 = eno_face(qz[0], qz[1], qz[2], qz[3], vz[1], vz[2]);
        const double Fz1 = eno_face(qz[1], qz[2], qz[3], qz[4], vz[2], vz[3]);
        const double Fz2 = eno_face(qz[2], qz[3], qz[4], qz[5], vz[3], vz[4]);
        double a0 = FLUXONLY ? o.x : 0.0, a1 = FLUXONLY ? o.y : 0.0;
        if (CONS) {
          a0 = a0 + inv_dx * Fz1; a0 = a0 - inv_dx * Fz0; a0 = a0 + inv_dx * Ff.x; a0 = a0 - inv_dx * Fb[f].x;
          a1 = a1 + inv_dx * Fz2; a1 = a1 - inv_dx * Fz1; a1 = a1 + inv_dx * Ff.y; a1 = a1 - inv_dx * Fb[f].y;
        } else {
          const double z0 = vz[2], z1 = vz[3], r0 = vr[1].x, r1 = vr[1].y;
          a0 = a0 + inv_dx * Fz1 * z0; a0 = a0 - inv_dx * Fz0 * z0; a0 = a0 + inv_dx * Ff.x * r0; a0 = a0 - inv_dx * Fb[f].x * r0;
          a1 = a1 + inv_dx * Fz2 * z1; a1 = a1 - inv_dx * Fz1 * z1; a1 = a1 + inv_dx * Ff.y * r1; a1 = a1 - inv_dx * Fb[f].y * r1;
        }
        if (FLUXONLY) { if (ok0) o.x = a0; if (ok1) o.y = a1; }
        else { if (ok0) o.x = cen.x + a0; if (ok1) o.y = cen.y + a1; }
      }
      if (!FLUXONLY || ok0 || ok1) st_pair(rowp(outs[f], g.ld, j), k, g.ku0, g.ku1, vec, o);
      Fb[f] = Ff;
      q[f][0] = q[f][1]; q[f][1] = q[f][2]; q[f][2] = q[f][3];
      wc[f][0] = wc[f][1]; wc[f][1] = wc[f][2];
    }
    vr[0] = vr[1]; vr[1] = vr[2]; vr[2] = vr[3];
  }
}

inline int pick_rb(const GridD& d) {
  // rows per block: as long as the grid still has >= ~16 blocks per SM keep the window start-up
  // (1-4 extra row loads) amortised over 32 rows
  const long long colb = ((d.nz + 1) / 2 + MT - 1) / MT;
  for (int rb = 32; rb >= 8; rb >>= 1)
    if (colb * ((d.nr + rb - 1) / rb) * d.batch >= 148LL * 12) return rb;
  return 8;
}
inline dim3 march_grid(const GridD& d, int rb) {
  return dim3(((d.nz + 1) / 2 + MT - 1) / MT, (d.nr + rb - 1) / rb, d.batch);
}
// Edge blocks of a launch, enumerated with the predicates the kernels use (a block the host lists but the device finds
// interior returns at once and is done by the interior kernel; the reverse cannot happen).  Falls back to the plain 2-D
// grid -- every block launched, the interior ones leave -- when 128-bit accesses are off (then every block is an edge
// block), when more than four column blocks or row ranges are irregular, or with AXB_EDGE_FULL_GRID=1.
inline EdgeMap edge_map(const GridD& d, int rb, int re, int hr, int hz, bool vec, bool owned_window, dim3& grid) {
  EdgeMap m;
  memset(&m, 0, sizeof(m));
  const int nbx = ((d.nz + 1) / 2 + MT - 1) / MT, nby = (d.nr + rb - 1) / rb;
  m.nfy = (d.nr + re - 1) / re;
  grid = dim3(nbx, m.nfy, d.batch);
  static const bool off = getenv("AXB_EDGE_FULL_GRID") != nullptr;
  if (off || !vec) return m;
  for (int bx = 0; bx < nbx; ++bx) {
    if (cols_interior(d, bx, hz)) continue;
    if (m.n_bc == 4) return m;
    m.bc[m.n_bc++] = bx;
  }
  const int per = rb / re;
  for (int by = 0; by < nby; ++by) {
    const int p0 = by * rb, p1 = (p0 + rb < d.nr) ? p0 + rb : d.nr;
    if (rows_interior(d, p0, p1, rb, hr, owned_window)) continue;
    const int f0 = by * per, f1 = (f0 + per < m.nfy) ? f0 + per : m.nfy;
    if (m.n_rr && m.r1[m.n_rr - 1] == f0) { m.r1[m.n_rr - 1] = f1; continue; }
    if (m.n_rr == 4) { m.n_bc = 0; m.n_rr = 0; return m; }
    m.r0[m.n_rr] = f0; m.r1[m.n_rr] = f1; ++m.n_rr;
  }
  for (int i = 0; i < m.n_rr; ++i) m.rows_r += m.r1[i] - m.r0[i];
  m.n_col_part = m.n_bc * m.nfy;
  m.total = m.n_col_part + (nbx - m.n_bc) * m.rows_r;
  m.on = 1;
  grid = dim3(m.total > 0 ? m.total : 1, 1, d.batch);
  return m;
}
// rows per block of the edge kernels (see row_chunk): 8, and 4 on grids so small that the interior chunk is already 8 rows
// (measured, compact edge grid: 516 x 16384 -- one rank's slab of the 8-GPU run -- 0.466 / 0.447 / 0.450 ms per step with
// 32 / 8 / 4 rows, 128 x 256 0.089 / 0.089 / 0.073, 4096 x 16384 2.851 / 2.849 / 2.856); AXB_EDGE_RB=4/8/16/32 forces one
inline int pick_re(int rb) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("AXB_EDGE_RB");
    forced = e ? atoi(e) : 0;
    if (forced != 4 && forced != 8 && forced != 16 && forced != 32) forced = 0;
  }
  const int re = forced ? forced : (rb <= 8 ? 4 : 8);
  return re < rb ? re : rb;
}

// The edge kernel of a pair (few blocks, latency bound: 40-90 us at 4096 x 16384) runs on a side stream
// forked from and joined back into the caller's stream, so it hides behind the interior kernel.  The two
// kernels write disjoint cells and only read the operation's inputs.  Works under stream capture (the
// side stream joins the capture through the fork event).
struct EdgeFork {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int device = -1;
  bool ok = false;
};
inline EdgeFork* edge_fork() {
  static thread_local EdgeFork f[16];
  static const bool off = getenv("AXB_EDGE_SERIAL") != nullptr;
  if (off) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  EdgeFork& e = f[dev];
  if (e.device != dev) {
    e.device = dev;
    e.ok = cudaStreamCreateWithFlags(&e.side, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&e.fork, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&e.join, cudaEventDisableTiming) == cudaSuccess;
    cudaGetLastError();
  }
  return e.ok ? &e : nullptr;
}
// stream the edge kernel goes to (falls back to the caller's stream)
inline cudaStream_t edge_begin(cudaStream_t s, EdgeFork*& f) {
  f = edge_fork();
  if (f && cudaEventRecord(f->fork, s) == cudaSuccess && cudaStreamWaitEvent(f->side, f->fork, 0) == cudaSuccess)
    return f->side;
  f = nullptr;
  cudaGetLastError();
  return s;
}
inline void edge_end(cudaStream_t s, EdgeFork* f) {
  if (!f) return;
  cudaEventRecord(f->join, f->side);
  cudaStreamWaitEvent(s, f->join, 0);
}

}  // namespace

int march_velocity(const GridD& d, double* u_z, double* u_r, const double* psi, const double* r1d, double uz_add,
                   double ur_add, const double* add_dev, double* umax_out, bool vec, cudaStream_t s) {
  const int rb = pick_rb(d);
  EdgeFork* ef;
  cudaStream_t se = edge_begin(s, ef);
  const int re = pick_re(rb);
  dim3 ge;
  const EdgeMap em = edge_map(d, rb, re, 1, 1, vec, umax_out != nullptr, ge);
#define VEL(R, P, ST) km_velocity<R, P><<<P == 2 ? ge : march_grid(d, rb), MT, rb * sizeof(double), ST>>>(                \
    d, rb, re, em, u_z, u_r, psi, r1d, uz_add, ur_add, add_dev, umax_out, vec)
  if (umax_out) { VEL(true, 2, se); VEL(true, 1, s); }
  else { VEL(false, 2, se); VEL(false, 1, s); }
#undef VEL
  edge_end(s, ef);
  return (int)cudaGetLastError();
}

int march_penalise(const GridD& d, double* u_z, double* u_r, double* w, const double* uzu, const double* uru,
                   const double* chi, double lam, double dt, const double* dt_dev, double U_z, double U_r,
                   const double* U_dev, const double* r1d, double* sum_out, bool vec, cudaStream_t s) {
  const int rb = pick_rb(d);
  EdgeFork* ef;
  cudaStream_t se = edge_begin(s, ef);
  const int re = pick_re(rb);
  dim3 ge;
  const EdgeMap em = edge_map(d, rb, re, 1, 1, vec, sum_out != nullptr, ge);
#define PEN(R, P, ST) km_penalise<R, P><<<P == 2 ? ge : march_grid(d, rb), MT, 0, ST>>>(d, rb, re, em, u_z, u_r, w, uzu, uru, chi, \
                                                                                  lam, dt, dt_dev, U_z, U_r, U_dev, r1d, sum_out, vec)
  if (sum_out) { PEN(true, 2, se); PEN(true, 1, s); }
  else { PEN(false, 2, se); PEN(false, 1, s); }
#undef PEN
  edge_end(s, ef);
  return (int)cudaGetLastError();
}

int march_diffusion(int stage, const GridD& d, double* out, const double* in, const double* src2, const double* r1d,
                    double nu, const double* nu_dev, double dt, const double* dt_dev, bool vec, cudaStream_t s) {
  const int rb = pick_rb(d);
  EdgeFork* ef;
  cudaStream_t se = edge_begin(s, ef);
  const int re = pick_re(rb);
  dim3 ge;
  const EdgeMap em = edge_map(d, rb, re, 1, 1, vec, false, ge);
#define DIF(S, P, ST) km_diffusion<S, P><<<P == 2 ? ge : march_grid(d, rb), MT, rb * sizeof(double), ST>>>(                \
    d, rb, re, em, out, in, src2, r1d, nu, nu_dev, dt, dt_dev, vec)
  if (stage == 1) { DIF(1, 2, se); DIF(1, 1, s); }
  else { DIF(2, 2, se); DIF(2, 1, s); }
#undef DIF
  edge_end(s, ef);
  return (int)cudaGetLastError();
}

int march_diffusion_fused(const GridD& d, double* out, const double* in, const double* r1d, double nu, double dt,
                          const double* dt_dev, bool vec, cudaStream_t s) {
  const int rb = pick_rb(d);
  EdgeFork* ef;
  cudaStream_t se = edge_begin(s, ef);
  const int re = pick_re(rb);
  dim3 ge;
  const EdgeMap em = edge_map(d, rb, re, 2, 2, vec, false, ge);
  km_diffusion_fused<2><<<ge, MT, (rb + 2) * sizeof(double), se>>>(d, rb, re, em, out, in, r1d, nu, dt, dt_dev, vec);
  km_diffusion_fused<1><<<march_grid(d, rb), MT, (rb + 2) * sizeof(double), s>>>(d, rb, re, em, out, in, r1d, nu, dt, dt_dev, vec);
  edge_end(s, ef);
  return (int)cudaGetLastError();
}

int march_eno3(int nf, bool cons, bool mirror, bool fluxonly, const GridD& d, double* out0, double* out1,
               const double* in0, const double* in1, const double* u_z, const double* u_r, double inv_dx, double dt,
               const double* dt_dev, double sign0, double sign1, bool vec, cudaStream_t s) {
  const int rb = pick_rb(d);
  const int re = pick_re(rb);
  const dim3 grd = march_grid(d, rb);
  dim3 grde;
  const EdgeMap em = edge_map(d, rb, re, 2, 2, vec, false, grde);
  EdgeFork* ef;
  cudaStream_t se = edge_begin(s, ef);
#define LAUNCH(NF, C, M, F)                                                                                                  \
  do {                                                                                                                      \
    km_eno3<NF, C, M, F, 2><<<grde, MT, 0, se>>>(d, rb, re, em, out0, out1, in0, in1, u_z, u_r, inv_dx, dt, dt_dev, sign0,  \
                                                 sign1, vec);                                                               \
    km_eno3<NF, C, M, F, 1><<<grd, MT, 0, s>>>(d, rb, re, em, out0, out1, in0, in1, u_z, u_r, inv_dx, dt, dt_dev, sign0,    \
                                               sign1, vec);                                                                 \
  } while (0)
  if (nf == 1 && cons && mirror && !fluxonly) LAUNCH(1, true, true, false);
  else if (nf == 2 && !cons && mirror && !fluxonly) LAUNCH(2, false, true, false);
  else if (nf == 1 && cons && !mirror && fluxonly) LAUNCH(1, true, false, true);
  else if (nf == 1 && !cons && !mirror && fluxonly) LAUNCH(1, false, false, true);
  else if (nf == 1 && cons && !mirror && !fluxonly) LAUNCH(1, true, false, false);
  else if (nf == 1 && !cons && !mirror && !fluxonly) LAUNCH(1, false, false, false);
  else { edge_end(s, ef); return AXB_ENOSUP; }
#undef LAUNCH
  edge_end(s, ef);
  return (int)cudaGetLastError();
}

// Test hook (no kernel launch, runs without a GPU): the edge blocks one launch of a row-marching pass would start on
// grid `g` -- (column block, fine row chunk) pairs in launch order -- with info = {rows per interior chunk, rows per
// edge chunk, compact grid on / off, launched blocks}.  tests/test_abi_cpu.py checks them against the definition.
extern "C" int axb_debug_edge_blocks(const axb_grid_t* g, int hr, int hz, int vec, int owned_window, int32_t* pairs,
                                     int cap, int32_t* info) {
  if (!g || !pairs || !info || cap < 0 || hr < 1 || hz < 1) return AXB_EINVAL;
  const GridD d = to_dev(g);
  if (d.nr < 1 || d.nz < 1) return AXB_EINVAL;
  const int rb = pick_rb(d), re = pick_re(rb);
  dim3 grid;
  const EdgeMap em = edge_map(d, rb, re, hr, hz, vec != 0, owned_window != 0, grid);
  int n = 0;
  for (unsigned ey = 0; ey < grid.y; ++ey)
    for (unsigned ex = 0; ex < grid.x; ++ex) {
      int bx, fy;
      if (!edge_decode(em, (int)ex, (int)ey, bx, fy)) continue;
      if (n < cap) { pairs[2 * n] = bx; pairs[2 * n + 1] = fy; }
      ++n;
    }
  info[0] = rb; info[1] = re; info[2] = em.on; info[3] = n;
  return AXB_OK;
}
