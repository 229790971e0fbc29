// reinit.cu -- narrow-band level-set re-initialisation, sm_100a.
//
// Replaces `skfmm.distance(phi, dx=dx, narrow=band)` of the soft-body drivers
// (examples/SoftSphereStreaming/soft_sphere_streaming.py:196-199,
//  examples/SoftSphereInTaylorGreenVortex/soft_sphere_in_taylor_green_vortex.py:160).
// scikit-fmm 2022.8.15 is a third-party dependency (poetry.lock:628-629) that is not vendored in the
// reference: PARITY UNPINNED.  oracle/axisym_oracle.py:fmm_distance restates its published
// fast-marching algorithm; tests/reinit_model.py models what this file does and is checked against
// that restatement on the CPU.
//
// The marcher accepts one cell at a time from a heap -- serial.  Here the same upwind update is iterated
// to its fixed point by all band cells at once:
//   k_reinit_front   whole grid, 8 B/pt read + 1 B/pt flag: cells whose 4-neighbourhood straddles the zero
//                    contour get the distance from the linear crossings (identical to the marcher's first
//                    step); they mark the 32x32-cell tiles within reach of the band;
//   k_reinit_compact active tiles -> list;  k_reinit_fill: the two value buffers of those tiles;
//   k_reinit_sweep   one block per active tile: tile + halo of 2 in shared memory, up to `inner` Jacobi
//                    iterations on chip (halo frozen; ends early when the tile is stationary), tiles whose
//                    neighbourhood did not change in the previous launch are skipped.  Every non-front cell
//                    recomputes its value from the neighbours the marcher would have frozen before it: front
//                    cells and cells with |value| <= narrow that are causally smaller than the result (a
//                    dimension whose upwind value is not below the 2-D result is dropped).  Second order where
//                    the two upwind cells are usable and monotone, else first order -- distance_marcher's
//                    selection rule, literally.  On a smooth contour the band settles in 2-3 launches; what
//                    remains are dependency chains of ~sqrt(2 R W) cells along the contour where it is axis
//                    aligned (one cell per iteration), a handful of tiles.  After `free_launch` launches values
//                    may only shrink in magnitude: where two fronts collide inside the band, or neighbours are
//                    exactly tied, the second-order selection can otherwise flip for ever.
//   k_reinit_ring    accepted cells -> phi; cells next to an accepted cell get the marcher's tentative value
//                    (update from all accepted neighbours, no causality filter); everything else keeps its
//                    old value (the driver's `ball_phi[mask] = bad_phi[mask]`).
// The fixed point does not depend on the iteration order; wherever the distance field is smooth it equals the
// marcher's result bit for bit (same expressions, -fmad=false).  The entry synchronises (convergence flag),
// like the LS extrapolation.
#include <float.h>

#include "axb_common.cuh"

namespace {

constexpr int TW = 32, TH = 32;  // tile = 32 columns x 32 rows, one thread per cell
constexpr int HALO = 2, SW = TW + 2 * HALO, SH = TH + 2 * HALO;
constexpr double MAXD = DBL_MAX;

struct Ctr {  // device counters
  int n_active, changed, negdet, has_front;
};

// cell values / front flags as the update sees them
struct GlobalAcc {  // dense (nr, nz) buffers; cells of inactive tiles were never initialised: unreached
  const double* __restrict__ d;
  const unsigned char* __restrict__ f;
  const int* __restrict__ tile_flag;
  int nz, ntc;
  __device__ __forceinline__ double val(int jj, int kk) const {
    return tile_flag[(jj / TH) * ntc + kk / TW] ? d[(long long)jj * nz + kk] : MAXD;
  }
  __device__ __forceinline__ bool flg(int jj, int kk) const { return f[(long long)jj * nz + kk] != 0; }
};
struct TileAcc {  // shared-memory tile with halo, origin (j0, k0) = global index of element [0][0]; flag bit 0 = front
  const double (*d)[SW];
  const unsigned char (*f)[SW];
  int j0, k0;
  __device__ __forceinline__ double val(int jj, int kk) const { return d[jj - j0][kk - k0]; }
  __device__ __forceinline__ bool flg(int jj, int kk) const { return (f[jj - j0][kk - k0] & 1) != 0; }
};

template <class Acc>
__device__ __forceinline__ bool usable(const Acc& A, double v, double narrow, int jj, int kk) {
  return fabs(v) <= narrow || A.flg(jj, kk);
}

// (value1, value2) of one dimension: distance_marcher's selection among the usable neighbours
template <class Acc>
__device__ __forceinline__ void upwind(const Acc& A, double narrow, int nr, int nz, int j, int k, int dim, int order,
                                       double& v1, double& v2) {
  v1 = MAXD;
  v2 = MAXD;
#pragma unroll
  for (int s = -1; s <= 1; s += 2) {
    const int jj = dim == 0 ? j + s : j, kk = dim == 0 ? k : k + s;
    if (jj < 0 || jj >= nr || kk < 0 || kk >= nz) continue;
    const double dn = A.val(jj, kk);
    if (!usable(A, dn, narrow, jj, kk)) continue;
    if (fabs(dn) < fabs(v1)) {
      v1 = dn;
      const int j2 = dim == 0 ? j + 2 * s : j, k2 = dim == 0 ? k : k + 2 * s;
      if (order == 2 && j2 >= 0 && j2 < nr && k2 >= 0 && k2 < nz) {
        const double d2 = A.val(j2, k2);
        if (usable(A, d2, narrow, j2, k2) && ((d2 <= v1 && v1 >= 0) || (d2 >= v1 && v1 <= 0))) v2 = d2;
      }
    }
  }
}

__device__ __forceinline__ void dim_terms(double v1, double v2, double idx2, double& a, double& b, double& c) {
  const double aa = 9.0 / 4.0;
  if (v2 < MAXD) {
    const double tp = (1.0 / 3.0) * (4 * v1 - v2);
    a = idx2 * aa;
    b = -(idx2 * 2 * aa * tp);
    c = idx2 * aa * (tp * tp);
  } else {
    a = idx2;
    b = -(idx2 * 2 * v1);
    c = idx2 * (v1 * v1);
  }
}

// root of a T^2 + b T + (c - 1) = 0 on the side of the cell's sign; false when the discriminant is negative
__device__ __forceinline__ bool quadratic(double a, double b, double c, bool positive, double& r) {
  c = c - 1;
  const double det = b * b - 4 * a * c;
  if (det < 0) return false;
  r = positive ? (-b + sqrt(det)) * 0.5 / a : (-b - sqrt(det)) * 0.5 / a;  // x / 2.0 == x * 0.5 exactly
  return true;
}

// new value of cell (j, k).  CAUSAL: the sweep form (returns MAXD when nothing usable); otherwise the
// marcher's tentative value from all usable neighbours (`ok` false on a negative discriminant).
template <bool CAUSAL, class Acc>
__device__ __forceinline__ double update_cell(const Acc& A, double narrow, int nr, int nz, int j, int k, double idx2,
                                              int order, bool positive, bool& ok) {
  double v1[2], v2[2];
  upwind(A, narrow, nr, nz, j, k, 0, order, v1[0], v2[0]);
  upwind(A, narrow, nr, nz, j, k, 1, order, v1[1], v2[1]);
  const bool h0 = v1[0] < MAXD, h1 = v1[1] < MAXD;
  ok = true;
  if (!h0 && !h1) return MAXD;
  int dim;
  if (h0 && h1) {
    double a0, b0, c0, a1, b1, c1, r;
    dim_terms(v1[0], v2[0], idx2, a0, b0, c0);
    dim_terms(v1[1], v2[1], idx2, a1, b1, c1);
    const bool good = quadratic(a0 + a1, b0 + b1, c0 + c1, positive, r);
    if (!CAUSAL) {
      ok = good;
      return good ? r : MAXD;
    }
    const double big = fmax(fabs(v1[0]), fabs(v1[1]));
    if (good && fabs(r) > big) return r;
    dim = fabs(v1[0]) <= fabs(v1[1]) ? 0 : 1;
  } else {
    dim = h0 ? 0 : 1;
  }
  double a, b, c, r;
  dim_terms(v1[dim], v2[dim], idx2, a, b, c);
  const bool good = quadratic(a, b, c, positive, r);
  if (!CAUSAL) ok = good;
  return good ? r : MAXD;
}

// -------------------------------------------------------------------------------------
// distance of a cell whose 4-neighbourhood straddles the zero contour (the marcher's initial front).
// q[dim][side] are the neighbours (side 0: index - 1, side 1: index + 1), h[dim][side] whether they exist.
__device__ __forceinline__ bool front_cell(double p, const double q[2][2], const bool h[2][2], double dx, double& dist) {
  if (p == 0.0) {
    dist = 0.0;
    return true;
  }
  double ldist[2] = {0.0, 0.0};
  bool borders = false;
#pragma unroll
  for (int dim = 0; dim < 2; ++dim) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (h[dim][s] && p * q[dim][s] < 0) {
        borders = true;
        const double c = dx * p / (p - q[dim][s]);
        if (ldist[dim] == 0 || ldist[dim] > c) ldist[dim] = c;
      }
    }
  }
  if (!borders) return false;
  double dsum = 0.0;
#pragma unroll
  for (int dim = 0; dim < 2; ++dim)
    if (ldist[dim] > 0) dsum += 1 / ldist[dim] / ldist[dim];
  dist = p < 0 ? -sqrt(1 / dsum) : sqrt(1 / dsum);
  return true;
}

constexpr int FT = 128, FR = 8;  // front pass: 128 threads x 2 columns, marching over 8 rows

// Whole-grid pass, 8 B/pt read + 1 B/pt written: a thread owns two adjacent columns and walks down FR rows with
// the r-neighbourhood in registers (every row of phi is loaded once; the z neighbours of the pair come from L1).
__global__ void __launch_bounds__(FT)
    k_reinit_front(int nr, int nz, long long ld, double dx, const double* __restrict__ phi, double* __restrict__ dA,
                   unsigned char* __restrict__ flag, int* __restrict__ tile_flag, int reach, int ntr, int ntc, bool vec,
                   Ctr* ctr) {
  const int k = 2 * (blockIdx.x * FT + threadIdx.x);
  if (k >= nz) return;
  const int ja = blockIdx.y * FR, jb = min(ja + FR, nr);
  const bool has1 = k + 1 < nz, hasl = k > 0, hasr = k + 2 < nz;
  double2 prev = make_double2(0, 0), cur = ld_pair(phi + (long long)ja * ld, k, nz, vec), next = cur;
  if (ja > 0) prev = ld_pair(phi + (long long)(ja - 1) * ld, k, nz, vec);
  for (int j = ja; j < jb; ++j) {
    const double* row = phi + (long long)j * ld;
    const bool hasu = j + 1 < nr, hasd = j > 0;
    if (hasu) next = ld_pair(row + ld, k, nz, vec);
    const double left = hasl ? row[k - 1] : 0.0, right = hasr ? row[k + 2] : 0.0;
    // cheap test first: no sign change (and no exact zero) in the pair's neighbourhood -> nothing to do
    const bool quiet0 = cur.x != 0.0 && !(hasd && cur.x * prev.x < 0) && !(hasu && cur.x * next.x < 0) &&
                        !(hasl && cur.x * left < 0) && !(has1 && cur.x * cur.y < 0);
    const bool quiet1 = !has1 || (cur.y != 0.0 && !(hasd && cur.y * prev.y < 0) && !(hasu && cur.y * next.y < 0) &&
                                  !(cur.y * cur.x < 0) && !(hasr && cur.y * right < 0));
    unsigned char f0 = 0, f1 = 0;
    if (!(quiet0 && quiet1)) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c == 1 && !has1) continue;
        const double p = c ? cur.y : cur.x;
        const double q[2][2] = {{c ? prev.y : prev.x, c ? next.y : next.x}, {c ? cur.x : left, c ? right : cur.y}};
        const bool h[2][2] = {{hasd, hasu}, {c ? true : hasl, c ? hasr : has1}};
        double dist;
        if (front_cell(p, q, h, dx, dist)) {
          const int kk = k + c;
          (c ? f1 : f0) = 1;
          dA[(long long)j * nz + kk] = dist;  // the other cells of the active tiles are filled by k_reinit_fill
          ctr->has_front = 1;
          const int t0 = max(j - reach, 0) / TH, t1 = min(j + reach, nr - 1) / TH;
          const int c0 = max(kk - reach, 0) / TW, c1 = min(kk + reach, nz - 1) / TW;
          for (int t = t0; t <= t1 && t < ntr; ++t)
            for (int cc = c0; cc <= c1 && cc < ntc; ++cc) tile_flag[t * ntc + cc] = 1;
        }
      }
    }
    unsigned char* frow = flag + (long long)j * nz;
    frow[k] = f0;
    if (has1) frow[k + 1] = f1;
    prev = cur;
    cur = next;
  }
}

__global__ void k_reinit_compact(const int* __restrict__ tile_flag, int ntiles, int* __restrict__ tile_list,
                                 int* __restrict__ chg_a, int* __restrict__ chg_b, Ctr* ctr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int on = tile_flag[t] != 0;
  chg_a[t] = on;  // "changed before the first launch": every active tile is visited once
  chg_b[t] = 0;
  if (on) tile_list[atomicAdd(&ctr->n_active, 1)] = t;
}

// both value buffers of the active tiles: front distance (written by k_reinit_front) or "unreached"
__global__ void __launch_bounds__(TW* TH)
    k_reinit_fill(int nr, int nz, double* __restrict__ dA, double* __restrict__ dB,
                  const unsigned char* __restrict__ flag, const int* __restrict__ tile_list, int ntc) {
  const int t = tile_list[blockIdx.x];
  const int k = (t % ntc) * TW + threadIdx.x;
  const int j = (t / ntc) * TH + threadIdx.y;
  if (j >= nr || k >= nz) return;
  const long long i = (long long)j * nz + k;
  const double v = flag[i] ? dA[i] : MAXD;
  dA[i] = v;
  dB[i] = v;
}

// One launch = up to `inner` Jacobi iterations of every active tile in shared memory.  `chg_prev[t]` says whether
// tile t changed in the previous launch; a tile whose own and neighbouring tiles did not change has nothing to do
// (its inputs are what they were and both global buffers already hold its values).
// 512 threads per tile (two cells each in the selection phase): two tiles share an SM, so the ~250 active tiles of
// the 2048 x 8192 case run in one wave -- the kernel is bound by the sequential depth of the iterations, not by
// throughput.
constexpr int SWEEP_ROWS = TH / 2, SWEEP_THREADS = TW * SWEEP_ROWS;

__global__ void __launch_bounds__(SWEEP_THREADS, 2)
    k_reinit_sweep(int nr, int nz, long long ld, double idx2, const double* __restrict__ phi,
                   const double* __restrict__ din, double* __restrict__ dout, const unsigned char* __restrict__ flag,
                   const int* __restrict__ tile_flag, const int* __restrict__ tile_list, int ntr, int ntc,
                   const int* __restrict__ chg_prev, int* __restrict__ chg_cur, double narrow, int order,
                   int monotone, int inner, Ctr* ctr) {
  __shared__ double sd[2][SH][SW];
  __shared__ unsigned char sf[SH][SW];     // bit 0: front cell, bit 1: phi > eps (own cells)
  __shared__ unsigned char sc[2][SH][SW];  // "changed in the last iteration", per cell and buffer parity
  __shared__ unsigned short s_list[TW * TH];
  __shared__ int s_n[2];
  __shared__ int s_act;
  const int tid = threadIdx.y * TW + threadIdx.x;
  const int t = tile_list[blockIdx.x];
  const int tj = t / ntc, tc = t % ntc;
  if (tid == 0) {
    int a = 0;
    for (int dj = -1; dj <= 1; ++dj)
      for (int dc = -1; dc <= 1; ++dc) {
        const int pj = tj + dj, pc = tc + dc;
        if (pj >= 0 && pj < ntr && pc >= 0 && pc < ntc) a |= chg_prev[pj * ntc + pc];
      }
    s_act = a;
    s_n[0] = 0;
    s_n[1] = 0;
  }
  __syncthreads();
  if (!s_act) {  // block-uniform
    if (tid == 0) chg_cur[t] = 0;
    return;
  }
  const int j0 = tj * TH - HALO, k0 = tc * TW - HALO;
  for (int idx = tid; idx < SH * SW; idx += SWEEP_THREADS) {
    const int lj = idx / SW, lk = idx % SW;
    const int jj = j0 + lj, kk = k0 + lk;
    double v = MAXD;
    unsigned char fl = 0;
    if (jj >= 0 && jj < nr && kk >= 0 && kk < nz && tile_flag[(jj / TH) * ntc + kk / TW]) {
      const long long i = (long long)jj * nz + kk;
      v = din[i];
      fl = flag[i];
    }
    sd[0][lj][lk] = v;
    sd[1][lj][lk] = v;
    sf[lj][lk] = fl;
    sc[0][lj][lk] = 0;  // halo, front and outside cells never change within a launch
    sc[1][lj][lk] = 0;
  }
  __syncthreads();
  // this thread's two cells: rows threadIdx.y and threadIdx.y + SWEEP_ROWS of the tile
  const int lk = threadIdx.x + HALO, k = k0 + lk;
  int ljc[2];
  bool active[2];
  double start[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    ljc[c] = threadIdx.y + c * SWEEP_ROWS + HALO;
    const int j = j0 + ljc[c];
    const bool inside = j < nr && k < nz;
    active[c] = inside && !(sf[ljc[c]][lk] & 1);  // front cells are fixed
    if (inside && phi[(long long)j * ld + k] > DBL_EPSILON) sf[ljc[c]][lk] |= 2;  // own elements only
    start[c] = sd[0][ljc[c]][lk];
  }
  __syncthreads();
  // Jacobi iterations.  Per iteration only a thin curve of cells has new inputs; those cells are compacted into a
  // list and evaluated by the first threads of the block, so the cost follows the cells that move, not the tile.
  int b = 0;
  for (int q = 0; q < inner; ++q) {
    int* cnt = &s_n[q & 1];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (!active[c]) continue;
      const int lj = ljc[c];
      // a cell whose eight stencil inputs did not change in the last iteration would recompute the value it has
      const bool need = q == 0 || sc[b][lj - 1][lk] || sc[b][lj + 1][lk] || sc[b][lj][lk - 1] || sc[b][lj][lk + 1] ||
                        sc[b][lj - 2][lk] || sc[b][lj + 2][lk] || sc[b][lj][lk - 2] || sc[b][lj][lk + 2];
      if (need) {
        s_list[atomicAdd(cnt, 1)] = (unsigned short)(lj * SW + lk);
      } else {
        sd[b ^ 1][lj][lk] = sd[b][lj][lk];
        sc[b ^ 1][lj][lk] = 0;
      }
    }
    if (tid == 0) s_n[(q & 1) ^ 1] = 0;  // the other counter: last read before the barrier that ended iteration q-1
    __syncthreads();
    int ch = 0;
    const int n = *cnt;
    for (int e = tid; e < n; e += SWEEP_THREADS) {
      const int cell = s_list[e];
      const int clj = cell / SW, clk = cell % SW;
      const double cur = sd[b][clj][clk];
      TileAcc A{sd[b], sf, j0, k0};
      bool ok;
      double r = update_cell<true>(A, narrow, nr, nz, j0 + clj, k0 + clk, idx2, order, (sf[clj][clk] & 2) != 0, ok);
      if (monotone && !(fabs(r) < fabs(cur))) r = cur;
      const int c1 = __double_as_longlong(r) != __double_as_longlong(cur);
      ch |= c1;
      sd[b ^ 1][clj][clk] = r;
      sc[b ^ 1][clj][clk] = (unsigned char)c1;
    }
    b ^= 1;
    if (!__syncthreads_or(ch)) break;  // also orders this iteration's writes before the next one's reads
  }
  int c1 = 0;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (!active[c]) continue;
    const double fin = sd[b][ljc[c]][lk];
    dout[(long long)(j0 + ljc[c]) * nz + k] = fin;
    c1 |= __double_as_longlong(fin) != __double_as_longlong(start[c]);
  }
  const int any1 = __syncthreads_or(c1);
  if (tid == 0) {
    chg_cur[t] = any1;
    if (any1) ctr->changed = 1;
  }
}

__global__ void __launch_bounds__(TW* TH)
    k_reinit_ring(int nr, int nz, long long ld, double idx2, double* __restrict__ phi, const double* __restrict__ d,
                  const unsigned char* __restrict__ flag, const int* __restrict__ tile_flag,
                  const int* __restrict__ tile_list, int ntc, double narrow, int order,
                  unsigned char* __restrict__ mask_out, Ctr* ctr) {
  const int t = tile_list[blockIdx.x];
  const int k = (t % ntc) * TW + threadIdx.x;
  const int j = (t / ntc) * TH + threadIdx.y;
  if (j >= nr || k >= nz) return;
  const long long i = (long long)j * nz + k;
  const GlobalAcc A{d, flag, tile_flag, nz, ntc};
  const double mine = d[i];
  // accepted = what the marcher froze: front cells and cells whose value is within the band
  if (fabs(mine) <= narrow || flag[i]) {
    phi[(long long)j * ld + k] = mine;
    if (mask_out) mask_out[i] = 0;
    return;
  }
  bool touches = false;
  if (j > 0) touches |= usable(A, A.val(j - 1, k), narrow, j - 1, k);
  if (j < nr - 1) touches |= usable(A, A.val(j + 1, k), narrow, j + 1, k);
  if (k > 0) touches |= usable(A, A.val(j, k - 1), narrow, j, k - 1);
  if (k < nz - 1) touches |= usable(A, A.val(j, k + 1), narrow, j, k + 1);
  if (!touches) return;
  bool ok;
  const double r = update_cell<false>(A, narrow, nr, nz, j, k, idx2, order, phi[(long long)j * ld + k] > DBL_EPSILON, ok);
  if (!ok) {
    ctr->negdet = 1;
    return;
  }
  phi[(long long)j * ld + k] = r;
  if (mask_out) mask_out[i] = 0;
}

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

// pinned host mirror of the counters (a pageable destination makes every read-back a staged, slower copy);
// allocated once per process, the drivers are single threaded (SURVEY.md 8b "Threading")
Ctr* host_ctr() {
  static Ctr* p = nullptr;
  if (!p && cudaHostAlloc((void**)&p, sizeof(Ctr), cudaHostAllocDefault) != cudaSuccess) p = nullptr;
  return p;
}

}  // namespace

extern "C" {

int64_t axb_reinit_workspace_bytes(int nr, int nz) {
  if (nr < 1 || nz < 1) return 0;
  const int64_t n = (int64_t)nr * nz;
  const int64_t ntiles = (int64_t)((nr + TH - 1) / TH) * ((nz + TW - 1) / TW);
  return 2 * align256(8 * n) + align256(n) + 4 * align256(4 * ntiles) + 256;
}

// info_host[0] = sweep launches run, info_host[1] = status bits: 1 no zero contour, 2 negative discriminant in
// the tentative ring, 4 no fixed point within the launch bound.
int axb_reinit_distance(const axb_grid_t* g, double* phi, double narrow, int order, unsigned char* mask_out,
                        void* work, int64_t work_bytes, int* info_host, axb_stream_t s) {
  if (!phi || !work || !info_host) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  if (!axb_al8(phi) || (((uintptr_t)work) & 255u)) return AXB_EALIGN;
  if (order != 1 && order != 2) return AXB_EINVAL;
  if (!(narrow > 0.0) || !(g->dx > 0.0)) return AXB_EINVAL;
  if (g->kz0 != 0 || g->nz_global != g->nz || g->ku0 != 0 || g->ku1 != g->nz) return AXB_ENOSUP;  // whole domain only
  const int nr = g->nr, nz = g->nz;
  if (work_bytes < axb_reinit_workspace_bytes(nr, nz)) return AXB_EWORK;
  const int64_t n = (int64_t)nr * nz;
  const int ntr = (nr + TH - 1) / TH, ntc = (nz + TW - 1) / TW;
  const int64_t ntiles = (int64_t)ntr * ntc;
  char* w = (char*)work;
  double* dA = (double*)w;            w += align256(8 * n);
  double* dB = (double*)w;            w += align256(8 * n);
  unsigned char* flag = (unsigned char*)w; w += align256(n);
  int* tile_flag = (int*)w;           w += align256(4 * ntiles);
  int* tile_list = (int*)w;           w += align256(4 * ntiles);
  int* chg_prev = (int*)w;            w += align256(4 * ntiles);
  int* chg_cur = (int*)w;             w += align256(4 * ntiles);
  Ctr* ctr = (Ctr*)w;
  info_host[0] = 0;
  info_host[1] = 0;

  const double wcells = ceil(narrow / g->dx);
  if (wcells > 1.0e6) return AXB_EINVAL;
  const int W = (int)wcells;
  // accepted cells lie within W (+1) cells of a front cell, the tentative ring one further, its stencil two more
  const int reach = W + 4;
  // on-chip iterations per launch: a dependency chain advances one cell per iteration, a tile is 32 cells wide
  const int inner = 64;
  // launches needed on a smooth contour: 2-3 for the band, then the chains along the contour where it is axis
  // aligned, ~sqrt(2 R W) cells for a radius of curvature of R cells (<= the grid size), ~TW cells per launch
  const int longest = nr > nz ? nr : nz;
  const int chain = 2 * W + (int)ceil(2.0 * sqrt(2.0 * (double)longest * (double)W));
  const int free_launch = 8 + 2 * ((chain + TW - 1) / TW);
  const int max_launch = 2 * free_launch;

  cudaError_t e;
  if ((e = cudaMemsetAsync(tile_flag, 0, align256(4 * ntiles), s)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(ctr, 0, sizeof(Ctr), s)) != cudaSuccess) return (int)e;
  if (mask_out && (e = cudaMemsetAsync(mask_out, 1, n, s)) != cudaSuccess) return (int)e;
  const dim3 blk(TW, TH);
  const double idx2 = 1 / g->dx / g->dx;  // the marcher's 1/dx^2, rounded the way it rounds it
  const bool vec = !(g->ld & 1) && axb_al16(phi);
  k_reinit_front<<<dim3((nz + 2 * FT - 1) / (2 * FT), (nr + FR - 1) / FR), FT, 0, s>>>(
      nr, nz, g->ld, g->dx, phi, dA, flag, tile_flag, reach, ntr, ntc, vec, ctr);
  AXB_LAUNCHED();
  k_reinit_compact<<<(unsigned)((ntiles + 255) / 256), 256, 0, s>>>(tile_flag, (int)ntiles, tile_list, chg_prev,
                                                                    chg_cur, ctr);
  AXB_LAUNCHED();
  Ctr* hp = host_ctr();
  if (!hp) return (int)cudaErrorMemoryAllocation;
  Ctr& h = *hp;
  if ((e = cudaMemcpyAsync(&h, ctr, sizeof(Ctr), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return (int)e;
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  if (!h.has_front || h.n_active == 0) {
    info_host[1] = 1;
    return AXB_OK;
  }
  const int nact = h.n_active;
  k_reinit_fill<<<nact, blk, 0, s>>>(nr, nz, dA, dB, flag, tile_list, ntc);
  AXB_LAUNCHED();
  double *din = dA, *dout = dB;
  int it = 0;
  bool converged = false;
  while (it < max_launch && !converged) {
    if ((e = cudaMemsetAsync(&ctr->changed, 0, sizeof(int), s)) != cudaSuccess) return (int)e;
    ++it;
    k_reinit_sweep<<<nact, dim3(TW, SWEEP_ROWS), 0, s>>>(nr, nz, g->ld, idx2, phi, din, dout, flag, tile_flag, tile_list, ntr, ntc,
                                        chg_prev, chg_cur, narrow, order, it > free_launch ? 1 : 0, inner, ctr);
    AXB_LAUNCHED();
    double* tmp = din; din = dout; dout = tmp;
    int* tc_ = chg_prev; chg_prev = chg_cur; chg_cur = tc_;
    // the band cannot settle in one launch (tile halos are frozen within a launch); afterwards look at the flag
    // after every second launch -- a launch that finds every tile stationary costs a few microseconds
    if (it == 1 || ((it & 1) == 0 && it < max_launch)) continue;
    if ((e = cudaMemcpyAsync(&h, ctr, sizeof(Ctr), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return (int)e;
    converged = (h.changed == 0);
  }
  info_host[0] = it;
  if (!converged) info_host[1] |= 4;
  // `din` now holds the latest values
  k_reinit_ring<<<nact, blk, 0, s>>>(nr, nz, g->ld, idx2, phi, din, flag, tile_flag, tile_list, ntc, narrow, order,
                                     mask_out, ctr);
  AXB_LAUNCHED();
  if ((e = cudaMemcpyAsync(&h, ctr, sizeof(Ctr), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return (int)e;
  if (h.negdet) info_host[1] |= 2;
  return (int)cudaGetLastError();
}

}  // extern "C"
