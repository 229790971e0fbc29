// reinit.cu -- narrow-band level-set re-initialisation, sm_100a.
//
// Replaces `skfmm.distance(phi, dx=dx, narrow=band)` of the soft-body drivers
// (examples/SoftSphereStreaming/soft_sphere_streaming.py:196-199,
//  examples/SoftSphereInTaylorGreenVortex/soft_sphere_in_taylor_green_vortex.py:160).
// scikit-fmm 2022.8.15 is a third-party dependency (poetry.lock:628-629) that is not vendored in the
// reference: PARITY UNPINNED.  oracle/axisym_oracle.py:fmm_distance restates its published
// fast-marching algorithm; tools/reinit_model.py models what this file does and is checked against
// that restatement on the CPU.
//
// The marcher accepts one cell at a time from a heap -- serial.  Here the same upwind update is iterated
// to its fixed point by all band cells at once (Jacobi, two buffers, so the result does not depend on
// scheduling):
//   k_reinit_front   whole grid, 8 B/pt read: cells whose 4-neighbourhood straddles the zero contour get the
//                    distance from the linear crossings (identical to the marcher's first step); they mark
//                    the 64x8-cell tiles within reach of the band;
//   k_reinit_compact active tiles -> list (the sweeps only visit those: O(band) work, L2 resident);
//   k_reinit_sweep   every non-front cell of the active tiles recomputes its value from the neighbours the
//                    marcher would have frozen before it: front cells and cells with |value| <= narrow that
//                    are causally smaller than the result (a dimension whose upwind value is not below the
//                    2-D result is dropped).  Second order where the two upwind cells are usable and
//                    monotone, else first order -- distance_marcher's selection rule, literally.
//                    A state that repeats with period 2 (tied neighbours flipping in the last bits) also ends
//                    the iteration (k_reinit_pick keeps the value of smaller magnitude).  After `free_iter`
//                    sweeps values may only shrink in magnitude: where two fronts collide inside the band the
//                    second-order selection can otherwise flip for ever.
//   k_reinit_ring    accepted cells -> phi; cells next to an accepted cell get the marcher's tentative value
//                    (update from all accepted neighbours, no causality filter); everything else keeps its
//                    old value (the driver's `ball_phi[mask] = bad_phi[mask]`).
// Wherever the distance field is smooth the fixed point equals the marcher's result bit for bit (same
// expressions, -fmad=false).  The entry synchronises (convergence flag), like the LS extrapolation.
#include <float.h>

#include "axb_common.cuh"

namespace {

constexpr int TW = 64, TH = 8;  // tile = 64 columns x 8 rows, one thread per cell
constexpr double MAXD = DBL_MAX;

struct Ctr {  // device counters (changed / changed2 adjacent: reset together)
  int n_active, changed, changed2, negdet, has_front;
};

__device__ __forceinline__ bool usable(double d, double narrow, const unsigned char* __restrict__ flag,
                                       long long idx) {
  return fabs(d) <= narrow || (flag != nullptr && flag[idx] != 0);
}

// (value1, value2) of one dimension: distance_marcher's selection among the usable neighbours
__device__ __forceinline__ void upwind(const double* __restrict__ d, const unsigned char* __restrict__ flag,
                                       double narrow, int nr, int nz, int j, int k, int dim, int order, double& v1,
                                       double& v2) {
  v1 = MAXD;
  v2 = MAXD;
#pragma unroll
  for (int s = -1; s <= 1; s += 2) {
    const int jj = dim == 0 ? j + s : j, kk = dim == 0 ? k : k + s;
    if (jj < 0 || jj >= nr || kk < 0 || kk >= nz) continue;
    const long long i1 = (long long)jj * nz + kk;
    const double dn = d[i1];
    if (!usable(dn, narrow, flag, i1)) continue;
    if (fabs(dn) < fabs(v1)) {
      v1 = dn;
      const int j2 = dim == 0 ? j + 2 * s : j, k2 = dim == 0 ? k : k + 2 * s;
      if (order == 2 && j2 >= 0 && j2 < nr && k2 >= 0 && k2 < nz) {
        const long long i2 = (long long)j2 * nz + k2;
        const double d2 = d[i2];
        if (usable(d2, narrow, flag, i2) && ((d2 <= v1 && v1 >= 0) || (d2 >= v1 && v1 <= 0))) v2 = d2;
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
  r = positive ? (-b + sqrt(det)) / 2.0 / a : (-b - sqrt(det)) / 2.0 / a;
  return true;
}

// new value of cell (j, k).  CAUSAL: the sweep form (returns MAXD when nothing usable); otherwise the
// marcher's tentative value from all usable neighbours (`ok` false on a negative discriminant).
template <bool CAUSAL>
__device__ __forceinline__ double update_cell(const double* __restrict__ d, const unsigned char* __restrict__ flag,
                                              double narrow, int nr, int nz, int j, int k, double dx, int order,
                                              bool positive, bool& ok) {
  const double idx2 = 1 / dx / dx;
  double v1[2], v2[2];
  upwind(d, flag, narrow, nr, nz, j, k, 0, order, v1[0], v2[0]);
  upwind(d, flag, narrow, nr, nz, j, k, 1, order, v1[1], v2[1]);
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
__global__ void __launch_bounds__(TW* TH)
    k_reinit_front(int nr, int nz, long long ld, double dx, const double* __restrict__ phi, double* __restrict__ dA,
                   double* __restrict__ dB, unsigned char* __restrict__ flag, int* __restrict__ tile_flag, int reach,
                   int ntr, int ntc, Ctr* ctr) {
  const int k = blockIdx.x * TW + threadIdx.x;
  const int j = blockIdx.y * TH + threadIdx.y;
  if (j >= nr || k >= nz) return;
  const double p = phi[(long long)j * ld + k];
  double dist = MAXD;
  bool front = false;
  if (p == 0.0) {
    dist = 0.0;
    front = true;
  } else {
    double ldist[2] = {0.0, 0.0};
    bool borders = false;
#pragma unroll
    for (int dim = 0; dim < 2; ++dim) {
#pragma unroll
      for (int s = -1; s <= 1; s += 2) {
        const int jj = dim == 0 ? j + s : j, kk = dim == 0 ? k : k + s;
        if (jj < 0 || jj >= nr || kk < 0 || kk >= nz) continue;
        const double q = phi[(long long)jj * ld + kk];
        if (p * q < 0) {
          borders = true;
          const double c = dx * p / (p - q);
          if (ldist[dim] == 0 || ldist[dim] > c) ldist[dim] = c;
        }
      }
    }
    if (borders) {
      double dsum = 0.0;
#pragma unroll
      for (int dim = 0; dim < 2; ++dim)
        if (ldist[dim] > 0) dsum += 1 / ldist[dim] / ldist[dim];
      dist = p < 0 ? -sqrt(1 / dsum) : sqrt(1 / dsum);
      front = true;
    }
  }
  const long long i = (long long)j * nz + k;
  dA[i] = dist;
  dB[i] = dist;
  flag[i] = front ? 1 : 0;
  if (front) {
    ctr->has_front = 1;
    const int t0 = max(j - reach, 0) / TH, t1 = min(j + reach, nr - 1) / TH;
    const int c0 = max(k - reach, 0) / TW, c1 = min(k + reach, nz - 1) / TW;
    for (int t = t0; t <= t1 && t < ntr; ++t)
      for (int c = c0; c <= c1 && c < ntc; ++c) tile_flag[t * ntc + c] = 1;
  }
}

__global__ void k_reinit_compact(const int* __restrict__ tile_flag, int ntiles, int* __restrict__ tile_list,
                                 int* __restrict__ chg_a, int* __restrict__ chg_b, Ctr* ctr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int on = tile_flag[t] != 0;
  chg_a[t] = on;  // "changed before the first sweep": every active tile is visited once
  chg_b[t] = 0;
  if (on) tile_list[atomicAdd(&ctr->n_active, 1)] = t;
}

// One Jacobi sweep over the active tiles.  `chg_prev[t]` says whether tile t changed in the previous sweep;
// a tile whose own and neighbouring tiles (the stencil reaches 2 cells, less than a tile) did not change has
// nothing to do: its inputs are what they were, and both buffers already hold its values.  Late sweeps, when
// only the long tangential dependency chains near the axis-aligned points of the contour are still moving,
// therefore touch a handful of tiles.
__global__ void __launch_bounds__(TW* TH)
    k_reinit_sweep(int nr, int nz, long long ld, double dx, const double* __restrict__ phi,
                   const double* __restrict__ din, double* __restrict__ dout, const unsigned char* __restrict__ flag,
                   bool use_flag, const int* __restrict__ tile_list, int ntr, int ntc, const int* __restrict__ chg_prev,
                   int* __restrict__ chg_cur, double narrow, int order, int monotone, Ctr* ctr) {
  __shared__ int s_act;
  const int t = tile_list[blockIdx.x];
  const int tj = t / ntc, tc = t % ntc;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    int a = 0;
    for (int dj = -1; dj <= 1; ++dj)
      for (int dc = -1; dc <= 1; ++dc) {
        const int pj = tj + dj, pc = tc + dc;
        if (pj >= 0 && pj < ntr && pc >= 0 && pc < ntc) a |= chg_prev[pj * ntc + pc];
      }
    s_act = a;
  }
  __syncthreads();
  if (!s_act) {  // block-uniform
    if (threadIdx.x == 0 && threadIdx.y == 0) chg_cur[t] = 0;
    return;
  }
  const int k = tc * TW + threadIdx.x;
  const int j = tj * TH + threadIdx.y;
  int c1 = 0, c2 = 0;
  if (j < nr && k < nz) {
    const long long i = (long long)j * nz + k;
    if (!flag[i]) {  // front cells are fixed (both buffers hold their distance)
      const double cur = din[i];
      bool ok;
      double r = update_cell<true>(din, use_flag ? flag : nullptr, narrow, nr, nz, j, k, dx, order,
                                   phi[(long long)j * ld + k] > DBL_EPSILON, ok);
      if (monotone && !(fabs(r) < fabs(cur))) r = cur;
      const double before = dout[i];  // the state two sweeps ago
      dout[i] = r;
      c1 = __double_as_longlong(r) != __double_as_longlong(cur);
      c2 = __double_as_longlong(r) != __double_as_longlong(before);
    }
  }
  const int any1 = __syncthreads_or(c1);
  const int any2 = __syncthreads_or(c2);
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    chg_cur[t] = any1;
    if (any1) ctr->changed = 1;
    if (any2) ctr->changed2 = 1;
  }
}

// period-2 cycle (tied neighbours that take each other as upwind cell when the rounding of the 2-D root says
// so; the two states differ in the last bits): keep, per cell, the value of smaller magnitude.
// `older` = state n, `latest` = state n+1 (== state n-1); the result goes to `latest`.
__global__ void __launch_bounds__(TW* TH)
    k_reinit_pick(int nr, int nz, const double* __restrict__ older, double* __restrict__ latest,
                  const unsigned char* __restrict__ flag, const int* __restrict__ tile_list, int ntc) {
  const int t = tile_list[blockIdx.x];
  const int k = (t % ntc) * TW + threadIdx.x;
  const int j = (t / ntc) * TH + threadIdx.y;
  if (j >= nr || k >= nz) return;
  const long long i = (long long)j * nz + k;
  if (flag[i]) return;
  const double a = older[i], b = latest[i];
  latest[i] = fabs(a) <= fabs(b) ? a : b;
}

__global__ void __launch_bounds__(TW* TH)
    k_reinit_ring(int nr, int nz, long long ld, double dx, double* __restrict__ phi, const double* __restrict__ d,
                  const unsigned char* __restrict__ flag, const int* __restrict__ tile_list, int ntc, double narrow,
                  int order, unsigned char* __restrict__ mask_out, Ctr* ctr) {
  const int t = tile_list[blockIdx.x];
  const int k = (t % ntc) * TW + threadIdx.x;
  const int j = (t / ntc) * TH + threadIdx.y;
  if (j >= nr || k >= nz) return;
  const long long i = (long long)j * nz + k;
  const double mine = d[i];
  // accepted = what the marcher froze: front cells and cells whose value is within the band
  if (fabs(mine) <= narrow || flag[i]) {
    phi[(long long)j * ld + k] = mine;
    if (mask_out) mask_out[i] = 0;
    return;
  }
  bool touches = false;
  if (j > 0) touches |= fabs(d[i - nz]) <= narrow || flag[i - nz];
  if (j < nr - 1) touches |= fabs(d[i + nz]) <= narrow || flag[i + nz];
  if (k > 0) touches |= fabs(d[i - 1]) <= narrow || flag[i - 1];
  if (k < nz - 1) touches |= fabs(d[i + 1]) <= narrow || flag[i + 1];
  if (!touches) return;
  bool ok;
  const double r = update_cell<false>(d, flag, narrow, nr, nz, j, k, dx, order,
                                      phi[(long long)j * ld + k] > DBL_EPSILON, ok);
  if (!ok) {
    ctr->negdet = 1;
    return;
  }
  phi[(long long)j * ld + k] = r;
  if (mask_out) mask_out[i] = 0;
}

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

}  // namespace

extern "C" {

int64_t axb_reinit_workspace_bytes(int nr, int nz) {
  if (nr < 1 || nz < 1) return 0;
  const int64_t n = (int64_t)nr * nz;
  const int64_t ntiles = (int64_t)((nr + TH - 1) / TH) * ((nz + TW - 1) / TW);
  return 2 * align256(8 * n) + align256(n) + 4 * align256(4 * ntiles) + 256;
}

// info_host[0] = sweeps run, info_host[1] = status bits: 1 no zero contour, 2 negative discriminant in the
// tentative ring, 4 sweeps did not reach a fixed point within the bound.
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
  const int reach = W + 3;  // accepted cells lie within W cells of a front cell, the ring one further, +1 slack
  // sweeps needed on a smooth contour: ~2W across the band plus the tangential dependency chains where the
  // contour is axis aligned, ~sqrt(2 R W) cells for a radius of curvature of R cells (<= the grid size)
  const int longest = nr > nz ? nr : nz;
  const int free_iter = 8 * W + 64 + (int)ceil(2.0 * sqrt(2.0 * (double)longest * (double)W));
  const int max_iter = 2 * free_iter;
  const bool use_flag = narrow < g->dx;  // front distances are <= dx: flags only matter for thinner bands

  cudaError_t e;
  if ((e = cudaMemsetAsync(tile_flag, 0, align256(4 * ntiles), s)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(ctr, 0, sizeof(Ctr), s)) != cudaSuccess) return (int)e;
  if (mask_out && (e = cudaMemsetAsync(mask_out, 1, n, s)) != cudaSuccess) return (int)e;
  const dim3 blk(TW, TH);
  k_reinit_front<<<dim3(ntc, ntr), blk, 0, s>>>(nr, nz, g->ld, g->dx, phi, dA, dB, flag, tile_flag, reach, ntr, ntc,
                                               ctr);
  AXB_LAUNCHED();
  k_reinit_compact<<<(unsigned)((ntiles + 255) / 256), 256, 0, s>>>(tile_flag, (int)ntiles, tile_list, chg_prev,
                                                                    chg_cur, ctr);
  AXB_LAUNCHED();
  Ctr h;
  if ((e = cudaMemcpyAsync(&h, ctr, sizeof(Ctr), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return (int)e;
  if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  if (!h.has_front || h.n_active == 0) {
    info_host[1] = 1;
    return AXB_OK;
  }
  const int nact = h.n_active;
  double *din = dA, *dout = dB;
  int it = 0;
  bool converged = false;
  while (it < max_iter && !converged) {
    // a group of sweeps, the last one with the change flag armed
    const int group = (it < W) ? (W - it) : (it < 4 * W + 16 ? 4 : 8);
    for (int q = 0; q < group && it < max_iter; ++q) {
      const bool last = (q == group - 1) || (it + 1 == max_iter);
      if (last && (e = cudaMemsetAsync(&ctr->changed, 0, 2 * sizeof(int), s)) != cudaSuccess) return (int)e;
      ++it;
      k_reinit_sweep<<<nact, blk, 0, s>>>(nr, nz, g->ld, g->dx, phi, din, dout, flag, use_flag, tile_list, ntr, ntc,
                                          chg_prev, chg_cur, narrow, order, it > free_iter ? 1 : 0, ctr);
      AXB_LAUNCHED();
      double* tmp = din; din = dout; dout = tmp;
      int* tc_ = chg_prev; chg_prev = chg_cur; chg_cur = tc_;
    }
    if ((e = cudaMemcpyAsync(&h, ctr, sizeof(Ctr), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return (int)e;
    converged = (h.changed == 0);
    if (!converged && h.changed2 == 0) {
      // state n+1 == state n-1: `din` holds n+1, `dout` holds n
      k_reinit_pick<<<nact, blk, 0, s>>>(nr, nz, dout, din, flag, tile_list, ntc);
      AXB_LAUNCHED();
      converged = true;
    }
  }
  info_host[0] = it;
  if (!converged) info_host[1] |= 4;
  // `din` now holds the latest values
  k_reinit_ring<<<nact, blk, 0, s>>>(nr, nz, g->ld, g->dx, phi, din, flag, tile_list, ntc, narrow, order, mask_out,
                                     ctr);
  AXB_LAUNCHED();
  if ((e = cudaMemcpyAsync(&h, ctr, sizeof(Ctr), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return (int)e;
  if (h.negdet) info_host[1] |= 2;
  return (int)cudaGetLastError();
}

}  // extern "C"
