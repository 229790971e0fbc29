// tridiag.cu -- batched tridiagonal r solves of the fast-diagonalisation plan with the pivots
// factored once.
//
// After the z transform every z-mode k is an independent system
//     (c0 I + c1 (A_r + lam_z[k] I)) x = rhs_k,      A_r = tridiag(sub, diag, sup)
// (the same operator pyaxisymflow/kernels/FastDiagonalisationStokesSolver.py:130-156 inverts through
// the eigen-decomposition of A_r).  The LU pivots depend on (row, mode) only, so they are computed
// once per plan (axb_tridiag_factor_columns, a division chain per column) and each solve is two
// streaming sweeps without a division:
//     forward :  y_m = rhs_m * scale_m - (a_m / den_{m-1}) y_{m-1}
//     backward:  x_m = (y_m - u_m x_{m+1}) / den_m
// One 16-lane CTA owns 16 adjacent columns (128-byte row segments) and walks down the rows; the
// right-hand side and the reciprocal pivots arrive through an 8-stage cp.async ring in shared
// memory, so each SM keeps ~100 KB of loads in flight although only 16384 chains exist.
#include "axb_common.cuh"

namespace {

constexpr int TC = 16;   // columns per CTA
constexpr int TR = 8;    // rows per stage
constexpr int TS = 8;    // stages

__device__ __forceinline__ void cp16(void* smem, const void* gmem, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(128)
    k_tri_factor(int nr, int nz, const double* __restrict__ sub, const double* __restrict__ diag,
                 const double* __restrict__ sup, const double* __restrict__ lam, double c0, double c1,
                 double* __restrict__ inv) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nz) return;
  const double lk = lam[k];
  double prev = 0.0;                                  // u_{m-1} / den_{m-1}
  for (int m = 0; m < nr; ++m) {
    const double b = c0 + c1 * (diag[m] + lk);
    const double a = (m > 0) ? c1 * sub[m - 1] : 0.0;
    const double den = b - a * prev;
    const double r = 1.0 / den;
    inv[(long long)m * nz + k] = r;
    prev = (m < nr - 1) ? c1 * sup[m] * r : 0.0;
  }
}

// stage `st` <- rows [m0, m0 + TR) (DIR = +1) or [m0 - TR + 1, m0] walked downwards (DIR = -1); row u of
// the stage is m0 + DIR * u.  16 lanes x 16 bytes cover two 128-byte row segments per instruction.
template <int DIR>
__device__ __forceinline__ void issue_stage(double* sx, double* sp, int st, int m0, int nr, const double* __restrict__ X,
                                            long long ld, const double* __restrict__ inv, int nz, int k0, int lane) {
  const int half = lane >> 3, c = (lane & 7) * 2;
#pragma unroll
  for (int u2 = 0; u2 < TR; u2 += 2) {
    const int u = u2 + half;
    const int m = m0 + DIR * u;
    const bool ok = (m >= 0) && (m < nr);
    const long long mm = ok ? m : 0;
    cp16(sx + (st * TR + u) * TC + c, X + mm * ld + k0 + c, ok);
    cp16(sp + (st * TR + u) * TC + c, inv + mm * nz + k0 + c, ok);
  }
}

// DIR = +1: forward elimination in place (X <- y); DIR = -1: back substitution in place (X <- x).
template <int DIR>
__global__ void __launch_bounds__(TC)
    k_tri_sweep(int nr, int nz, double* __restrict__ X, long long ld, const double* __restrict__ inv,
                const double* __restrict__ lo, const double* __restrict__ up, const double* __restrict__ scale,
                double c1) {
  __shared__ __align__(16) double sx[TS * TR * TC];
  __shared__ __align__(16) double sp[TS * TR * TC];
  const int lane = threadIdx.x;
  const int k0 = blockIdx.x * TC;
  const int nb = (nr + TR - 1) / TR;
  const int start = (DIR > 0) ? 0 : nr - 1;
#pragma unroll
  for (int i = 0; i < TS - 1; ++i) {
    if (i < nb) issue_stage<DIR>(sx, sp, i, start + DIR * i * TR, nr, X, ld, inv, nz, k0, lane);
    cp_commit();
  }
  double carry = 0.0;        // y_{m-1} (forward) / x_{m+1} (backward)
  double pprev = 0.0;        // 1 / den_{m-1} (forward only)
  for (int i = 0; i < nb; ++i) {
    __syncwarp(0xffffu);                                             // stage (i - 1) % TS is free again
    const int nxt = i + TS - 1;
    if (nxt < nb) issue_stage<DIR>(sx, sp, nxt % TS, start + DIR * nxt * TR, nr, X, ld, inv, nz, k0, lane);
    cp_commit();
    cp_wait<TS - 1>();
    __syncwarp(0xffffu);
    const int st = i % TS;
#pragma unroll
    for (int u = 0; u < TR; ++u) {
      const int m = start + DIR * (i * TR + u);
      if (m < 0 || m >= nr) break;
      const double r = sx[(st * TR + u) * TC + lane];
      const double p = sp[(st * TR + u) * TC + lane];
      if (DIR > 0) {
        const double a = (m > 0) ? c1 * lo[m - 1] : 0.0;
        const double rhs = scale ? r * scale[m] : r;
        carry = rhs - (a * pprev) * carry;
        pprev = p;
      } else {
        const double cu = (m < nr - 1) ? c1 * up[m] : 0.0;
        carry = (r - cu * carry) * p;
      }
      X[(long long)m * ld + k0 + lane] = carry;
    }
  }
}

}  // namespace

bool tri_fast_ok(int nz, const double* X, long long ld, const double* inv) {
  return inv && (nz % TC == 0) && axb_al16(X) && axb_al16(inv) && (ld % 2 == 0);
}

int launch_tri_factored(int nr, int nz, double* X, long long ld, const double* inv, const double* sub,
                        const double* sup, const double* scale, double c1, cudaStream_t s) {
  if (nr < 2 || nz < 1 || !X || !inv || !sub || !sup || ld < nz) return AXB_EINVAL;
  if (!tri_fast_ok(nz, X, ld, inv)) return AXB_EINVAL;
  k_tri_sweep<1><<<nz / TC, TC, 0, s>>>(nr, nz, X, ld, inv, sub, sup, scale, c1);
  AXB_LAUNCHED();
  k_tri_sweep<-1><<<nz / TC, TC, 0, s>>>(nr, nz, X, ld, inv, sub, sup, scale, c1);
  AXB_LAUNCHED();
  return (int)cudaGetLastError();
}

extern "C" {

int axb_tridiag_factor_columns(int nr, int nz, const double* sub, const double* diag, const double* sup,
                               const double* lam, double c0, double c1, double* inv_pivots, axb_stream_t s) {
  if (nr < 2 || nz < 1 || !sub || !diag || !sup || !lam || !inv_pivots) return AXB_EINVAL;
  k_tri_factor<<<(nz + 127) / 128, 128, 0, (cudaStream_t)s>>>(nr, nz, sub, diag, sup, lam, c0, c1, inv_pivots);
  AXB_LAUNCHED();
  return (int)cudaGetLastError();
}

int axb_tridiag_solve_factored(int nr, int nz, double* X, int64_t ld, const double* inv_pivots, const double* sub,
                               const double* sup, const double* scale, double c1, axb_stream_t s) {
  return launch_tri_factored(nr, nz, X, ld, inv_pivots, sub, sup, scale, c1, (cudaStream_t)s);
}

}  // extern "C"
