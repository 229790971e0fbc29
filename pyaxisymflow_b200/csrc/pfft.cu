// pfft.cu -- z transforms of the fast-diagonalisation solve for PERIODIC z
// (pyaxisymflow/kernels/FastDiagonalisationStokesSolver.py:88-93: the periodic z operator, whose eigenvectors are
// the Fourier modes; the reference finds them with la.eig and applies them as two dense (nr x N)(N x N) products,
// :137-156; driver examples/PeriodicFlowPastSphere/periodic_flow_past_sphere.py:83-104, inner grid N = Nz - 4).
//
// A real FFT of every row, ONE pass over HBM per transform, for any even N = 2M whose half length M has only
// small prime factors (e.g. 4092 = 2 * 2 * 3 * 11 * 31):
//
//   forward : z[n] = x[2n] + i x[2n+1]  (the row read as complex pairs) -> complex FFT_M in shared memory by Stockham
//             auto-sort passes, one pass per factor R, every output the R-term sum  sum_t in[j + t M/R] w^(t e)  with
//             w = exp(-2 pi i / M) from one on-chip table (the pass twiddle and the R-point DFT kernel are the same
//             table walked with stride e) -> real-FFT untangling -> half-complex row
//             [Re X_0 .. Re X_M | Im X_1 .. Im X_{M-1}]  (N reals; columns N .. pitch-1 of the spectral buffer are
//             zero-filled so that the tridiagonal sweeps may run on a 16-column-aligned width)
//   inverse : the same passes on conj(Z) (inverse FFT through conjugation), Z rebuilt from the half-complex row.
//
// tools/rfft_model.py restates the index arithmetic in NumPy and is checked against numpy.fft on the CPU.
// Column c of the half-complex layout belongs to mode m(c) = c (c <= M), c - M (c > M); the r solve of column c uses
// lam_z[c] = (2 - 2 cos(2 pi m / N)) / dx^2, the same for the real and the imaginary part of a mode.
#include <math.h>

#include "axb_common.cuh"

namespace {

struct PfftFactors {
  int n;         // number of passes
  int r[12];     // radix of each pass
};

__device__ __forceinline__ double2 cmulf(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// all passes of the length-M complex FFT; data starts in `a`, returns the buffer holding the result.
// A work item is (butterfly j, group of PG outputs u0 .. u0+PG-1): every input of the butterfly is read once per
// item and multiplied by its pass twiddle (global table, L1 resident) once, then feeds PG accumulators with the
// R-point DFT weights W_R^(u t) = wt[t R + u], an R x R table per pass built once per CTA, read at the same
// address by all lanes of a warp (items of a warp share the output group) with immediate offsets -- no index
// arithmetic in the inner loop.
constexpr int PG = 4;
constexpr int PB = 512;          // threads per CTA
__device__ __forceinline__ void pfft_build_weights(double2* wt, const double2* __restrict__ tabM, int M, const PfftFactors& F) {
  int off = 0;
  for (int p = 0; p < F.n; ++p) {
    const int R = F.r[p], L = M / R;
    for (int i = threadIdx.x; i < R * R; i += blockDim.x) {
      const int t = i / R, u = i - t * R;
      wt[off + i] = tabM[((u * t) % R) * L];
    }
    off += R * R;
  }
}
__device__ __forceinline__ double2* pfft_passes(double2* a, double2* b, const double2* __restrict__ tabM,
                                               const double2* __restrict__ wt, int M, const PfftFactors& F) {
  int ns = 1, off = 0;
  for (int p = 0; p < F.n; ++p) {
    const int R = F.r[p];
    const int L = M / R;
    const int step_j = M / (ns * R);              // table stride of the pass twiddle per unit of jm
    const int NG = (R + PG - 1) / PG;
    for (int w = threadIdx.x; w < L * NG; w += blockDim.x) {
      const int ug = w / L;
      const int j = w - ug * L;
      const int u0 = ug * PG;
      const int nv = min(PG, R - u0);
      const int jm = j % ns;
      const int dtw = jm * step_j;
      const double2 v0 = a[j];
      double2 acc[PG];
#pragma unroll
      for (int g = 0; g < PG; ++g) acc[g] = v0;
      int ktw = 0;
      const double2* wrow = wt + off + u0;
#pragma unroll 2
      for (int t = 1; t < R; ++t) {
        double2 v = a[j + t * L];
        if (ns > 1) {
          ktw += dtw;
          if (ktw >= M) ktw -= M;
          v = cmulf(v, __ldg(&tabM[ktw]));
        }
        wrow += R;
#pragma unroll
        for (int g = 0; g < PG; ++g) {
          if (g < nv) {
            const double2 c = wrow[g];
            acc[g].x += v.x * c.x - v.y * c.y;
            acc[g].y += v.x * c.y + v.y * c.x;
          }
        }
      }
      const int base = (j - jm) * R + jm;
#pragma unroll
      for (int g = 0; g < PG; ++g)
        if (g < nv) b[base + (u0 + g) * ns] = acc[g];
    }
    __syncthreads();
    double2* t = a; a = b; b = t;
    ns *= R;
    off += R * R;
  }
  return a;
}

// forward: rows of N reals -> half-complex rows (pitch ld_dst >= N, tail zero-filled up to `pad_to`)
__global__ void __launch_bounds__(PB, 2)
    k_rfft_rows(int rows, int N, PfftFactors F, const double* __restrict__ src, long long ld_src, double* __restrict__ dst,
                long long ld_dst, int pad_to, const double2* __restrict__ tabM, const double2* __restrict__ tabN, double scale,
                int vec) {
  extern __shared__ __align__(16) unsigned char pf_smem[];
  const int M = N >> 1;
  double2* a = reinterpret_cast<double2*>(pf_smem);
  double2* b = a + M;
  double2* wt = b + M;
  pfft_build_weights(wt, tabM, M, F);
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const double* x = src + (long long)row * ld_src;
    if (vec) {
      const double2* x2 = reinterpret_cast<const double2*>(x);
      for (int n = threadIdx.x; n < M; n += blockDim.x) a[n] = x2[n];
    } else {
      for (int n = threadIdx.x; n < M; n += blockDim.x) a[n] = make_double2(x[2 * n], x[2 * n + 1]);
    }
    __syncthreads();
    const double2* Z = pfft_passes(a, b, tabM, wt, M, F);
    double* X = dst + (long long)row * ld_dst;
    for (int k = threadIdx.x; k <= M; k += blockDim.x) {
      const double2 zk = Z[k == M ? 0 : k];
      const double2 zq = Z[k == 0 ? 0 : M - k];
      const double2 zm = make_double2(zq.x, -zq.y);
      const double2 E = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y + zm.y));
      const double2 D = make_double2(zk.x - zm.x, zk.y - zm.y);
      const double2 O = make_double2(0.5 * D.y, -0.5 * D.x);             // -i/2 (zk - zm)
      const double2 w = tabN[k];
      const double2 P = cmulf(w, O);
      X[k] = (E.x + P.x) * scale;
      if (k > 0 && k < M) X[M + k] = (E.y + P.y) * scale;
    }
    for (int c = N + threadIdx.x; c < pad_to; c += blockDim.x) X[c] = 0.0;
    __syncthreads();                                                       // Z (in a or b) is dead: next row may load
  }
}

// inverse: half-complex rows -> rows of N reals, times `scale` (the caller folds 1/M = 2/N in)
__global__ void __launch_bounds__(PB, 2)
    k_irfft_rows(int rows, int N, PfftFactors F, const double* __restrict__ src, long long ld_src, double* __restrict__ dst,
                 long long ld_dst, const double2* __restrict__ tabM, const double2* __restrict__ tabN, double scale, int vec) {
  extern __shared__ __align__(16) unsigned char pf_smem[];
  const int M = N >> 1;
  double2* a = reinterpret_cast<double2*>(pf_smem);
  double2* b = a + M;
  double2* wt = b + M;
  pfft_build_weights(wt, tabM, M, F);
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const double* h = src + (long long)row * ld_src;
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
      // X[k] and conj(X[M - k]); Im X_0 = Im X_M = 0
      const double2 xk = make_double2(h[k], k == 0 ? 0.0 : h[M + k]);
      const int mk = M - k;
      const double2 xm = make_double2(h[mk], (mk == M) ? 0.0 : -h[M + mk]);
      const double2 E = make_double2(0.5 * (xk.x + xm.x), 0.5 * (xk.y + xm.y));
      const double2 D = make_double2(0.5 * (xk.x - xm.x), 0.5 * (xk.y - xm.y));
      const double2 w = tabN[k];
      const double2 O = cmulf(D, make_double2(w.x, -w.y));                // times exp(+2 pi i k / N)
      // Z = E + i O; the passes get conj(Z)
      a[k] = make_double2(E.x - O.y, -(E.y + O.x));
    }
    __syncthreads();
    const double2* z = pfft_passes(a, b, tabM, wt, M, F);
    double* x = dst + (long long)row * ld_dst;
    if (vec) {
      double2* x2 = reinterpret_cast<double2*>(x);
      for (int n = threadIdx.x; n < M; n += blockDim.x) x2[n] = make_double2(z[n].x * scale, -z[n].y * scale);
    } else {
      for (int n = threadIdx.x; n < M; n += blockDim.x) {
        x[2 * n] = z[n].x * scale;
        x[2 * n + 1] = -z[n].y * scale;
      }
    }
    __syncthreads();
  }
}

bool pfft_factorize(int M, PfftFactors* F) {
  F->n = 0;
  int m = M;
  int twos = 0;
  while (m % 2 == 0) { m /= 2; ++twos; }
  for (int i = 0; i < twos / 2; ++i) { if (F->n >= 12) return false; F->r[F->n++] = 4; }
  if (twos & 1) { if (F->n >= 12) return false; F->r[F->n++] = 2; }
  for (int p = 3; p <= 64 && m > 1; p += 2)
    while (m % p == 0) {
      if (F->n >= 12) return false;
      F->r[F->n++] = p;
      m /= p;
    }
  return m == 1 && M >= 2;
}
// shared memory: two row buffers + the R x R weight tables of all passes
size_t pfft_smem_bytes(int M, const PfftFactors& F) {
  size_t w = 0;
  for (int p = 0; p < F.n; ++p) w += (size_t)F.r[p] * F.r[p];
  return ((size_t)2 * M + w) * sizeof(double2);
}

}  // namespace

// shared with fd_gemm.cu (axb_fd_solve)
int launch_rfft_rows(int inverse, int rows, int N, const double* src, long long ld_src, double* dst, long long ld_dst,
                     int pad_to, const double* tables, double scale, cudaStream_t st) {
  if (rows < 1 || !src || !dst || !tables || N < 4 || (N & 1) || ld_src < N || ld_dst < N) return AXB_EINVAL;
  const int M = N / 2;
  PfftFactors F;
  if (!pfft_factorize(M, &F)) return AXB_ENOSUP;
  const size_t smem = pfft_smem_bytes(M, F);
  if (smem > 200 * 1024) return AXB_ENOSUP;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_rfft_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_irfft_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  int resident = (int)((227u * 1024u) / (smem + 1024));
  if (resident > 2) resident = 2;             // 512 threads x <= 64 registers: two CTAs per SM
  if (resident < 1) resident = 1;
  const int grid = rows < sms * resident ? rows : sms * resident;
  const double2* tabM = reinterpret_cast<const double2*>(tables);
  const double2* tabN = tabM + M;
  const int vec = (axb_al16(src) && axb_al16(dst) && (ld_src % 2 == 0) && (ld_dst % 2 == 0)) ? 1 : 0;
  if (!inverse)
    k_rfft_rows<<<grid, PB, smem, st>>>(rows, N, F, src, ld_src, dst, ld_dst, pad_to < N ? N : pad_to, tabM, tabN, scale,
                                         vec);
  else
    k_irfft_rows<<<grid, PB, smem, st>>>(rows, N, F, src, ld_src, dst, ld_dst, tabM, tabN, scale, vec);
  AXB_LAUNCHED();
  return (int)cudaGetLastError();
}

extern "C" {

// tables: [ exp(-2 pi i k / M), k = 0 .. M-1 | exp(-2 pi i k / N), k = 0 .. M ] as (re, im) pairs, N + 1 pairs
int axb_rfft_rows(int rows, int n, const double* src, int64_t ld_src, double* dst, int64_t ld_dst, int pad_to,
                  const double* tables, double scale, axb_stream_t s) {
  return launch_rfft_rows(0, rows, n, src, ld_src, dst, ld_dst, pad_to, tables, scale, (cudaStream_t)s);
}
int axb_irfft_rows(int rows, int n, const double* src, int64_t ld_src, double* dst, int64_t ld_dst,
                   const double* tables, double scale, axb_stream_t s) {
  return launch_rfft_rows(1, rows, n, src, ld_src, dst, ld_dst, 0, tables, scale, (cudaStream_t)s);
}
int axb_rfft_supported(int n) {
  PfftFactors F;
  return (n >= 4 && !(n & 1) && pfft_factorize(n / 2, &F) && pfft_smem_bytes(n / 2, F) <= 200 * 1024) ? 1 : 0;
}

}  // extern "C"
