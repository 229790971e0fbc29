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

// all passes of the length-M complex FFT; data starts in `a`, returns the buffer holding the result
__device__ __forceinline__ double2* pfft_passes(double2* a, double2* b, const double2* __restrict__ tab, int M,
                                               const PfftFactors& F) {
  int ns = 1;
  for (int p = 0; p < F.n; ++p) {
    const int R = F.r[p];
    const int L = M / R;
    const int step_j = M / (ns * R);              // table stride of the pass twiddle per unit of jm
    for (int q = threadIdx.x; q < M; q += blockDim.x) {
      const int jm = q % ns;
      const int t1 = q / ns;
      const int u = t1 % R;
      const int j = (t1 / R) * ns + jm;
      int e = (int)(((long long)jm * step_j + (long long)u * L) % M);
      double2 acc = a[j];                          // t = 0: twiddle 1
      int idx = e;
      for (int t = 1; t < R; ++t) {
        const double2 v = a[j + t * L];
        const double2 w = tab[idx];
        acc.x += v.x * w.x - v.y * w.y;
        acc.y += v.x * w.y + v.y * w.x;
        idx += e;
        if (idx >= M) idx -= M;
      }
      b[q] = acc;
    }
    __syncthreads();
    double2* t = a; a = b; b = t;
    ns *= R;
  }
  return a;
}

// forward: rows of N reals -> half-complex rows (pitch ld_dst >= N, tail zero-filled up to `pad_to`)
__global__ void __launch_bounds__(256)
    k_rfft_rows(int rows, int N, PfftFactors F, const double* __restrict__ src, long long ld_src, double* __restrict__ dst,
                long long ld_dst, int pad_to, const double2* __restrict__ tabM, const double2* __restrict__ tabN, double scale,
                int vec) {
  extern __shared__ __align__(16) unsigned char pf_smem[];
  const int M = N >> 1;
  double2* a = reinterpret_cast<double2*>(pf_smem);
  double2* b = a + M;
  double2* tab = b + M;
  for (int i = threadIdx.x; i < M; i += blockDim.x) tab[i] = tabM[i];
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const double* x = src + (long long)row * ld_src;
    if (vec) {
      const double2* x2 = reinterpret_cast<const double2*>(x);
      for (int n = threadIdx.x; n < M; n += blockDim.x) a[n] = x2[n];
    } else {
      for (int n = threadIdx.x; n < M; n += blockDim.x) a[n] = make_double2(x[2 * n], x[2 * n + 1]);
    }
    __syncthreads();
    const double2* Z = pfft_passes(a, b, tab, M, F);
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
__global__ void __launch_bounds__(256)
    k_irfft_rows(int rows, int N, PfftFactors F, const double* __restrict__ src, long long ld_src, double* __restrict__ dst,
                 long long ld_dst, const double2* __restrict__ tabM, const double2* __restrict__ tabN, double scale, int vec) {
  extern __shared__ __align__(16) unsigned char pf_smem[];
  const int M = N >> 1;
  double2* a = reinterpret_cast<double2*>(pf_smem);
  double2* b = a + M;
  double2* tab = b + M;
  for (int i = threadIdx.x; i < M; i += blockDim.x) tab[i] = tabM[i];
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
    const double2* z = pfft_passes(a, b, tab, M, F);
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

}  // namespace

// shared with fd_gemm.cu (axb_fd_solve)
int launch_rfft_rows(int inverse, int rows, int N, const double* src, long long ld_src, double* dst, long long ld_dst,
                     int pad_to, const double* tables, double scale, cudaStream_t st) {
  if (rows < 1 || !src || !dst || !tables || N < 4 || (N & 1) || ld_src < N || ld_dst < N) return AXB_EINVAL;
  const int M = N / 2;
  PfftFactors F;
  if (!pfft_factorize(M, &F)) return AXB_ENOSUP;
  const size_t smem = (size_t)3 * M * sizeof(double2);
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
  if (resident > 8) resident = 8;
  if (resident < 1) resident = 1;
  const int grid = rows < sms * resident ? rows : sms * resident;
  const double2* tabM = reinterpret_cast<const double2*>(tables);
  const double2* tabN = tabM + M;
  const int vec = (axb_al16(src) && axb_al16(dst) && (ld_src % 2 == 0) && (ld_dst % 2 == 0)) ? 1 : 0;
  if (!inverse)
    k_rfft_rows<<<grid, 256, smem, st>>>(rows, N, F, src, ld_src, dst, ld_dst, pad_to < N ? N : pad_to, tabM, tabN, scale,
                                         vec);
  else
    k_irfft_rows<<<grid, 256, smem, st>>>(rows, N, F, src, ld_src, dst, ld_dst, tabM, tabN, scale, vec);
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
  return (n >= 4 && !(n & 1) && pfft_factorize(n / 2, &F) && (size_t)3 * (n / 2) * sizeof(double2) <= 200 * 1024) ? 1 : 0;
}

}  // extern "C"
