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
  int pg[12];    // odd radices: output pairs per work item (1..3)
};

__device__ __forceinline__ double2 cmulf(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// All passes of the length-M complex FFT; data starts in `a`, returns the buffer holding the result
// (tools/rfft_model.py: cfft_passes_paired).
//   * The pass twiddle W_M^(jm t M / (ns R)) of input (j, t) is applied when the PREVIOUS pass stores that element
//     (every element is read by exactly one (j, t) of the next pass): a butterfly reads plain values, and nothing is
//     twiddled more than once.
//   * An odd radix R pairs the inputs t, R-t and the outputs u, R-u:
//         S_t = x_t + x_(R-t),  D_t = x_t - x_(R-t),  A_u = sum_t S_t cos(2 pi u t / R),  B_u = sum_t D_t sin(2 pi u t / R)
//         X_u = x_0 + A_u - i B_u,   X_(R-u) = x_0 + A_u + i B_u,   X_0 = x_0 + sum_t S_t
//     -- a quarter of the multiplications of the R x R sum (R = 31: 900 instead of 3720 FMAs per butterfly).  A work
//     item is (butterfly j, group of PGP output pairs); the (cos, sin) weights of a pass are an H x H table
//     (H = (R-1)/2) in shared memory, read at the same address by all lanes of a warp.  PGP is chosen per pass on the
//     host so that the items of the pass fill the 512 threads in as few rounds as possible (F.pg).
//   * Radix 2 and 4 are the usual butterflies.
constexpr int PB = 512;          // threads per CTA

// a / d for 0 <= a < 2^15, 1 <= d < 2^15 without the integer-division sequence: (a + 0.5) / d is at least 0.5 / d
// > 1.5e-5 away from an integer, far more than the rounding of the float product (< 2^15 * 2^-22), so the floor is exact
__device__ __forceinline__ int qdiv_small(int a, float inv_d) { return __float2int_rd(((float)a + 0.5f) * inv_d); }

struct NextTw {                  // pre-twiddle of the next pass, by output position
  int on, Ln, nsn, stepn;
  float inv_Ln, inv_nsn;
};
__device__ __forceinline__ void pf_store(double2* __restrict__ b, int pos, double2 v, const NextTw& nt,
                                         const double2* __restrict__ tabM) {
  if (nt.on) {
    const int tn = qdiv_small(pos, nt.inv_Ln);
    const int jn = pos - tn * nt.Ln;
    const int k = (jn - qdiv_small(jn, nt.inv_nsn) * nt.nsn) * tn * nt.stepn;   // (jn mod nsn) tn stepn < M
    if (k) v = cmulf(v, __ldg(&tabM[k]));
  }
  b[pos] = v;
}

__device__ __forceinline__ void pfft_build_weights(double2* wt, const double2* __restrict__ tabM, int M, const PfftFactors& F) {
  int off = 0;
  for (int p = 0; p < F.n; ++p) {
    const int R = F.r[p], L = M / R;
    if (!(R & 1)) continue;
    const int H = (R - 1) / 2;
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
      const int t = i / H + 1, u = i - (t - 1) * H + 1;          // [t][u]: the outputs of a group are adjacent
      const double2 v = tabM[((u * t) % R) * L];                 // (cos, -sin) of 2 pi u t / R
      wt[off + i] = make_double2(v.x, -v.y);
    }
    off += H * H;
  }
}

template <int PGP>
__device__ __forceinline__ void pfft_pass_odd(const double2* __restrict__ a, double2* __restrict__ b,
                                              const double2* __restrict__ wt, int M, int R, int ns, float inv_L,
                                              float inv_ns, const NextTw nt, const double2* __restrict__ tabM) {
  const int H = (R - 1) / 2, L = M / R, NG = (H + PGP - 1) / PGP;
  for (int w = threadIdx.x; w < L * NG; w += blockDim.x) {
    const int ug = qdiv_small(w, inv_L);
    const int j = w - ug * L;
    const int u0 = ug * PGP;                                     // pairs u0 + 1 .. u0 + nv
    const int nv = min(PGP, H - u0);
    const double2 x0 = a[j];
    double2 sum0 = x0;
    double2 A[PGP], B[PGP];
#pragma unroll
    for (int g = 0; g < PGP; ++g) { A[g] = make_double2(0.0, 0.0); B[g] = make_double2(0.0, 0.0); }
    const double2* row = wt + u0;
#pragma unroll 1
    for (int t = 1; t <= H; ++t) {
      const double2 xp = a[j + t * L], xm = a[j + (R - t) * L];
      const double2 S = make_double2(xp.x + xm.x, xp.y + xm.y), D = make_double2(xp.x - xm.x, xp.y - xm.y);
      sum0.x += S.x;
      sum0.y += S.y;
#pragma unroll
      for (int g = 0; g < PGP; ++g) {
        if (g < nv) {
          const double2 cs = row[g];
          A[g].x += S.x * cs.x;
          A[g].y += S.y * cs.x;
          B[g].x += D.x * cs.y;
          B[g].y += D.y * cs.y;
        }
      }
      row += H;
    }
    const int jm = j - qdiv_small(j, inv_ns) * ns;
    const int base = (j - jm) * R + jm;
#pragma unroll
    for (int g = 0; g < PGP; ++g) {
      if (g < nv) {
        const int u = u0 + g + 1;
        const double ax = x0.x + A[g].x, ay = x0.y + A[g].y;
        pf_store(b, base + u * ns, make_double2(ax + B[g].y, ay - B[g].x), nt, tabM);
        pf_store(b, base + (R - u) * ns, make_double2(ax - B[g].y, ay + B[g].x), nt, tabM);
      }
    }
    if (ug == 0) pf_store(b, base, sum0, nt, tabM);
  }
}

__device__ __forceinline__ double2* pfft_passes(double2* a, double2* b, const double2* __restrict__ tabM,
                                               const double2* __restrict__ wt, int M, const PfftFactors& F) {
  int ns = 1, off = 0;
  for (int p = 0; p < F.n; ++p) {
    const int R = F.r[p];
    const int L = M / R;
    NextTw nt;
    nt.on = (p + 1 < F.n);
    if (nt.on) {
      const int Rn = F.r[p + 1];
      nt.Ln = M / Rn;
      nt.nsn = ns * R;
      nt.stepn = M / (nt.nsn * Rn);
    } else {
      nt.Ln = 1; nt.nsn = 1; nt.stepn = 0;
    }
    nt.inv_Ln = 1.0f / (float)nt.Ln;
    nt.inv_nsn = 1.0f / (float)nt.nsn;
    const float inv_L = 1.0f / (float)L, inv_ns = 1.0f / (float)ns;
    if (R == 2) {
      for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const double2 x0 = a[j], x1 = a[j + L];
        const int jm = j - qdiv_small(j, inv_ns) * ns;
        const int base = (j - jm) * 2 + jm;
        pf_store(b, base, make_double2(x0.x + x1.x, x0.y + x1.y), nt, tabM);
        pf_store(b, base + ns, make_double2(x0.x - x1.x, x0.y - x1.y), nt, tabM);
      }
    } else if (R == 4) {
      for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const double2 x0 = a[j], x1 = a[j + L], x2 = a[j + 2 * L], x3 = a[j + 3 * L];
        const double2 t0 = make_double2(x0.x + x2.x, x0.y + x2.y), t1 = make_double2(x0.x - x2.x, x0.y - x2.y);
        const double2 t2 = make_double2(x1.x + x3.x, x1.y + x3.y), t3 = make_double2(x1.x - x3.x, x1.y - x3.y);
        const int jm = j - qdiv_small(j, inv_ns) * ns;
        const int base = (j - jm) * 4 + jm;
        pf_store(b, base, make_double2(t0.x + t2.x, t0.y + t2.y), nt, tabM);
        pf_store(b, base + ns, make_double2(t1.x + t3.y, t1.y - t3.x), nt, tabM);          // t1 - i t3
        pf_store(b, base + 2 * ns, make_double2(t0.x - t2.x, t0.y - t2.y), nt, tabM);
        pf_store(b, base + 3 * ns, make_double2(t1.x - t3.y, t1.y + t3.x), nt, tabM);      // t1 + i t3
      }
    } else {
      switch (F.pg[p]) {
        case 1: pfft_pass_odd<1>(a, b, wt + off, M, R, ns, inv_L, inv_ns, nt, tabM); break;
        case 2: pfft_pass_odd<2>(a, b, wt + off, M, R, ns, inv_L, inv_ns, nt, tabM); break;
        default: pfft_pass_odd<3>(a, b, wt + off, M, R, ns, inv_L, inv_ns, nt, tabM); break;
      }
      off += ((R - 1) / 2) * ((R - 1) / 2);
    }
    __syncthreads();
    double2* t = a; a = b; b = t;
    ns *= R;
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
  if (!(m == 1 && M >= 2)) return false;
  // odd radices: output pairs per work item -- the fewest rounds of PB threads, then the least work per item
  for (int p = 0; p < F->n; ++p) {
    const int R = F->r[p];
    F->pg[p] = 1;
    if (!(R & 1)) continue;
    const int H = (R - 1) / 2, L = M / R;
    long long best = -1;
    for (int g = 1; g <= 3; ++g) {                               // 4 pairs per item would spill at 64 registers
      const int items = L * ((H + g - 1) / g);
      const long long rounds = (items + PB - 1) / PB;
      const long long cost = rounds * (6 + 5 * g);               // per t: 2 loads + 4 adds + g x (1 load + 4 FMA)
      if (best < 0 || cost < best) { best = cost; F->pg[p] = g; }
    }
  }
  return true;
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
