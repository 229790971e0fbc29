// zfft.cu -- z-direction cosine transforms of the fast-diagonalisation solve, done as FFTs.
//
// With homogeneous Neumann walls the eigenvectors of A_z are v_k[j] = cos(pi k (2j+1) / (2N))
// (pyaxisymflow/kernels/FastDiagonalisationStokesSolver.py:107-125 finds them with la.eig): the
// forward z transform of the solve is a DCT-II of every row and the backward one a DCT-III.
// For N = 2^p a row fits in shared memory, so each transform is ONE pass over HBM (read a row,
// write a row) instead of an (nr x N) x (N x N) GEMM:
//
//   DCT-II :  x -> permuted real sequence v -> z[n] = v[2n] + i v[2n+1] -> FFT_{N/2} in shared
//             memory -> real-FFT untangling -> quarter-wave rotation -> X[k]
//   DCT-III:  the same steps backwards (inverse FFT through the swap re<->im identity)
//
// The FFT is an in-place Stockham auto-sort: every thread owns 16 complex points per pass, reads
// them all, __syncthreads, then writes its radix-16 (last pass: radix 2/4/8) butterflies to the
// auto-sort positions.  Twiddles come from exactly rounded host tables.
#include <math.h>

#include "axb_common.cuh"

namespace {

constexpr int EPT = 16;  // complex points per thread and pass
constexpr double RH = 0.70710678118654752440;

__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }   // one spare slot per 16: pass-1 stores
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// W16^k = exp(-2 pi i k / 16), k = 0..7
__device__ __forceinline__ double2 w16(int k) {
  constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173, H = 0.70710678118654752440;
  switch (k) {
    case 0: return make_double2(1.0, 0.0);
    case 1: return make_double2(C1, -S1);
    case 2: return make_double2(H, -H);
    case 3: return make_double2(S1, -C1);
    case 4: return make_double2(0.0, -1.0);
    case 5: return make_double2(-S1, -C1);
    case 6: return make_double2(-H, -H);
    default: return make_double2(-C1, -S1);
  }
}

// forward R-point DFT in registers, natural order in and out (decimation in time)
template <int R>
__device__ __forceinline__ void dft(double2 (&v)[R]) {
  if constexpr (R == 2) {
    const double2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else {
    double2 e[R / 2], o[R / 2];
#pragma unroll
    for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    dft<R / 2>(e);
    dft<R / 2>(o);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      double2 t;
      if (k == 0) t = o[0];
      else if (4 * k == R) t = make_double2(o[k].y, -o[k].x);       // times -i
      else t = cmul(o[k], w16(k * (16 / R)));
      v[k] = cadd(e[k], t);
      v[k + R / 2] = csub(e[k], t);
    }
  }
}

// One Stockham pass of radix R over a row of M points held in s (padded indexing).
template <int R>
__device__ __forceinline__ void fft_pass(double2* s, int M, int Ns, int t0, int T, bool live,
                                         const double2* __restrict__ tabM) {
  constexpr int B = EPT / R;
  double2 v[B][R];
  const int stride = M / R;
  if (live) {
#pragma unroll
    for (int b = 0; b < B; ++b)
#pragma unroll
      for (int t = 0; t < R; ++t) v[b][t] = s[pad(t0 + b * T + t * stride)];
  }
  __syncthreads();
  if (live) {
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int j = t0 + b * T;
      const int jm = j & (Ns - 1);
      if (Ns > 1) {
        const int step = jm * (stride / Ns);                        // jm * M / (Ns R)
        if constexpr (R == 16) {
          // w^t = w^(4a) w^b from six table entries (one rounding deep) instead of fifteen loads
          double2 wa[4], wb[4];
#pragma unroll
          for (int q = 1; q < 4; ++q) { wb[q] = tabM[q * step]; wa[q] = tabM[4 * q * step]; }
#pragma unroll
          for (int t = 1; t < R; ++t) {
            const int a = t >> 2, c = t & 3;
            const double2 w = (a == 0) ? wb[c] : (c == 0 ? wa[a] : cmul(wa[a], wb[c]));
            v[b][t] = cmul(v[b][t], w);
          }
        } else {
#pragma unroll
          for (int t = 1; t < R; ++t) v[b][t] = cmul(v[b][t], tabM[t * step]);
        }
      }
      dft<R>(v[b]);
      const int base = (j - jm) * R + jm;
#pragma unroll
      for (int t = 0; t < R; ++t) s[pad(base + t * Ns)] = v[b][t];
    }
  }
  __syncthreads();
}

// The last pass (Ns * R == M) writes every butterfly back to the slots it read, so it needs no
// barrier between the reads and the writes and the butterflies of a thread can run one by one.
template <int R>
__device__ __forceinline__ void fft_last_pass(double2* s, int M, int t0, int T, bool live,
                                              const double2* __restrict__ tabM) {
  constexpr int B = EPT / R;
  const int Ns = M / R;
  if (live) {
#pragma unroll 2
    for (int b = 0; b < B; ++b) {
      const int j = t0 + b * T;                                     // j < Ns
      double2 v[R];
#pragma unroll
      for (int t = 0; t < R; ++t) v[t] = s[pad(j + t * Ns)];
#pragma unroll
      for (int t = 1; t < R; ++t) v[t] = cmul(v[t], tabM[t * j]);   // exp(-2 pi i j t / M)
      dft<R>(v);
#pragma unroll
      for (int t = 0; t < R; ++t) s[pad(j + t * Ns)] = v[t];
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void fft_row(double2* s, int M, int logM, int t0, int T, bool live,
                                        const double2* __restrict__ tabM) {
  int Ns = 1;
  for (int p = 0; p < (logM >> 2); ++p) {
    fft_pass<16>(s, M, Ns, t0, T, live, tabM);
    Ns <<= 4;
  }
  switch (logM & 3) {
    case 1: fft_last_pass<2>(s, M, t0, T, live, tabM); break;
    case 2: fft_last_pass<4>(s, M, t0, T, live, tabM); break;
    case 3: fft_last_pass<8>(s, M, t0, T, live, tabM); break;
    default: break;
  }
}

// tables: [ tabM : M | tabN : M + 1 | tab4N : M + 1 ] as double2,
//   tabM[k] = exp(-2 pi i k / M), tabN[k] = exp(-2 pi i k / N), tab4N[k] = exp(-i pi k / (2N))
struct Tabs {
  const double2 *M, *N, *Q;
};
__device__ __forceinline__ Tabs split_tabs(const double2* t, int M) {
  Tabs r;
  r.M = t;
  r.N = t + M;
  r.Q = t + 2 * M + 1;
  return r;
}

// X[k] = s_k * sum_j x[j] cos(pi k (2j+1) / (2N)),  s_0 = scale0, s_k = scale otherwise
template <bool VEC>
__global__ void __launch_bounds__(512)
    k_dct2_rows(int rows, int N, int logM, int rpc, const double* __restrict__ src, long long ld_src,
                double* __restrict__ dst, long long ld_dst, const double2* __restrict__ tabs, double scale0,
                double scale) {
  extern __shared__ double2 smem[];
  const int M = N >> 1, T = M / EPT;
  const int rl = threadIdx.x / T, t0 = threadIdx.x - rl * T;
  double2* s = smem + (size_t)rl * (M + (M >> 4) + 1);
  const Tabs tb = split_tabs(tabs, M);
  for (int rb = blockIdx.x; rb * rpc < rows; rb += gridDim.x) {
    const int row = rb * rpc + rl;
    const bool live = row < rows;
    if (live) {
      const double* x = src + (long long)row * ld_src;
#pragma unroll
      for (int i = 0; i < EPT / 2; ++i) {
        const int n = t0 + i * T;                                   // n < M/2
        double x0, x1, x2, x3;
        if (VEC) {
          const double2 a = *reinterpret_cast<const double2*>(x + 4 * n);
          const double2 b = *reinterpret_cast<const double2*>(x + 4 * n + 2);
          x0 = a.x; x1 = a.y; x2 = b.x; x3 = b.y;
        } else {
          x0 = x[4 * n]; x1 = x[4 * n + 1]; x2 = x[4 * n + 2]; x3 = x[4 * n + 3];
        }
        s[pad(n)] = make_double2(x0, x2);
        s[pad(M - 1 - n)] = make_double2(x3, x1);
      }
    }
    __syncthreads();
    fft_row(s, M, logM, t0, T, live, tb.M);
    if (live) {
      double* X = dst + (long long)row * ld_dst;
      if (t0 == 0) {
        const double2 z0 = s[0];
        X[0] = (z0.x + z0.y) * scale0;
        X[M] = (z0.x - z0.y) * 0.70710678118654752440 * scale;
      }
#pragma unroll
      for (int i = 0; i < EPT / 2; ++i) {
        const int k = 1 + t0 + i * T;                               // 1 .. M/2
        const double2 zk = s[pad(k)], zm = s[pad(M - k)];
        const double ex = 0.5 * (zk.x + zm.x), ey = 0.5 * (zk.y - zm.y);
        const double2 d = make_double2(0.5 * (zk.x - zm.x), 0.5 * (zk.y + zm.y));
        // one table read per k: exp(-i pi (M-k) / 2N) = exp(-i pi / 4) conj(q), exp(-2 pi i k / N) = q^4
        const double2 q = tb.Q[k];
        const double2 qm = make_double2(RH * (q.x - q.y), -RH * (q.x + q.y));
        const double2 q2 = cmul(q, q);
        const double2 p = cmul(cmul(q2, q2), d);
        const double2 vk = make_double2(ex + p.y, ey - p.x);
        const double2 vm = make_double2(ex - p.y, -ey - p.x);
        const double2 a = cmul(q, vk), b = cmul(qm, vm);
        X[k] = a.x * scale;
        X[N - k] = -a.y * scale;
        X[M - k] = b.x * scale;
        X[M + k] = -b.y * scale;
      }
    }
    __syncthreads();
  }
}

// y[j] = sum_k a[k] cos(pi k (2j+1) / (2N))
template <bool VEC>
__global__ void __launch_bounds__(512)
    k_dct3_rows(int rows, int N, int logM, int rpc, const double* __restrict__ src, long long ld_src,
                double* __restrict__ dst, long long ld_dst, const double2* __restrict__ tabs) {
  extern __shared__ double2 smem[];
  const int M = N >> 1, T = M / EPT;
  const int rl = threadIdx.x / T, t0 = threadIdx.x - rl * T;
  double2* s = smem + (size_t)rl * (M + (M >> 4) + 1);
  const Tabs tb = split_tabs(tabs, M);
  for (int rb = blockIdx.x; rb * rpc < rows; rb += gridDim.x) {
    const int row = rb * rpc + rl;
    const bool live = row < rows;
    if (live) {
      const double* a = src + (long long)row * ld_src;
      if (t0 == 0) {
        const double a0 = a[0], am = a[M] * 0.70710678118654752440;
        s[0] = make_double2(a0 - am, a0 + am);                      // stored swapped (im, re)
      }
#pragma unroll
      for (int i = 0; i < EPT / 2; ++i) {
        const int k = 1 + t0 + i * T;
        const double2 qk = tb.Q[k];
        const double2 qm = make_double2(RH * (qk.x - qk.y), -RH * (qk.x + qk.y));
        const double2 q2 = cmul(qk, qk);
        // B_k = conj(q_k) (a_k - i a_{N-k}) / 2
        const double2 bk = cmul(make_double2(qk.x, -qk.y), make_double2(0.5 * a[k], -0.5 * a[N - k]));
        const double2 bm = cmul(make_double2(qm.x, -qm.y), make_double2(0.5 * a[M - k], -0.5 * a[M + k]));
        const double sx = bk.x + bm.x, sy = bk.y - bm.y;
        const double2 dd = make_double2(bk.x - bm.x, bk.y + bm.y);
        const double2 wn = cmul(q2, q2);
        const double2 q = cmul(make_double2(wn.x, -wn.y), dd);
        s[pad(k)] = make_double2(sy + q.x, sx - q.y);               // swapped
        s[pad(M - k)] = make_double2(-sy + q.x, sx + q.y);
      }
    }
    __syncthreads();
    fft_row(s, M, logM, t0, T, live, tb.M);
    if (live) {
      double* y = dst + (long long)row * ld_dst;
#pragma unroll
      for (int i = 0; i < EPT / 2; ++i) {
        const int n = t0 + i * T;
        const double2 za = s[pad(n)], zb = s[pad(M - 1 - n)];       // swapped: (.y, .x) = (re, im)
        if (VEC) {
          *reinterpret_cast<double2*>(y + 4 * n) = make_double2(za.y, zb.x);
          *reinterpret_cast<double2*>(y + 4 * n + 2) = make_double2(za.x, zb.y);
        } else {
          y[4 * n] = za.y; y[4 * n + 1] = zb.x; y[4 * n + 2] = za.x; y[4 * n + 3] = zb.y;
        }
      }
    }
    __syncthreads();
  }
}

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

}  // namespace

int launch_dct_rows(int inverse, int rows, int N, const double* src, long long ld_src, double* dst, long long ld_dst,
                    const double* tabs, double scale0, double scale, cudaStream_t st) {
  if (rows < 1 || !src || !dst || !tabs || ld_src < N || ld_dst < N) return AXB_EINVAL;
  if (N < 64 || N > 16384 || (N & (N - 1))) return AXB_EINVAL;
  const int M = N / 2, T = M / EPT;
  int rpc = 1;
  while (rpc * T < 128) rpc *= 2;
  const size_t smem = (size_t)rpc * (M + (M >> 4) + 1) * sizeof(double2);
  const bool vec = axb_al16(src) && axb_al16(dst) && (ld_src % 2 == 0) && (ld_dst % 2 == 0);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_dct2_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_dct2_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_dct3_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_dct3_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int per_sm = (int)((227u * 1024u) / (smem + 1024));
  const int resident = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const int nblocks = (rows + rpc - 1) / rpc;
  const int grid = nblocks < sms * resident ? nblocks : sms * resident;
  const int threads = rpc * T;
  const double2* tb = reinterpret_cast<const double2*>(tabs);
  const int logM = ilog2(M);
  if (!inverse) {
    if (vec) k_dct2_rows<true><<<grid, threads, smem, st>>>(rows, N, logM, rpc, src, ld_src, dst, ld_dst, tb, scale0, scale);
    else k_dct2_rows<false><<<grid, threads, smem, st>>>(rows, N, logM, rpc, src, ld_src, dst, ld_dst, tb, scale0, scale);
  } else {
    if (vec) k_dct3_rows<true><<<grid, threads, smem, st>>>(rows, N, logM, rpc, src, ld_src, dst, ld_dst, tb);
    else k_dct3_rows<false><<<grid, threads, smem, st>>>(rows, N, logM, rpc, src, ld_src, dst, ld_dst, tb);
  }
  AXB_LAUNCHED();
  return (int)cudaGetLastError();
}

extern "C" {

int axb_dct2_rows(int rows, int n, const double* src, int64_t ld_src, double* dst, int64_t ld_dst, const double* tables,
                  double scale0, double scale, axb_stream_t s) {
  return launch_dct_rows(0, rows, n, src, ld_src, dst, ld_dst, tables, scale0, scale, (cudaStream_t)s);
}
int axb_dct3_rows(int rows, int n, const double* src, int64_t ld_src, double* dst, int64_t ld_dst, const double* tables,
                  axb_stream_t s) {
  return launch_dct_rows(1, rows, n, src, ld_src, dst, ld_dst, tables, 1.0, 1.0, (cudaStream_t)s);
}

}  // extern "C"
