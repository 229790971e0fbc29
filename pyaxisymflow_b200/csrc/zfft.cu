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
#include <cuda.h>   // CUtensorMap (types only; the encoder comes through cudaGetDriverEntryPoint)
#include <math.h>
#include <stdlib.h>
#include <string.h>

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

// The shared-memory kernels have no room to stage the next row on chip (two CTAs of 70 KB at N = 8192), so its HBM
// read is started as an L2 prefetch -- one bulk instruction per row, no registers, no shared memory -- while the current
// row is transformed; the load phase of the next iteration then runs at L2 latency.
__device__ __forceinline__ void l2_prefetch_row(const double* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(p), "r"(bytes) : "memory");
}

// X[k] = s_k * sum_j x[j] cos(pi k (2j+1) / (2N)),  s_0 = scale0, s_k = scale otherwise
template <bool VEC>
__global__ void __launch_bounds__(512)
    k_dct2_rows(int rows, int N, int logM, int rpc, const double* __restrict__ src, long long ld_src,
                double* __restrict__ dst, long long ld_dst, const double2* __restrict__ tabs, double scale0,
                double scale, int prefetch) {
  extern __shared__ double2 smem[];
  const int M = N >> 1, T = M / EPT;
  const int rl = threadIdx.x / T, t0 = threadIdx.x - rl * T;
  double2* s = smem + (size_t)rl * (M + (M >> 4) + 1);
  const Tabs tb = split_tabs(tabs, M);
  for (int rb = blockIdx.x; rb * rpc < rows; rb += gridDim.x) {
    const int row = rb * rpc + rl;
    const bool live = row < rows;
    if (VEC && prefetch && t0 == 0) {
      const long long nrow = (long long)row + (long long)gridDim.x * rpc;
      if (nrow < rows) l2_prefetch_row(src + nrow * ld_src, (unsigned)N * 8u);
    }
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
                double* __restrict__ dst, long long ld_dst, const double2* __restrict__ tabs, int prefetch) {
  extern __shared__ double2 smem[];
  const int M = N >> 1, T = M / EPT;
  const int rl = threadIdx.x / T, t0 = threadIdx.x - rl * T;
  double2* s = smem + (size_t)rl * (M + (M >> 4) + 1);
  const Tabs tb = split_tabs(tabs, M);
  for (int rb = blockIdx.x; rb * rpc < rows; rb += gridDim.x) {
    const int row = rb * rpc + rl;
    const bool live = row < rows;
    if (VEC && prefetch && t0 == 0) {
      const long long nrow = (long long)row + (long long)gridDim.x * rpc;
      if (nrow < rows) l2_prefetch_row(src + nrow * ld_src, (unsigned)N * 8u);
    }
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

// ---------------------------------------------------------------------------------------------
// Register-resident variant for M = 2 * 16^P (N = 64, 1024, 16384): the 16 points a thread owns stay
// in registers from pass to pass and shared memory is only the exchange medium -- real parts
// first, then imaginary parts, so the exchange buffer is half a row.  The freed shared memory
// holds the NEXT row, fetched by a bulk async copy (TMA, mbarrier complete_tx) while the current one
// is transformed: the HBM read is off the critical path.  The last radix-2 stage of the FFT is
// folded into the untangling step (DCT-II) / the output step (DCT-III), which then work on
// index quadruples {k, k + M/2, M/2 - k, M - k}.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int pd(int p) { return p + (p >> 4); }
__device__ __forceinline__ unsigned rr_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rr_mb_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "RR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra RR_DONE;\n"
      "bra RR_WAIT;\n"
      "RR_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Q[k] = exp(-i pi k / 2N), 0 <= k <= M/2, from the shared-memory copy of its first M/8 + 1 entries:
// Q[M/2 - k] = exp(-i pi / 8) conj(Q[k]),  Q[M/4 - k] = exp(-i pi / 16) conj(Q[k])
// PAD: the table is stored with one spare slot per 32 entries (entry t at t + (t >> 5)).  The DCT-III of the warp-local
// kernel reads it at k = n1 + 32 a + 512 i, 32 entries = 512 bytes apart from lane to lane -- all 16 lanes of a half-warp
// in the same banks (ncu: 32 wavefronts per LDS.128 instead of 4, 14.7 M of that kernel's 23 M conflicts); with the
// spare slots the stride is 33 entries and a quarter-warp covers all banks.
template <bool PAD = false>
__device__ __forceinline__ double2 q_at(const double2* __restrict__ tq, int k, int M) {
  const bool r1 = k > (M >> 2);
  const int k1 = r1 ? (M >> 1) - k : k;
  const bool r2 = k1 > (M >> 3);
  const int k2 = r2 ? (M >> 2) - k1 : k1;
  double2 q = tq[PAD ? k2 + (k2 >> 5) : k2];
  if (r2) q = cmul(make_double2(0.98078528040323044913, -0.19509032201612826785), make_double2(q.x, -q.y));
  if (r1) q = cmul(make_double2(0.92387953251128675613, -0.38268343236508977173), make_double2(q.x, -q.y));
  return q;
}

// untangle one k (1 <= k <= M/2): zk = Z[k], zm = Z[M - k], q = exp(-i pi k / 2N), tn = q^4 = exp(-2 pi i k / N)
__device__ __forceinline__ void dct2_emit(double* __restrict__ X, int k, int M, int N, double2 zk, double2 zm, double2 q,
                                          double2 tn, double scale) {
  const double ex = 0.5 * (zk.x + zm.x), ey = 0.5 * (zk.y - zm.y);
  const double2 d = make_double2(0.5 * (zk.x - zm.x), 0.5 * (zk.y + zm.y));
  const double2 qm = make_double2(RH * (q.x - q.y), -RH * (q.x + q.y));
  const double2 p = cmul(tn, d);
  const double2 vk = make_double2(ex + p.y, ey - p.x);
  const double2 vm = make_double2(ex - p.y, -ey - p.x);
  const double2 a = cmul(q, vk), b = cmul(qm, vm);
  X[k] = a.x * scale;
  X[N - k] = -a.y * scale;
  X[M - k] = b.x * scale;
  X[M + k] = -b.y * scale;
}
// inverse of the above: spectrum entries a[k], a[N-k], a[M-k], a[M+k] -> Z[k], Z[M-k] (stored swapped)
__device__ __forceinline__ void dct3_pair(double ak, double ank, double amk, double apk, double2 qk, double2& zk,
                                          double2& zm) {
  const double2 qm = make_double2(RH * (qk.x - qk.y), -RH * (qk.x + qk.y));
  const double2 q2 = cmul(qk, qk);
  const double2 bk = cmul(make_double2(qk.x, -qk.y), make_double2(0.5 * ak, -0.5 * ank));
  const double2 bm = cmul(make_double2(qm.x, -qm.y), make_double2(0.5 * amk, -0.5 * apk));
  const double sx = bk.x + bm.x, sy = bk.y - bm.y;
  const double2 dd = make_double2(bk.x - bm.x, bk.y + bm.y);
  const double2 wn = cmul(q2, q2);
  const double2 q = cmul(make_double2(wn.x, -wn.y), dd);
  zk = make_double2(sy + q.x, sx - q.y);
  zm = make_double2(-sy + q.x, sx + q.y);
}

template <bool INV>
__global__ void __launch_bounds__(512, 1)
    k_dct_rows_rr(const __grid_constant__ CUtensorMap tmS, int split, int rows, int N, int logM, int rpc,
                  const double* __restrict__ src, long long ld_src, double* __restrict__ dst, long long ld_dst,
                  const double2* __restrict__ tabs, double scale0, double scale) {
  extern __shared__ __align__(128) unsigned char rr_smem[];
  const int M = N >> 1, H = M >> 1, T = M >> 4;
  const int rl = threadIdx.x >> (logM - 4), j = threadIdx.x - rl * T;
  const int EXS = M + (M >> 4) + 16;                                   // doubles per row of the exchange buffer
  double* stage_all = reinterpret_cast<double*>(rr_smem);
  double* ex_all = stage_all + (size_t)rpc * N;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(ex_all + (size_t)rpc * EXS);
  const double* st = stage_all + (size_t)rl * N;
  double* ex = ex_all + (size_t)rl * EXS;
  const Tabs tb = split_tabs(tabs, M);
  const int P = (logM - 1) >> 2;                                       // radix-16 passes
  const unsigned bar_a = rr_u32(bar);
  // on-chip copies of the tables: Q[0 .. M/8] and, per pass p >= 1, the base twiddles w^1 and w^4 of its
  // 16^p residues (everything else is derived by conjugate reflections, squarings and products)
  double2* tq = reinterpret_cast<double2*>(bar + 2);
  double2* tw = tq + (M >> 3) + 1;
  for (int i = threadIdx.x; i <= (M >> 3); i += blockDim.x) tq[i] = tb.Q[i];
  {
    int off = 0;
    for (int p = 1; p < P; ++p) {
      const int ns = 1 << (4 * p);
      for (int i = threadIdx.x; i < ns; i += blockDim.x) {
        const int step = i * (T >> (4 * p));
        tw[off + i] = tb.M[step];
        tw[off + ns + i] = tb.M[4 * step];
      }
      off += 2 * ns;
    }
  }
  const double2 wm1 = tb.M[1];                                         // exp(-2 pi i / M)

  auto issue = [&](int rb) {                                           // thread 0: fetch row block rb
    const int r0 = rb * rpc;
    const int nl = min(rpc, rows - r0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_a), "r"(nl * N * 8) : "memory");
    for (int q = 0; q < nl; ++q) {
      const double* g = src + (long long)(r0 + q) * ld_src;
      const unsigned d = rr_u32(stage_all + (size_t)q * N);
      if (split) {
        // the row viewed as quadruples (x[4n .. 4n+3]): two tensor loads with a 2-element inner box stage
        // A[n] = (x[4n], x[4n+1]) in st[0 .. N/2) and B[n] = (x[4n+2], x[4n+3]) in st[N/2 .. N)
#pragma unroll
        for (int par = 0; par < 2; ++par)
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(
                  d + par * (N * 4)),
              "l"(reinterpret_cast<unsigned long long>(&tmS)), "r"(2 * par), "r"(0), "r"(0), "r"(r0 + q), "r"(bar_a)
              : "memory");
        continue;
      }
      for (int c = 0; c < N * 8; c += 32768) {                         // <= 32 KB per copy
        const int bytes = min(32768, N * 8 - c);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d + c),
            "l"(reinterpret_cast<const char*>(g) + c), "r"(bytes), "r"(bar_a)
            : "memory");
      }
    }
  };
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (blockIdx.x * rpc < rows) issue(blockIdx.x);
  }
  __syncthreads();

  int it = 0;
  for (int rb = blockIdx.x; rb * rpc < rows; rb += gridDim.x, ++it) {
    const int row = rb * rpc + rl;
    const bool live = row < rows;
    rr_mb_wait(bar_a, it & 1);
    double2 v[16];
    double2 mir[8];
    if (!INV && split) {
      // staged as A[n] = (x[4n], x[4n+1]), B[n] = (x[4n+2], x[4n+3]): conflict-free 128-bit reads
      const double2* qa = reinterpret_cast<const double2*>(st);
      const double2* qb = reinterpret_cast<const double2*>(st + (N >> 1));
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int n = j + t * T;                                       // < M/2
        v[t] = make_double2(qa[n].x, qb[n].x);
      }
#pragma unroll
      for (int t = 8; t < 16; ++t) {
        const int n = M - 1 - (j + t * T);                             // mirrored half
        v[t] = make_double2(qb[n].y, qa[n].y);
      }
    } else if (!INV) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int n = j + t * T;                                       // < M/2
        v[t] = make_double2(st[4 * n], st[4 * n + 2]);
      }
#pragma unroll
      for (int t = 8; t < 16; ++t) {
        const int n = M - 1 - (j + t * T);                             // mirrored half
        v[t] = make_double2(st[4 * n + 3], st[4 * n + 1]);
      }
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int k = j + t * T;                                       // < M/2
        if (k == 0) {
          const double a0 = st[0], am = st[M] * RH;
          v[0] = make_double2(a0 - am, a0 + am);                       // Z[0], stored swapped (im, re)
          double2 dummy;
          dct3_pair(st[H], st[N - H], st[M - H], st[M + H], make_double2(0.92387953251128675613, -0.38268343236508977173),
                    mir[0], dummy);                                    // Z[M/2]
        } else {
          dct3_pair(st[k], st[N - k], st[M - k], st[M + k], q_at(tq, k, M), v[t], mir[t]);
        }
      }
    }
    __syncthreads();                                                   // the staged rows have been consumed
    if (threadIdx.x == 0 && (rb + (int)gridDim.x) * rpc < rows) issue(rb + gridDim.x);
    if (INV) {
      // Z[M - k] belongs to the thread that owns slot (M - k) / T of butterfly (M - k) % T
      double2* ex2 = reinterpret_cast<double2*>(ex);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int k = j + t * T;
        const int np = (k == 0) ? H : M - k;
        ex2[((np >> (logM - 4)) - 8) * T + (np & (T - 1))] = mir[t];
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 8; ++u) v[8 + u] = ex2[u * T + j];
      __syncthreads();
    }
    // ---- P radix-16 Stockham passes; outputs of pass p sit at base + t * Ns
    int Ns = 1, base = j * 16, twoff = 0;
#pragma unroll 1
    for (int p = 0; p < P; ++p) {
      const int jm = j & (Ns - 1);
      if (Ns > 1) {
        double2 wa[4], wb[4];
        wb[1] = tw[twoff + jm];
        wa[1] = tw[twoff + Ns + jm];
        wb[2] = cmul(wb[1], wb[1]);
        wb[3] = cmul(wb[2], wb[1]);
        wa[2] = cmul(wa[1], wa[1]);
        wa[3] = cmul(wa[2], wa[1]);
        twoff += 2 * Ns;
#pragma unroll
        for (int t = 1; t < 16; ++t) {
          const int a = t >> 2, c = t & 3;
          const double2 w = (a == 0) ? wb[c] : (c == 0 ? wa[a] : cmul(wa[a], wb[c]));
          v[t] = cmul(v[t], w);
        }
      }
      dft<16>(v);
      base = (j - jm) * 16 + jm;
      if (p + 1 < P) {                                                 // exchange into the next pass' layout
#pragma unroll
        for (int t = 0; t < 16; ++t) ex[pd(base + t * Ns)] = v[t].x;
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t].x = ex[pd(j + t * T)];
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 16; ++t) ex[pd(base + t * Ns)] = v[t].y;
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t].y = ex[pd(j + t * T)];
        __syncthreads();
        Ns <<= 4;
      }
    }
    // ---- last exchange: quadruples for the folded radix-2 stage
    double2 e0 = make_double2(0.0, 0.0), eh = make_double2(0.0, 0.0);
    const int off = INV ? 0 : 1;                                       // DCT-II quads start at k = 1
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int t = 0; t < 16; ++t) ex[pd(base + t * Ns)] = half ? v[t].y : v[t].x;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = off + j + i * T;
        const int hk = INV ? H - 1 - k : H - k, mk = INV ? M - 1 - k : M - k;
        const double r0 = ex[pd(k)], r1 = ex[pd(k + H)], r2 = ex[pd(hk)], r3 = ex[pd(mk)];
        if (half) { v[4 * i].y = r0; v[4 * i + 1].y = r1; v[4 * i + 2].y = r2; v[4 * i + 3].y = r3; }
        else { v[4 * i].x = r0; v[4 * i + 1].x = r1; v[4 * i + 2].x = r2; v[4 * i + 3].x = r3; }
      }
      if (!INV && j == 0) {
        if (half) { e0.y = ex[0]; eh.y = ex[pd(H)]; }
        else { e0.x = ex[0]; eh.x = ex[pd(H)]; }
      }
      __syncthreads();
    }
    if (live) {
      double* out = dst + (long long)row * ld_dst;
      if (!INV) {
        if (j == 0) {
          const double2 z0 = cadd(e0, eh), zh = csub(e0, eh);
          out[0] = (z0.x + z0.y) * scale0;
          out[M] = (z0.x - z0.y) * RH * scale;
          dct2_emit(out, H, M, N, zh, zh, make_double2(0.92387953251128675613, -0.38268343236508977173),
                    make_double2(0.0, -1.0), scale);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = 1 + j + i * T;                                 // 1 .. M/4
          const double2 q = q_at(tq, k, M);
          const double2 qq = cmul(q, q);
          const double2 tn = cmul(qq, qq);                              // exp(-2 pi i k / N)
          const double2 w = cmul(tn, tn);                               // exp(-2 pi i k / M)
          const double2 wb = cmul(w, v[4 * i + 1]);
          const double2 cwd = cmul(make_double2(w.x, -w.y), v[4 * i + 3]);
          const double2 z_k = cadd(v[4 * i], wb), z_kh = csub(v[4 * i], wb);          // Z[k], Z[k + M/2]
          const double2 z_hk = csub(v[4 * i + 2], cwd), z_mk = cadd(v[4 * i + 2], cwd);   // Z[M/2 - k], Z[M - k]
          // exp(-i pi (M/2 - k) / 2N) = exp(-i pi / 8) conj(q); its 4th power is -i conj(tn)
          const double2 q2 = cmul(make_double2(0.92387953251128675613, -0.38268343236508977173),
                                  make_double2(q.x, -q.y));
          dct2_emit(out, k, M, N, z_k, z_mk, q, tn, scale);
          dct2_emit(out, H - k, M, N, z_hk, z_kh, q2, make_double2(-tn.y, -tn.x), scale);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = j + i * T;                                     // 0 .. M/4 - 1
          const int n2 = H - 1 - n;
          // exp(-2 pi i n / M) = Q[n]^8; exp(-2 pi i (M/2 - 1 - n) / M) = -conj(exp(-2 pi i (n + 1) / M))
          const double2 q = q_at(tq, n, M);
          const double2 qq = cmul(q, q);
          const double2 q4 = cmul(qq, qq);
          const double2 wn = cmul(q4, q4);
          const double2 w1 = cmul(wn, wm1);
          const double2 wb = cmul(wn, v[4 * i + 1]);
          const double2 wd = cmul(make_double2(-w1.x, w1.y), v[4 * i + 3]);
          const double2 z_n = cadd(v[4 * i], wb), z_nh = csub(v[4 * i], wb);          // z[n], z[n + M/2]
          const double2 z_c = cadd(v[4 * i + 2], wd), z_m = csub(v[4 * i + 2], wd);   // z[M/2-1-n], z[M-1-n]
          // stored swapped: (.y, .x) = (re, im)
          *reinterpret_cast<double2*>(out + 4 * n) = make_double2(z_n.y, z_m.x);
          *reinterpret_cast<double2*>(out + 4 * n + 2) = make_double2(z_n.x, z_m.y);
          *reinterpret_cast<double2*>(out + 4 * n2) = make_double2(z_c.y, z_nh.x);
          *reinterpret_cast<double2*>(out + 4 * n2 + 2) = make_double2(z_c.x, z_nh.y);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-local variant for N = 16384 (M = 8192 = 32 x 256), modelled line by line in tools/dct_w_model.py.
// The register-resident kernel above exchanges its 16 points per thread through shared memory after EVERY radix-16
// pass with block-wide barriers, so all 16 warps sit in the same phase -- butterflies (FP64 pipe), then exchange
// (shared-memory pipe), then barrier -- and neither pipe is busy half of the time (ncu: FP64 pipe 36 %).  Here the
// row is split m = n1 + 32 n2:
//   phase 1  the 256-point FFT over n2 of residue n1 is done by ONE HALF-WARP (two radix-16 passes, exchange X1
//            inside the half-warp: __syncwarp only);
//   phase 2  one block-wide exchange X2 and a radix-16 over d (n1 = c + 2 d) give E / O, the two half-length FFTs;
//   phase 3  radix 2 + real-FFT untangling + quarter-wave rotation work on quadruples {k, k+M/2, M/2-k, M-k} that
//            live in 4 threads of one warp (exchange X3: __syncwarp only).
// Everything between two X2 phases -- three DFT-16s, the twiddles, X3, the untangling, the stores, the next row's
// loads and X1 -- is free of block barriers, so the warps drift apart and the FP64 pipe of one overlaps the
// shared-memory traffic of another.  The next row is staged by cp.async (16-byte chunks, XOR-swizzled so that the
// stride-32-sector reads of phase 1 are conflict free; completion through an mbarrier) while the current one is
// transformed.  All exchanges are conflict free per 16 lanes (XOR placements, no padding): the exchange buffer is
// exactly one row of doubles (real parts, then imaginary parts).
// ---------------------------------------------------------------------------------------------
constexpr int WN = 16384, WM = WN / 2, WH = WM / 2;

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}

// `opaque(t)` hides a loop-invariant value from the optimiser: everything derived from it afterwards is recomputed
// where it is used (integer ALU work, the FP64 and shared-memory pipes are the busy ones) instead of being kept in
// registers across the whole row loop -- the 16 complex points of a thread already take 64 of its 128 registers.
__device__ __forceinline__ int opaque(int t) {
  asm volatile("" : "+r"(t));
  return t;
}

// Stage layout.  The row arrives by TMA (one tensor map over (16 doubles, 8 x 1 KB, 8 x 128 B, 16 x 8 KB, rows), 128-byte
// swizzle): source byte offset B = 8192 i3 + 1024 i1 + 128 i2 + 16 cc + .. lands in shared-memory line
// 64 i3 + 8 i2 + i1 at chunk cc ^ i1.  Phase 1 reads sectors 32 apart (i1 = a & 7 differs from lane to lane), so the
// XOR spreads the 16 lanes of a half-warp over all banks (tools/dct_w_model.py counts the conflicts: none).
__device__ __forceinline__ int stage_addr8(int s, int h, int word) {     // sector s (32 B), half h, 8-byte word
  const int i1 = (s >> 5) & 7, i2 = (s >> 2) & 7, i3 = s >> 8;
  return i3 * 1024 + i2 * 128 + i1 * 16 + 2 * (((2 * s + h) & 7) ^ i1) + word;
}

// DCT-III stage: the spectrum is read at k, N-k, M-k, M+k with k = n1 + 32 (a + 16 b), 32 DOUBLES apart from lane to
// lane, so the tensor map is (16 doubles, 8 x 256 B, 2 x 128 B, 64 x 2 KB, rows): source offset
// 8 k = 2048 i3 + 256 i1 + 128 i2 + 16 cc + .. lands in line 16 i3 + 8 i2 + i1 at chunk cc ^ i1 (i1 = a & 7).
// (All four indices have the parity of n1, and so have both half-warps of a warp: 2-way conflicts remain.)
__device__ __forceinline__ int stage3_addr8(int k) {
  const int cc = (k >> 1) & 7, i2 = (k >> 4) & 1, i1 = (k >> 5) & 7, i3 = k >> 8;
  return (16 * i3 + 8 * i2 + i1) * 16 + 2 * (cc ^ i1) + (k & 1);
}

template <bool INV>
__global__ void __launch_bounds__(512, 1)
    k_dct_rows_w(const __grid_constant__ CUtensorMap tmS, int rows, const double* __restrict__ src, long long ld_src,
                 double* __restrict__ dst, long long ld_dst, const double2* __restrict__ tabs, double scale0,
                 double scale, unsigned skew_ns) {
  extern __shared__ __align__(1024) unsigned char w_smem[];
  double* S = reinterpret_cast<double*>(w_smem);                         // staged row, swizzled 16-byte chunks
  double* XB = S + WN;                                                   // exchange buffer: WM doubles
  double2* Stab = reinterpret_cast<double2*>(XB + WM);                   // W_256^(d ka), 16 x 16
  double2* tq = Stab + 256;                                              // Q[0 .. M/8]
  double2* T4 = tq + (WM >> 3) + 1 + 33;                                 // W_4096^d, d < 16; [16] = exp(-2 pi i / M)
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(T4 + 17);
  const Tabs tb = split_tabs(tabs, WM);
  const int tid0 = threadIdx.x;
  const unsigned bar_a = rr_u32(bar);
  // ---- tables on chip
  for (int i = tid0; i < 256; i += 512) Stab[i] = tb.M[(32 * (i >> 4) * (i & 15)) & (WM - 1)];
  for (int i = tid0; i <= (WM >> 3); i += 512) tq[i + (i >> 5)] = tb.Q[i];   // padded: q_at<true>
  if (tid0 < 16) T4[tid0] = tb.M[2 * tid0];
  if (tid0 == 16) T4[16] = tb.M[1];

  const unsigned s_base = rr_u32(S);
  auto issue = [&](int row) {                                            // thread 0: fetch one row, 4 x 32 KB boxes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_a), "r"(WN * 8) : "memory");
#pragma unroll
    for (int q = 0; q < 4; ++q)
      asm volatile(
          "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::
              "r"(s_base + 32768u * q),
          "l"(reinterpret_cast<unsigned long long>(&tmS)), "r"(0), "r"(0), "r"(0), "r"((INV ? 16 : 4) * q), "r"(row), "r"(bar_a)
          : "memory");
  };
  if (tid0 == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if ((int)blockIdx.x < rows) issue(blockIdx.x);
  }
  __syncthreads();

  int it = 0;
#pragma unroll 1
  for (int row = blockIdx.x; row < rows; row += gridDim.x, ++it) {
    double2 v[16];
    {
      // ---- phase-1 roles: half-warp hw of warp w does the 256-point FFT of residue n1
      const int tid = opaque(tid0), w = tid >> 5, lane = tid & 31;
      const int hw = (lane >> 3) & 1;
      const int j = (lane & 7) | ((lane >> 4) << 3);
      // DCT-II : n1 = w | 31 - w (sectors m and M-1-m share a warp), hw1 mirrored (a = 15 - j) and rotated by 8
      // DCT-III: n1 = w | 32 - w (Z[k] and Z[M-k] come as pairs; warp 0: 0 | 16), natural roles
      int n1 = INV ? (hw ? 32 - w : w) : (hw ? 31 - w : w);
      if (INV && w == 0 && hw) n1 = 16;
      const int a = (!INV && hw) ? 15 - j : j;
      const int kb = (!INV && hw) ? (j ^ 8) : j;
      const int x1x = (INV && hw) ? 8 : 0;                               // X1 placement of the second half-warp
      const int d = n1 >> 1;
      // tw1 base: W_4096^(16 a + d) = W_256^a W_4096^d, negated when the slots are rotated by 8 ((-1)^kb)
      double2 w1 = cmul(Stab[16 + a], T4[d]);
      if (!INV && hw) { w1.x = -w1.x; w1.y = -w1.y; }
      double2 w4 = cmul(w1, w1);
      w4 = cmul(w4, w4);
      if constexpr (!INV) {
        // stage addresses (8-byte words): slots 0..7 then 8..15, see the model (stage_word)
        int sA_re, sA_im, sB_re, sB_im;
        {
          const int s0 = n1 + 32 * a;                                    // sectors with b < 8: words 0 (re), 2 (im)
          const int lo_re = stage_addr8(s0, 0, 0), lo_im = stage_addr8(s0, 1, 0);
          const int s1 = (31 - n1) + 32 * (15 - a);                      // mirrored sectors: words 3 (re), 1 (im)
          const int hi_re = stage_addr8(s1, 1, 1), hi_im = stage_addr8(s1, 0, 1);
          if (!hw) { sA_re = lo_re; sA_im = lo_im; sB_re = hi_re + 7 * 2048; sB_im = hi_im + 7 * 2048; }
          else { sA_re = hi_re + 7 * 2048; sA_im = hi_im + 7 * 2048; sB_re = lo_re; sB_im = lo_im; }
        }
        const int stepA = hw ? -2048 : 2048;                             // slot i < 8: + i stepA; i >= 8: - (i-8) stepA
        rr_mb_wait(bar_a, it & 1);
        // ---- phase 1a: 16 points z[n1 + 32 (a + 16 b)]
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = make_double2(S[sA_re + i * stepA], S[sA_im + i * stepA]);
#pragma unroll
        for (int i = 8; i < 16; ++i) v[i] = make_double2(S[sB_re - (i - 8) * stepA], S[sB_im - (i - 8) * stepA]);
      } else {
        // ---- phase 1a: Z[k] for the own slots b < 8 (k = n1 + 32 a + 512 b < M/2) from the quadruple
        //      a[k], a[N-k], a[M-k], a[M+k]; the same quadruple gives Z[M-k], slot 15 - b of the thread that holds
        //      residue 32 - n1 at a' = 15 - a (warp 0: residue 0 pairs n2 <-> 256 - n2, residue 16 n2 <-> 255 - n2)
        const int k0 = n1 + 32 * a;
        const int pk = stage3_addr8(k0), pnk = stage3_addr8(WN - k0), pmk = stage3_addr8(WM - k0),
                  ppk = stage3_addr8(WM + k0);                           // slot i: +-512 i
        const bool sp0 = (k0 == 0);                                      // the thread that holds Z[0] and Z[M/2]
        int dh = hw ^ 1, dj = 15 - j;
        if (w == 0) { dh = hw; dj = hw ? 15 - j : ((16 - j) & 15); }
        double2* XP = reinterpret_cast<double2*>(XB) + 256 * w;          // pair exchange: one warp's 256 complex
        const int dst0 = dh * 128 + (dj ^ (8 * dh)) + (sp0 ? 16 : 0);    // slot 15 - i (+ 1 for residue 0, a = 0)
        rr_mb_wait(bar_a, it & 1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = k0 + 512 * i;
          double2 zm;
          if (i == 0 && sp0) {
            const double a0 = S[stage3_addr8(0)], am = S[stage3_addr8(WM)] * RH;
            v[0] = make_double2(a0 - am, a0 + am);                       // Z[0], stored swapped (im, re)
            double2 dummy;
            dct3_pair(S[stage3_addr8(WH)], S[stage3_addr8(WN - WH)], S[stage3_addr8(WM - WH)], S[stage3_addr8(WM + WH)],
                      make_double2(0.92387953251128675613, -0.38268343236508977173), zm, dummy);   // Z[M/2]: slot 8
            XP[0] = zm;
          } else {
            dct3_pair(S[pk + 512 * i], S[pnk - 512 * i], S[pmk - 512 * i], S[ppk + 512 * i], q_at<true>(tq, k, WM), v[i], zm);
            XP[dst0 + 16 * (7 - i)] = zm;
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 8; i < 16; ++i) v[i] = XP[hw * 128 + 16 * (i - 8) + (j ^ (8 * hw))];
        __syncwarp();
      }
      dft<16>(v);
      {                                                                  // tw1: v[t] *= w1^t
        double2 wa[4], wb[4];
        wb[1] = w1;
        wa[1] = w4;
        wb[2] = cmul(wb[1], wb[1]);
        wb[3] = cmul(wb[2], wb[1]);
        wa[2] = cmul(wa[1], wa[1]);
        wa[3] = cmul(wa[2], wa[1]);
#pragma unroll
        for (int t = 1; t < 16; ++t) {
          const int hi = t >> 2, lo = t & 3;
          const double2 ww = (hi == 0) ? wb[lo] : (lo == 0 ? wa[hi] : cmul(wa[hi], wb[lo]));
          v[t] = cmul(v[t], ww);
        }
      }
      // ---- X1: 16 x 16 transpose inside the half-warp (real parts, then imaginary parts)
      const int x1w = 512 * w + 256 * hw + 16 * a;                       // element (a, k) at + (k ^ a)
      const int x1r = 512 * w + 256 * hw;                                // read (aa, kb) at + 16 aa + (kb ^ aa)
      const int ax = a ^ x1x, kx = kb ^ x1x;
#pragma unroll
      for (int k = 0; k < 16; ++k) XB[x1w + (k ^ ax)] = v[k].x;
      __syncwarp();
#pragma unroll
      for (int aa = 0; aa < 16; ++aa) v[aa].x = XB[x1r + 16 * aa + (kx ^ aa)];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 16; ++k) XB[x1w + (k ^ ax)] = v[k].y;
      __syncwarp();
#pragma unroll
      for (int aa = 0; aa < 16; ++aa) v[aa].y = XB[x1r + 16 * aa + (kx ^ aa)];
      dft<16>(v);
#pragma unroll
      for (int t = 1; t < 16; ++t) v[t] = cmul(v[t], Stab[16 * d + t]);  // tw2 (broadcast reads)
    }
    __syncthreads();                        // A: every warp has left the stage and its X1 region
    if (tid0 == 0 && row + (int)gridDim.x < rows) issue(row + gridDim.x);
    {
      // ---- X2: block-wide, A(n1, k2) = 256 n1 + (k2 ^ sigma(n1)); readers are the phase-3 threads
      const int tid = opaque(tid0), w = tid >> 5, lane = tid & 31;
      const int hw = (lane >> 3) & 1;
      const int j = (lane & 7) | ((lane >> 4) << 3);
      int n1 = INV ? (hw ? 32 - w : w) : (hw ? 31 - w : w);
      if (INV && w == 0 && hw) n1 = 16;
      const int kb = (!INV && hw) ? (j ^ 8) : j;
      const int sg = 8 * ((n1 & 1) ^ (n1 >> 4));
      const int x2w = 256 * n1 + (kb ^ sg);                              // + 16 q
      const int mu = lane >> 4, c = (lane >> 3) & 1, G = lane & 7;
      // DCT-II residues pair as k2 <-> 256 - k2 (0 and 128 with themselves), DCT-III as k2 <-> 255 - k2
      int k2 = mu ? (INV ? 255 : 256) - 8 * w - G : 8 * w + G;
      if (!INV && w == 0 && G == 0 && mu) k2 = 128;
      const int x2r_lo = 256 * c + (k2 ^ (8 * c)), x2r_hi = 256 * c + (k2 ^ (8 * (c ^ 1)));   // + 512 dd
#pragma unroll
      for (int q = 0; q < 16; ++q) XB[x2w + 16 * q] = v[q].x;
      __syncthreads();
#pragma unroll
      for (int dd = 0; dd < 8; ++dd) v[dd].x = XB[x2r_lo + 512 * dd];
#pragma unroll
      for (int dd = 8; dd < 16; ++dd) v[dd].x = XB[x2r_hi + 512 * dd];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 16; ++q) XB[x2w + 16 * q] = v[q].y;
      __syncthreads();
#pragma unroll
      for (int dd = 0; dd < 8; ++dd) v[dd].y = XB[x2r_lo + 512 * dd];
#pragma unroll
      for (int dd = 8; dd < 16; ++dd) v[dd].y = XB[x2r_hi + 512 * dd];
    }
    __syncthreads();                        // E: the exchange buffer is warp-private again
    // The block barriers of X2 leave all 16 warps at the same point of the same instruction sequence, and left alone
    // they stay in step: FP64 phases and shared-memory phases of all warps coincide and the two pipes take turns.
    // Holding back two of the four warps of every scheduler by about one phase makes them complementary.
    if (skew_ns) {
      const int g = (opaque(tid0) >> 7) & 3;                             // the four warps of a scheduler: 0 .. 3
      const unsigned ns = (skew_ns >> 16) ? (skew_ns & 0xffff) * g : ((g & 1) ? skew_ns : 0);
      if (ns) __nanosleep(ns);
    }
    dft<16>(v);                             // v[e] = E (c = 0) / O (c = 1) [k2 + 256 e]
    {
      // ---- phase 3: member mm = 2 mu + c of group G of warp w holds E/O of residue k2 (mu: the mirrored residue)
      const int tid = opaque(tid0), w = tid >> 5, lane = tid & 31;
      const int mu = lane >> 4, c = (lane >> 3) & 1, G = lane & 7, mm = 2 * mu + c;
      const bool special = !INV && (w == 0) && (G == 0);                 // residues 0 and 128 pair with themselves
      const int x3b = 512 * w + 64 * G;
      const int ph0 = (2 * G) & 15, ph1 = (2 * G + 1) & 15;             // XOR of the members with even / odd index
      const int phw = c ? ph1 : ph0;
      const int x3w = x3b + 16 * mm;                                     // element e at + (e ^ phw)
      // quadruple i: e_i = eb + i es; sources: members (mlo, mlo + 1) at e_i and (mhi, mhi + 1) at et - e_i
      int eb = mm, es = 4, et = 15, mlo = 0, mhi = 2, kres = 8 * w + G;
      if (special) {
        es = 2;
        if (mu) { eb = c; mlo = 2; kres = 128; }
        else { eb = 1 + c; et = 16; mhi = 0; kres = 0; }
      }
      const bool first = special && mm == 1;                             // this thread also emits k = 0 and k = M/2
      double2 e0 = make_double2(0.0, 0.0), o0 = e0;
      // X3: after it v[4 i + s] = E[k], O[k], E[M/2 - k], O[M/2 - k] of quadruple i (real parts, then imaginary)
#pragma unroll
      for (int e = 0; e < 16; ++e) XB[x3w + (e ^ phw)] = v[e].x;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = eb + i * es, em = et - e;
        v[4 * i].x = XB[x3b + 16 * mlo + (e ^ ph0)];
        v[4 * i + 1].x = XB[x3b + 16 * (mlo + 1) + (e ^ ph1)];
        v[4 * i + 2].x = XB[x3b + 16 * mhi + (em ^ ph0)];
        v[4 * i + 3].x = XB[x3b + 16 * (mhi + 1) + (em ^ ph1)];
      }
      if (first) { e0.x = XB[x3b + ph0]; o0.x = XB[x3b + 16 + ph1]; }
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 16; ++e) XB[x3w + (e ^ phw)] = v[e].y;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = eb + i * es, em = et - e;
        v[4 * i].y = XB[x3b + 16 * mlo + (e ^ ph0)];
        v[4 * i + 1].y = XB[x3b + 16 * (mlo + 1) + (e ^ ph1)];
        v[4 * i + 2].y = XB[x3b + 16 * mhi + (em ^ ph0)];
        v[4 * i + 3].y = XB[x3b + 16 * (mhi + 1) + (em ^ ph1)];
      }
      if (first) { e0.y = XB[x3b + ph0]; o0.y = XB[x3b + 16 + ph1]; }
      __syncwarp();
      double* out = dst + (long long)row * ld_dst;
      if constexpr (INV) {
        // ---- radix 2 + the two sectors y[4n .. 4n+3], y[4n' .. 4n'+3] (n' = M/2 - 1 - n) of every quadruple
        const double2 wm1 = T4[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = kres + 256 * (eb + i * es), n2 = WH - 1 - n;
          // exp(-2 pi i n / M) = Q[n]^8; exp(-2 pi i (M/2 - 1 - n) / M) = -conj(exp(-2 pi i (n + 1) / M))
          const double2 q = q_at<true>(tq, n, WM);
          const double2 qq = cmul(q, q);
          const double2 q4 = cmul(qq, qq);
          const double2 wn = cmul(q4, q4);
          const double2 w1n = cmul(wn, wm1);
          const double2 wb = cmul(wn, v[4 * i + 1]);
          const double2 wd = cmul(make_double2(-w1n.x, w1n.y), v[4 * i + 3]);
          const double2 z_n = cadd(v[4 * i], wb), z_nh = csub(v[4 * i], wb);          // z[n], z[n + M/2]
          const double2 z_c = cadd(v[4 * i + 2], wd), z_m = csub(v[4 * i + 2], wd);   // z[M/2-1-n], z[M-1-n]
          // stored swapped: (.y, .x) = (re, im)
          *reinterpret_cast<double2*>(out + 4 * n) = make_double2(z_n.y, z_m.x);
          *reinterpret_cast<double2*>(out + 4 * n + 2) = make_double2(z_n.x, z_m.y);
          *reinterpret_cast<double2*>(out + 4 * n2) = make_double2(z_c.y, z_nh.x);
          *reinterpret_cast<double2*>(out + 4 * n2 + 2) = make_double2(z_c.x, z_nh.y);
        }
        continue;
      }
      // ---- radix 2 + untangling + quarter-wave rotation, 8 outputs per quadruple
      if (first) {
        const double2 z0 = cadd(e0, o0), zh = csub(e0, o0);
        out[0] = (z0.x + z0.y) * scale0;
        out[WM] = (z0.x - z0.y) * RH * scale;
        dct2_emit(out, WH, WM, WN, zh, zh, make_double2(0.92387953251128675613, -0.38268343236508977173),
                  make_double2(0.0, -1.0), scale);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = kres + 256 * (eb + i * es);
        const double2 q = q_at<true>(tq, k, WM);
        const double2 qq = cmul(q, q);
        const double2 tn = cmul(qq, qq);                                 // exp(-2 pi i k / N)
        const double2 wk = cmul(tn, tn);                                 // exp(-2 pi i k / M)
        const double2 wb = cmul(wk, v[4 * i + 1]);
        const double2 cwd = cmul(make_double2(wk.x, -wk.y), v[4 * i + 3]);
        const double2 z_k = cadd(v[4 * i], wb), z_kh = csub(v[4 * i], wb);
        const double2 z_hk = csub(v[4 * i + 2], cwd), z_mk = cadd(v[4 * i + 2], cwd);
        const double2 q2 = cmul(make_double2(0.92387953251128675613, -0.38268343236508977173), make_double2(q.x, -q.y));
        dct2_emit(out, k, WM, WN, z_k, z_mk, q, tn, scale);
        dct2_emit(out, WH - k, WM, WN, z_hk, z_kh, q2, make_double2(-tn.y, -tn.x), scale);
      }
    }
  }
}

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

}  // namespace

typedef CUresult (*DctEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DctEncodeFn dct_encoder() {
  static DctEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (DctEncodeFn)f;
    cudaGetLastError();
  }
  return fn;
}

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
  static int rr_off = -1;
  if (rr_off < 0) rr_off = getenv("AXB_DCT_SMEM") ? 1 : 0;
  static int w_off = -1;
  if (w_off < 0) w_off = getenv("AXB_DCT_RR") ? 1 : 0;                 // A/B switch: the register-resident kernel
  if (vec && !rr_off && !w_off && N == WN) {
    const size_t wb_bytes = (size_t)(WN + WM) * sizeof(double) + (256 + (WM >> 3) + 1 + 33 + 17) * sizeof(double2) + 16;
    static bool w_once = false;
    if (!w_once) {
      cudaFuncSetAttribute(k_dct_rows_w<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(k_dct_rows_w<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      w_once = true;
    }
    const int grid = rows < sms ? rows : sms;
    static int skew = -1;
    if (skew < 0) skew = getenv("AXB_DCT_SKEW") ? atoi(getenv("AXB_DCT_SKEW")) : 600;
    CUtensorMap tmW;
    memset(&tmW, 0, sizeof(tmW));
    bool ok = false;
    if (DctEncodeFn enc = dct_encoder()) {
      // DCT-II : (16 doubles, 8 x 1 KB, 8 x 128 B, 16 x 8 KB, rows), the 1 KB dimension first (stage_addr8)
      // DCT-III: (16 doubles, 8 x 256 B, 2 x 128 B, 64 x 2 KB, rows) (stage3_addr8); 4 boxes of 32 KB either way
      const cuuint64_t dims2[5] = {16, 8, 8, 16, (cuuint64_t)rows}, dims3[5] = {16, 8, 2, 64, (cuuint64_t)rows};
      const cuuint64_t str2[4] = {1024, 128, 8192, (cuuint64_t)ld_src * 8}, str3[4] = {256, 128, 2048, (cuuint64_t)ld_src * 8};
      const cuuint32_t box2[5] = {16u, 8u, 8u, 4u, 1u}, box3[5] = {16u, 8u, 2u, 16u, 1u};
      const cuuint64_t* dims = inverse ? dims3 : dims2;
      const cuuint64_t* strides = inverse ? str3 : str2;
      const cuuint32_t* box = inverse ? box3 : box2;
      const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
      ok = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double*>(src), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (ok) {
      if (inverse)
        k_dct_rows_w<true><<<grid, 512, wb_bytes, st>>>(tmW, rows, src, ld_src, dst, ld_dst,
                                                         reinterpret_cast<const double2*>(tabs), scale0, scale,
                                                         (unsigned)skew);
      else
        k_dct_rows_w<false><<<grid, 512, wb_bytes, st>>>(tmW, rows, src, ld_src, dst, ld_dst,
                                                        reinterpret_cast<const double2*>(tabs), scale0, scale,
                                                        (unsigned)skew);
      AXB_LAUNCHED();
      return (int)cudaGetLastError();
    }
  }
  if (vec && !rr_off && ((ilog2(M) - 1) % 4 == 0)) {
    // register-resident kernel with the next row prefetched by TMA
    int rr_rpc = 1;
    while (rr_rpc * T < 128) rr_rpc *= 2;
    int tw_entries = 0;
    for (int p = 1; p < (ilog2(M) - 1) / 4; ++p) tw_entries += 2 << (4 * p);
    const size_t rr_bytes = ((size_t)rr_rpc * N + (size_t)rr_rpc * (M + (M >> 4) + 16)) * sizeof(double) + 16 +
                            ((size_t)(M >> 3) + 1 + tw_entries) * sizeof(double2);
    static bool rr_once = false;
    if (!rr_once) {
      cudaFuncSetAttribute(k_dct_rows_rr<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(k_dct_rows_rr<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      rr_once = true;
    }
    const int rr_blocks = (rows + rr_rpc - 1) / rr_rpc;
    int rr_res = (int)((227u * 1024u) / (rr_bytes + 1024));
    const int by_regs = 512 / (rr_rpc * T);                 // 128 registers per thread
    if (rr_res > by_regs) rr_res = by_regs;
    if (rr_res < 1) rr_res = 1;
    const int rr_grid = rr_blocks < sms * rr_res ? rr_blocks : sms * rr_res;
    const double2* tbp = reinterpret_cast<const double2*>(tabs);
    // DCT-II: the row is staged as two arrays of pairs by tensor loads with a 2-element inner box over the
    // (quadruple, 4) view of the row, so the first pass reads conflict-free 128-bit pairs
    CUtensorMap tmS;
    memset(&tmS, 0, sizeof(tmS));
    int split = 0;
    static int lin = -1;
    if (lin < 0) lin = getenv("AXB_DCT_LINEAR") ? 1 : 0;
    if (!inverse && !lin) {
      if (DctEncodeFn enc = dct_encoder()) {
        const cuuint64_t quads = N / 4, mid = quads < 256 ? quads : 256;
        const cuuint64_t dims[4] = {4, mid, quads / mid, (cuuint64_t)rows};
        const cuuint64_t strides[3] = {32, mid * 32, (cuuint64_t)ld_src * 8};
        const cuuint32_t box[4] = {2u, (cuuint32_t)mid, (cuuint32_t)(quads / mid), 1u};
        const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
        split = enc(&tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(src), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      }
    }
    if (!inverse)
      k_dct_rows_rr<false><<<rr_grid, rr_rpc * T, rr_bytes, st>>>(tmS, split, rows, N, ilog2(M), rr_rpc, src, ld_src, dst,
                                                                   ld_dst, tbp, scale0, scale);
    else
      k_dct_rows_rr<true><<<rr_grid, rr_rpc * T, rr_bytes, st>>>(tmS, 0, rows, N, ilog2(M), rr_rpc, src, ld_src, dst,
                                                                  ld_dst, tbp, scale0, scale);
    AXB_LAUNCHED();
    return (int)cudaGetLastError();
  }
  const int threads = rpc * T;
  // persistent CTAs: exactly as many as are resident at once (the kernels take 128 registers, so the register file and
  // not shared memory bounds the residency at N >= 2048; a grid sized by shared memory alone ran a second, half-empty
  // wave); AXB_DCT_GRID_SMEM=1 restores the shared-memory estimate for an A/B
  int occ = 0;
  if (!inverse) {
    if (vec) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dct2_rows<true>, threads, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dct2_rows<false>, threads, smem);
  } else {
    if (vec) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dct3_rows<true>, threads, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dct3_rows<false>, threads, smem);
  }
  static int by_smem = -1;
  if (by_smem < 0) by_smem = getenv("AXB_DCT_GRID_SMEM") ? 1 : 0;
  const int per_sm = (int)((227u * 1024u) / (smem + 1024));
  int resident = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  if (!by_smem && occ >= 1 && occ < resident) resident = occ;
  const int nblocks = (rows + rpc - 1) / rpc;
  const int grid = nblocks < sms * resident ? nblocks : sms * resident;
  const double2* tb = reinterpret_cast<const double2*>(tabs);
  const int logM = ilog2(M);
  static int pf = -1;
  if (pf < 0) pf = getenv("AXB_DCT_NO_PREFETCH") ? 0 : 1;            // A/B switch: no L2 prefetch of the next row
  if (!inverse) {
    // (measured: the prefetch pays for the DCT-III, 0.103 -> 0.096 ms at 8192 x 2048, whose load phase also reads the
    // rotation table; the DCT-II is unchanged to 1 % slower with it)
    if (vec) k_dct2_rows<true><<<grid, threads, smem, st>>>(rows, N, logM, rpc, src, ld_src, dst, ld_dst, tb, scale0, scale, 0);
    else k_dct2_rows<false><<<grid, threads, smem, st>>>(rows, N, logM, rpc, src, ld_src, dst, ld_dst, tb, scale0, scale, 0);
  } else {
    if (vec) k_dct3_rows<true><<<grid, threads, smem, st>>>(rows, N, logM, rpc, src, ld_src, dst, ld_dst, tb, pf);
    else k_dct3_rows<false><<<grid, threads, smem, st>>>(rows, N, logM, rpc, src, ld_src, dst, ld_dst, tb, 0);
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
