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
// Three generations of the sweep kernel live here (all walk 32- to 128-column groups down the rows in lockstep):
//   k_tri_sweep<DIR, TC>      cp.async ring, TC columns per CTA (AXB_TRI_COLS; operands that are not TMA friendly);
//   k_tri_sweep_tma<DIR>      one warp per 32 columns: TMA loads, chain and TMA stores in that warp (round-2 start,
//                             axb_set_tridiag_sweep(1));
//   k_tri_sweep_ws<DIR, NS>   the default: a producer lane feeds a ring of NS boxes (8 rows x 32 columns of the field
//                             and of the reciprocal pivots + the row coefficients) by TMA, a second warp runs the
//                             dependent chain out of registers and stores the rows.
#include <cuda.h>   // CUtensorMap (types only; the encoder comes through cudaGetDriverEntryPoint)
#include <stdlib.h>

#include "axb_common.cuh"

extern int g_axb_tri_one_warp;   // capi.cu: axb_set_tridiag_sweep

namespace {

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

// row_coef[m] = { c1 sub[m-1] (0 at m = 0), scale[m] (1 without scale), c1 sup[m] (0 at m = nr-1), 0 }
__global__ void k_tri_row_coef(int nr, const double* __restrict__ sub, const double* __restrict__ sup,
                               const double* __restrict__ scale, double c1, double* __restrict__ rc) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nr) return;
  rc[4 * m + 0] = (m > 0) ? c1 * sub[m - 1] : 0.0;
  rc[4 * m + 1] = scale ? scale[m] : 1.0;
  rc[4 * m + 2] = (m < nr - 1) ? c1 * sup[m] : 0.0;
  rc[4 * m + 3] = 0.0;
}

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
template <int DIR, int TC>
__device__ __forceinline__ void issue_stage(double* sx, double* sp, int st, int m0, int nr, const double* __restrict__ X,
                                            long long ld, const double* __restrict__ inv, int nz, int k0, int lane) {
  constexpr int LPR = TC / 2;                                  // lanes per row (16 bytes each)
  const int half = lane / LPR, c = (lane % LPR) * 2;
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
template <int TC>
__device__ __forceinline__ void cta_sync() {
  if constexpr (TC <= 32) __syncwarp((TC == 32) ? 0xffffffffu : ((1u << TC) - 1u));
  else __syncthreads();
}

template <int DIR, int TC>
__global__ void __launch_bounds__(TC)
    k_tri_sweep(int nr, int nz, double* __restrict__ X, long long ld, const double* __restrict__ inv,
                const double* __restrict__ rc) {
  extern __shared__ __align__(16) double tri_smem[];
  double* sx = tri_smem;
  double* sp = tri_smem + TS * TR * TC;
  const int lane = threadIdx.x;
  const int k0 = blockIdx.x * TC;
  const int nb = (nr + TR - 1) / TR;
  const int start = (DIR > 0) ? 0 : nr - 1;
#pragma unroll
  for (int i = 0; i < TS - 1; ++i) {
    if (i < nb) issue_stage<DIR, TC>(sx, sp, i, start + DIR * i * TR, nr, X, ld, inv, nz, k0, lane);
    cp_commit();
  }
  double carry = 0.0;        // y_{m-1} (forward) / x_{m+1} (backward)
  double pprev = 0.0;        // 1 / den_{m-1} (forward only)
  for (int i = 0; i < nb; ++i) {
    cta_sync<TC>();                                             // stage (i - 1) % TS is free again
    const int nxt = i + TS - 1;
    if (nxt < nb) issue_stage<DIR, TC>(sx, sp, nxt % TS, start + DIR * nxt * TR, nr, X, ld, inv, nz, k0, lane);
    cp_commit();
    cp_wait<TS - 1>();
    cta_sync<TC>();
    const int st = i % TS;
    // all loads of the stage first (no control flow in between), then the dependent chain: with only
    // ~3.5 warps per SM the latency has to be hidden inside the thread
    double r[TR], p[TR], co[TR], sc[TR];
#pragma unroll
    for (int u = 0; u < TR; ++u) {
      const int m = start + DIR * (i * TR + u);
      const int mc = m < 0 ? 0 : (m >= nr ? nr - 1 : m);
      r[u] = sx[(st * TR + u) * TC + lane];
      p[u] = sp[(st * TR + u) * TC + lane];
      if (DIR > 0) {
        co[u] = rc[4 * mc];
        sc[u] = rc[4 * mc + 1];
      } else {
        co[u] = rc[4 * mc + 2];
      }
    }
#pragma unroll
    for (int u = 0; u < TR; ++u) {
      const int m = start + DIR * (i * TR + u);
      if (DIR > 0) {
        carry = r[u] * sc[u] - (co[u] * pprev) * carry;
        pprev = p[u];
      } else {
        carry = (r[u] - co[u] * carry) * p[u];
      }
      if (m >= 0 && m < nr) X[(long long)m * ld + k0 + lane] = carry;
    }
  }
}

// ---- TMA variant -----------------------------------------------------------------------------
// One warp per CTA owns 32 columns.  Lane 0 keeps TS - 1 boxes (8 rows x 32 columns of the field and of
// the pivots) in flight with cp.async.bulk.tensor + mbarrier complete_tx; the warp consumes a box per
// iteration and the finished rows leave through a double-buffered TMA store, so the instruction
// stream of the only warp is the dependent chain and little else.  Out-of-range rows / columns are
// zero-filled on load and clipped on store by the tensor map.
constexpr int WC = 32;   // columns per warp
constexpr int WS = 8;    // stages
constexpr int W_STAGE = TR * WC;                       // doubles per box
constexpr unsigned STAGE_TX = (2 * W_STAGE + 4 * TR) * 8;   // bytes landing on a stage's mbarrier

__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mb_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "TRI_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra TRI_DONE;\n"
      "bra TRI_WAIT;\n"
      "TRI_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_ld(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_st(const CUtensorMap* map, int c0, int c1, unsigned src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(
                   reinterpret_cast<unsigned long long>(map)),
               "r"(c0), "r"(c1), "r"(src)
               : "memory");
}

template <int DIR>
__global__ void __launch_bounds__(32)
    k_tri_sweep_tma(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmP,
                    const __grid_constant__ CUtensorMap tmC, int nr) {
  __shared__ __align__(128) double sx[WS * W_STAGE];
  __shared__ __align__(128) double sp[WS * W_STAGE];
  __shared__ __align__(128) double sc4[WS * TR * 4];         // row coefficients of the boxes in flight
  __shared__ __align__(128) double so[2 * W_STAGE];
  __shared__ __align__(8) unsigned long long bars[WS];
  const int lane = threadIdx.x;
  const int k0 = blockIdx.x * WC;
  const int nb = (nr + TR - 1) / TR;
  // first row of box i: the same row-0-aligned boxes walked down (forward) or up (backward); only the
  // box at the far end can stick out (beyond nr: zero-filled loads, clipped stores -- TMA stores do
  // not take negative coordinates)
  auto row0 = [&](int i) { return (DIR > 0) ? i * TR : (nb - 1 - i) * TR; };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < WS; ++i) mb_init(s_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < WS - 1; ++i) {
      if (i < nb) {
        mb_expect(s_u32(&bars[i]), STAGE_TX);
        tma_ld(s_u32(sx + i * W_STAGE), &tmX, k0, row0(i), s_u32(&bars[i]));
        tma_ld(s_u32(sp + i * W_STAGE), &tmP, k0, row0(i), s_u32(&bars[i]));
        tma_ld(s_u32(sc4 + i * TR * 4), &tmC, 0, row0(i), s_u32(&bars[i]));
      }
    }
  }
  __syncwarp();
  double carry = 0.0, pprev = 0.0;
  for (int i = 0; i < nb; ++i) {
    __syncwarp();                                             // box (i - 1) % WS has been consumed
    const int nxt = i + WS - 1;
    if (lane == 0 && nxt < nb) {
      const int st = nxt % WS;
      mb_expect(s_u32(&bars[st]), STAGE_TX);
      tma_ld(s_u32(sx + st * W_STAGE), &tmX, k0, row0(nxt), s_u32(&bars[st]));
      tma_ld(s_u32(sp + st * W_STAGE), &tmP, k0, row0(nxt), s_u32(&bars[st]));
      tma_ld(s_u32(sc4 + st * TR * 4), &tmC, 0, row0(nxt), s_u32(&bars[st]));
    }
    const int st = i % WS;
    mb_wait(s_u32(&bars[st]), (i / WS) & 1);
    const int m0 = row0(i);
    double r[TR], p[TR], co[TR], sc[TR];
#pragma unroll
    for (int u = 0; u < TR; ++u) {
      r[u] = sx[st * W_STAGE + u * WC + lane];
      p[u] = sp[st * W_STAGE + u * WC + lane];
      if (DIR > 0) {
        const double2 c = *reinterpret_cast<const double2*>(sc4 + (st * TR + u) * 4);   // broadcast read
        co[u] = c.x;
        sc[u] = c.y;
      } else {
        co[u] = sc4[(st * TR + u) * 4 + 2];
      }
    }
    double* ob = so + (i & 1) * W_STAGE;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");   // store i - 2 has left ob
    __syncwarp();
    if (DIR > 0) {
#pragma unroll
      for (int u = 0; u < TR; ++u) {
        carry = r[u] * sc[u] - (co[u] * pprev) * carry;
        pprev = p[u];
        ob[u * WC + lane] = carry;
      }
    } else {
#pragma unroll
      for (int u = TR - 1; u >= 0; --u) {
        carry = (r[u] - co[u] * carry) * p[u];
        ob[u * WC + lane] = carry;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      tma_st(&tmX, k0, m0, s_u32(ob));
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

// ---- warp-specialised TMA variant -------------------------------------------------------------
// In k_tri_sweep_tma the only warp also issues the TMA loads, fences and stores of every box: ~100 cycles per
// row of which the dependent chain is ~10 (an already-complete mbarrier try_wait costs ~90 cycles, the proxy fence
// before a TMA store ~110, each tensor load issue an ELECT loop).  With <= 8192 columns there are fewer warps than SM
// sub-partitions, so that instruction stream IS the duration (2048 rows x 52 ns = 0.107 ms whatever the width).
// Here warp 0's elected lane is the producer (empty[] barriers -> expect_tx + 3 tensor loads per box) and warp 1
// consumes: the try_wait on box i + 1 is issued before the chain of box i and confirmed after it, box i + 1 moves
// from shared memory to a second register set while the rows of box i leave with plain 256-byte row stores (no
// staging buffer, no proxy fence, no bulk-group wait).
__device__ __forceinline__ unsigned mb_try(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mb_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}

template <int DIR>
struct SweepBox {
  double r[TR], p[TR], co[TR], sc[TR];
};

// NS = boxes in the ring (8 / 16 / 32: 4.3 KB each).  With <= 8192 columns there are <= 256 CTAs and the bytes in
// flight (CTAs x NS x 4.3 KB) bound the sweep before the chain does, so narrow problems get a deeper ring.
// (Measured at 1024 x 4096, ncu: the consumer warp issues ~170 instructions per box at one instruction per ~3 cycles --
// fixed-latency dependency waits of a single in-order warp -- so neither a deeper ring, nor a shorter chain (one FMA per
// row), nor three producer warps issuing one tensor load each changed the ~500 cycles per box.)
template <int DIR, int NS>
__global__ void __launch_bounds__(64)
    k_tri_sweep_ws(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmP,
                   const __grid_constant__ CUtensorMap tmC, int nr, int nz, double* __restrict__ X, long long ld) {
  extern __shared__ __align__(128) double ws_smem[];
  double* sx = ws_smem;
  double* sp = sx + NS * W_STAGE;
  double* sc4 = sp + NS * W_STAGE;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(sc4 + NS * TR * 4);
  unsigned long long* empty = full + NS;
  const int lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * WC;
  const int nb = (nr + TR - 1) / TR;
  auto row0 = [&](int i) { return (DIR > 0) ? i * TR : (nb - 1 - i) * TR; };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      mb_init(s_u32(&full[i]), 1);
      mb_init(s_u32(&empty[i]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {                                       // producer warp: one lane
    if (lane == 0) {
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS;
        if (i >= NS) mb_wait(s_u32(&empty[st]), ((i / NS) - 1) & 1);   // the consumer has taken use (i / NS) - 1
        mb_expect(s_u32(&full[st]), STAGE_TX);
        tma_ld(s_u32(sx + st * W_STAGE), &tmX, k0, row0(i), s_u32(&full[st]));
        tma_ld(s_u32(sp + st * W_STAGE), &tmP, k0, row0(i), s_u32(&full[st]));
        tma_ld(s_u32(sc4 + st * TR * 4), &tmC, 0, row0(i), s_u32(&full[st]));
      }
    }
    return;
  }
  const bool col_ok = k0 + lane < nz;
  auto fetch = [&](SweepBox<DIR>& b, int i) {                   // box i: shared memory -> registers
    const int st = i % NS;
#pragma unroll
    for (int u = 0; u < TR; ++u) {
      b.r[u] = sx[st * W_STAGE + u * WC + lane];
      b.p[u] = sp[st * W_STAGE + u * WC + lane];
      if (DIR > 0) {
        const double2 c = *reinterpret_cast<const double2*>(sc4 + (st * TR + u) * 4);   // broadcast read
        b.co[u] = c.x;
        b.sc[u] = c.y;
      } else {
        b.co[u] = sc4[(st * TR + u) * 4 + 2];
      }
    }
  };
  double carry = 0.0, pprev = 0.0;
  double out[TR];
  auto chain = [&](const SweepBox<DIR>& b) {
    if (DIR > 0) {
#pragma unroll
      for (int u = 0; u < TR; ++u) {
        carry = b.r[u] * b.sc[u] - (b.co[u] * pprev) * carry;
        pprev = b.p[u];
        out[u] = carry;
      }
    } else {
#pragma unroll
      for (int u = TR - 1; u >= 0; --u) {
        carry = (b.r[u] - b.co[u] * carry) * b.p[u];
        out[u] = carry;
      }
    }
  };
  const bool cols_full = k0 + WC <= nz;                         // CTA-uniform
  auto store = [&](int i) {
    const int m0 = row0(i);
    double* xo = X + (long long)m0 * ld + k0 + lane;
    if (cols_full && m0 + TR <= nr) {                           // whole box inside: eight unpredicated row stores
#pragma unroll
      for (int u = 0; u < TR; ++u) xo[(long long)u * ld] = out[u];
    } else {
#pragma unroll
      for (int u = 0; u < TR; ++u)
        if (col_ok && m0 + u < nr) xo[(long long)u * ld] = out[u];
    }
  };
  // one iteration: (try box i + 1) -> chain of box i -> (confirm, fetch box i + 1 into `nxt`) -> store box i, release
  auto step = [&](const SweepBox<DIR>& cur, SweepBox<DIR>& nxt, int i) {
    const bool more = i + 1 < nb;
    const unsigned nbar = s_u32(&full[(i + 1) % NS]);
    const unsigned npar = ((i + 1) / NS) & 1;
    unsigned ok = 1;
    if (more) ok = mb_try(nbar, npar);
    chain(cur);
    if (more) {
      if (!ok) mb_wait(nbar, npar);
      fetch(nxt, i + 1);
    }
    store(i);
    __syncwarp();                                               // every lane has box i in registers
    if (lane == 0) mb_arrive(s_u32(&empty[i % NS]));
  };
  SweepBox<DIR> A, B;
  mb_wait(s_u32(&full[0]), 0);
  fetch(A, 0);
  int i = 0;
  for (; i + 1 < nb; i += 2) {
    step(A, B, i);
    step(B, A, i + 1);
  }
  if (i < nb) step(A, B, i);
}

template <int NS>
static void launch_ws(int grid, const CUtensorMap& tmX, const CUtensorMap& tmP, const CUtensorMap& tmC, int nr, int nz,
                      double* X, long long ld, cudaStream_t s) {
  constexpr size_t smem = (size_t)(2 * NS * W_STAGE + NS * TR * 4 + 2 * NS) * sizeof(double);
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(k_tri_sweep_ws<1, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_tri_sweep_ws<-1, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    once = true;
  }
  k_tri_sweep_ws<1, NS><<<grid, 64, smem, s>>>(tmX, tmP, tmC, nr, nz, X, ld);
  AXB_LAUNCHED();
  k_tri_sweep_ws<-1, NS><<<grid, 64, smem, s>>>(tmX, tmP, tmC, nr, nz, X, ld);
  AXB_LAUNCHED();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tri_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
    cudaGetLastError();
  }
  return fn;
}
bool tri_map(CUtensorMap* m, const double* ptr, int nr, int nz, long long ld, int box_cols = WC) {
  EncodeTiledFn enc = tri_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)nz, (cuuint64_t)nr};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)TR};
  const cuuint32_t estr[2] = {1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Partition (SPIKE) correction for an r range split over ranks: after the local solve g of this rank's
// rows, x = g - V xl - W xr with the two neighbouring interface unknowns
//   xl[k] = sum_i CL[i][k] G[i][k],  xr[k] = sum_i CR[i][k] G[i][k]      (i < n_iface = 2 * ranks),
// G = the first / last local-solve rows of every rank, CL / CR = rows of the inverted reduced system.
__global__ void __launch_bounds__(128)
    k_tri_partition_correct(int rows, int nz, double* __restrict__ X, long long ld, const double* __restrict__ V,
                            const double* __restrict__ W, const double* __restrict__ G, const double* __restrict__ CL,
                            const double* __restrict__ CR, int n_iface, int rows_per_block,
                            const int* __restrict__ vcut_blk, const int* __restrict__ wcut_blk) {
  // The spikes decay away from the interfaces (mode k of the Poisson-like operator like rho_k^m, rho_k < 1): for all
  // but the lowest modes V is below 1e-20 of its maximum after a few dozen rows, W before the last few dozen.  vcut_blk /
  // wcut_blk (optional) hold, per 256-column group, the first row from which V is negligible and the first row that W
  // reaches; a block whose rows lie in between has nothing to correct and leaves (X is corrected in place).
  if (vcut_blk) {
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    if (r0 >= vcut_blk[blockIdx.x] && r1 <= wcut_blk[blockIdx.x]) return;
  }
  const int k = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (k >= nz) return;
  const bool two = k + 1 < nz;
  double xl0 = 0, xl1 = 0, xr0 = 0, xr1 = 0;
  for (int i = 0; i < n_iface; ++i) {
    const long long o = (long long)i * nz + k;
    const double g0 = G[o], g1 = two ? G[o + 1] : 0.0;
    xl0 += CL[o] * g0;
    xr0 += CR[o] * g0;
    if (two) { xl1 += CL[o + 1] * g1; xr1 += CR[o + 1] * g1; }
  }
  const int m0 = blockIdx.y * rows_per_block, m1 = min(rows, m0 + rows_per_block);
  for (int m = m0; m < m1; ++m) {
    const long long o = (long long)m * ld + k, t = (long long)m * nz + k;
    X[o] = (X[o] - V[t] * xl0) - W[t] * xr0;
    if (two) X[o + 1] = (X[o + 1] - V[t + 1] * xl1) - W[t + 1] * xr1;
  }
}

bool tri_fast_ok(int nz, const double* X, long long ld, const double* inv) {
  return inv && (nz % 16 == 0) && axb_al16(X) && axb_al16(inv) && (ld % 2 == 0);
}

template <int TC>
static void launch_sweeps(int nr, int nz, double* X, long long ld, const double* inv, const double* rc,
                          cudaStream_t s) {
  constexpr size_t smem = 2 * TS * TR * TC * sizeof(double);
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(k_tri_sweep<1, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_tri_sweep<-1, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    once = true;
  }
  k_tri_sweep<1, TC><<<nz / TC, TC, smem, s>>>(nr, nz, X, ld, inv, rc);
  AXB_LAUNCHED();
  k_tri_sweep<-1, TC><<<nz / TC, TC, smem, s>>>(nr, nz, X, ld, inv, rc);
  AXB_LAUNCHED();
}

int launch_tri_factored(int nr, int nz, double* X, long long ld, const double* inv, const double* rc,
                        cudaStream_t s) {
  if (nr < 2 || nz < 1 || !X || !inv || !rc || ld < nz) return AXB_EINVAL;
  if (!tri_fast_ok(nz, X, ld, inv)) return AXB_EINVAL;
  // widest column block that divides nz and still leaves >= ~1 CTA per SM
  static int force = -1;
  if (force < 0) {
    const char* e = getenv("AXB_TRI_COLS");     // 16 / 32 / 64 / 128: cp.async variants; unset: TMA
    force = e ? atoi(e) : 0;
  }
  if (g_axb_tri_one_warp < 0) {
    const char* e = getenv("AXB_TRI_ONE_WARP");   // 1: the single-warp TMA kernel (loads, chain and TMA stores in one warp)
    g_axb_tri_one_warp = (e && atoi(e) == 1) ? 1 : 0;
  }
  const int one_warp = g_axb_tri_one_warp;
  if (force == 0) {
    CUtensorMap tmX, tmP, tmC;
    if (tri_map(&tmX, X, nr, nz, ld) && tri_map(&tmP, inv, nr, nz, nz) && tri_map(&tmC, rc, nr, 4, 4, 4)) {
      const int grid = (nz + WC - 1) / WC;
      if (one_warp == 1) {
        k_tri_sweep_tma<1><<<grid, 32, 0, s>>>(tmX, tmP, tmC, nr);
        AXB_LAUNCHED();
        k_tri_sweep_tma<-1><<<grid, 32, 0, s>>>(tmX, tmP, tmC, nr);
        AXB_LAUNCHED();
      } else {
        // ring depth by CTAs per SM (bytes in flight = CTAs x depth x 4.3 KB; 5 / 3 / 1 CTAs of 35 / 70 / 139 KB fit on an
        // SM); AXB_TRI_RING=8/16/32 forces a depth
        static int ring = -1, sms = 0;
        if (ring < 0) {
          const char* e = getenv("AXB_TRI_RING");
          ring = e ? atoi(e) : 0;
          int dev = 0;
          cudaGetDevice(&dev);
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        int ns = (grid <= sms) ? 32 : (grid <= 2 * sms) ? 16 : 8;
        if (ring == 8 || ring == 16 || ring == 32) ns = ring;
        if (ns == 32) launch_ws<32>(grid, tmX, tmP, tmC, nr, nz, X, ld, s);
        else if (ns == 16) launch_ws<16>(grid, tmX, tmP, tmC, nr, nz, X, ld, s);
        else launch_ws<8>(grid, tmX, tmP, tmC, nr, nz, X, ld, s);
      }
      return (int)cudaGetLastError();
    }
  }
  int tc = 16;
  if (nz % 128 == 0 && nz >= 128 * 96) tc = 128;
  else if (nz % 64 == 0 && nz >= 64 * 96) tc = 64;
  else if (nz % 32 == 0) tc = 32;
  if ((force == 16 || force == 32 || force == 64 || force == 128) && nz % force == 0) tc = force;
  switch (tc) {
    case 128: launch_sweeps<128>(nr, nz, X, ld, inv, rc, s); break;
    case 64: launch_sweeps<64>(nr, nz, X, ld, inv, rc, s); break;
    case 32: launch_sweeps<32>(nr, nz, X, ld, inv, rc, s); break;
    default: launch_sweeps<16>(nr, nz, X, ld, inv, rc, s); break;
  }
  return (int)cudaGetLastError();
}

extern "C" {

int axb_tridiag_factor_columns(int nr, int nz, const double* sub, const double* diag, const double* sup,
                               const double* lam, const double* scale, double c0, double c1, double* inv_pivots,
                               double* row_coef, axb_stream_t s) {
  if (nr < 2 || nz < 1 || !sub || !diag || !sup || !lam || !inv_pivots || !row_coef) return AXB_EINVAL;
  k_tri_factor<<<(nz + 127) / 128, 128, 0, (cudaStream_t)s>>>(nr, nz, sub, diag, sup, lam, c0, c1, inv_pivots);
  AXB_LAUNCHED();
  k_tri_row_coef<<<(nr + 127) / 128, 128, 0, (cudaStream_t)s>>>(nr, sub, sup, scale, c1, row_coef);
  AXB_LAUNCHED();
  return (int)cudaGetLastError();
}

int axb_tridiag_partition_correct_banded(int rows, int nz, double* X, int64_t ld, const double* V, const double* W,
                                         const double* G, const double* CL, const double* CR, int n_iface,
                                         const int32_t* vcut_blk, const int32_t* wcut_blk, axb_stream_t s) {
  if (rows < 1 || nz < 1 || !X || !V || !W || !G || !CL || !CR || n_iface < 2 || ld < nz) return AXB_EINVAL;
  if ((vcut_blk == nullptr) != (wcut_blk == nullptr)) return AXB_EINVAL;
  const int rpb = 32;
  k_tri_partition_correct<<<dim3((nz / 2 + 128) / 128, (rows + rpb - 1) / rpb), 128, 0, (cudaStream_t)s>>>(
      rows, nz, X, ld, V, W, G, CL, CR, n_iface, rpb, vcut_blk, wcut_blk);
  AXB_LAUNCHED();
  return (int)cudaGetLastError();
}
int axb_tridiag_partition_correct(int rows, int nz, double* X, int64_t ld, const double* V, const double* W,
                                  const double* G, const double* CL, const double* CR, int n_iface, axb_stream_t s) {
  return axb_tridiag_partition_correct_banded(rows, nz, X, ld, V, W, G, CL, CR, n_iface, nullptr, nullptr, s);
}

int axb_tridiag_solve_factored(int nr, int nz, double* X, int64_t ld, const double* inv_pivots, const double* row_coef,
                               axb_stream_t s) {
  return launch_tri_factored(nr, nz, X, ld, inv_pivots, row_coef, (cudaStream_t)s);
}

}  // extern "C"
