// axb_common.cuh -- shared device/host helpers for libaxisym_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "axisym_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libaxisym_b200 is written for sm_100a (B200) only"
#endif

extern int64_t g_axb_launches;  // defined in capi.cu

#define AXB_LAUNCHED() (++g_axb_launches)
#define AXB_RETURN_LAST()                 \
  do {                                    \
    cudaError_t e__ = cudaGetLastError(); \
    return (int)e__;                      \
  } while (0)

static inline int axb_check_grid(const axb_grid_t* g) {
  if (!g) return AXB_EINVAL;
  if (g->nr < 1 || g->nz < 1 || g->ld < g->nz) return AXB_EINVAL;
  if (g->ku0 < 0 || g->ku1 > g->nz || g->ku0 > g->ku1) return AXB_EINVAL;
  if (g->nz_global < 1) return AXB_EINVAL;
  if (g->ju1 != 0 && (g->ju0 < 0 || g->ju1 > g->nr || g->ju0 > g->ju1)) return AXB_EINVAL;
  if (g->batch > 1) return AXB_ENOSUP;       // batched entries validate with axb_check_grid_batched
  return AXB_OK;
}
// for the entries that serve an ensemble with one launch (member index = blockIdx.z)
static inline int axb_check_grid_batched(const axb_grid_t* g) {
  if (!g) return AXB_EINVAL;
  if (g->batch <= 1) return axb_check_grid(g);
  axb_grid_t one = *g;
  one.batch = 0;
  const int rc = axb_check_grid(&one);
  if (rc) return rc;
  if (g->batch > 65535 || g->batch_stride < g->nz || g->scalar_stride < 0) return AXB_EINVAL;
  return AXB_OK;
}
static inline bool axb_al8(const void* p) { return (((uintptr_t)p) & 7u) == 0; }
static inline bool axb_al16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// Device-side copy of the grid descriptor (passed by value to kernels).
struct GridD {
  int nr, nz;
  long long ld;
  double dx;
  int kz0, nzg, ku0, ku1;
  int ju0, ju1;      // owned rows: the ones fused reductions count
  int batch, sstride;   // ensemble members per launch (>= 1), doubles between their device scalars
  long long bstride;    // elements between the members' fields
};
static inline GridD to_dev(const axb_grid_t* g) {
  GridD d;
  d.nr = g->nr; d.nz = g->nz; d.ld = g->ld; d.dx = g->dx;
  d.kz0 = g->kz0; d.nzg = g->nz_global; d.ku0 = g->ku0; d.ku1 = g->ku1;
  d.ju0 = g->ju1 ? g->ju0 : 0; d.ju1 = g->ju1 ? g->ju1 : g->nr;
  d.batch = g->batch > 1 ? g->batch : 1;
  d.sstride = g->batch > 1 ? g->scalar_stride : 0;
  d.bstride = g->batch > 1 ? g->batch_stride : 0;
  return d;
}
#ifdef __CUDACC__
// Batched launches: blockIdx.z is the ensemble member.  A kernel moves its field pointers by member_field(g)
// elements and its device-scalar pointers by member_scalar(g) doubles (both 0 for a single field); null stays null.
__device__ __forceinline__ long long member_field(const GridD& g) { return (long long)blockIdx.z * g.bstride; }
__device__ __forceinline__ int member_scalar(const GridD& g) { return (int)blockIdx.z * g.sstride; }
template <class T>
__device__ __forceinline__ T* moved(T* p, long long o) { return p ? p + o : p; }
#endif

// 2 columns per thread, (32 x 8) threads per block: a block covers 64 columns x 8 rows.
constexpr int TBX = 32, TBY = 8;
static inline dim3 grid2d(const GridD& g) {
  const int kpairs = (g.nz + 1) / 2;
  return dim3((kpairs + TBX - 1) / TBX, (g.nr + TBY - 1) / TBY, g.batch);
}

// (f[k], f[k+1]) of one row; 128-bit load when the row is 16-byte aligned, k even and
// k+1 inside the row, else two guarded 64-bit loads (second lane clamps to k).
__device__ __forceinline__ double2 ld_pair(const double* __restrict__ row, int k, int nz, bool vec) {
  if (vec && (k + 1 < nz)) return *reinterpret_cast<const double2*>(row + k);
  double2 r;
  r.x = row[k];
  r.y = row[(k + 1 < nz) ? k + 1 : k];
  return r;
}
__device__ __forceinline__ void st_pair(double* __restrict__ row, int k, int k_lo, int k_hi, bool vec,
                                        double2 v) {
  if (vec && k >= k_lo && k + 1 < k_hi) {
    *reinterpret_cast<double2*>(row + k) = v;
  } else {
    if (k >= k_lo && k < k_hi) row[k] = v.x;
    if (k + 1 >= k_lo && k + 1 < k_hi) row[k + 1] = v.y;
  }
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide reductions for blocks of up to 1024 threads; result valid in thread 0
__device__ __forceinline__ double block_max(double v) {
  __shared__ double sm[32];
  const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
  const int wid = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;
  const int nw = (blockDim.x * blockDim.y + 31) >> 5;
  v = warp_max(v);
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < nw) ? sm[lane] : -INFINITY;
    v = warp_max(v);
  }
  return v;
}
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sm[32];
  const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
  const int wid = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;
  const int nw = (blockDim.x * blockDim.y + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < nw) ? sm[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}
// max of NON-NEGATIVE doubles: their bit patterns order like unsigned integers
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
// max of arbitrary doubles via CAS
__device__ __forceinline__ void atomic_max_any(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) >= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}
