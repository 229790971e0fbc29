// particles.cu -- the particle <-> mesh interpolation family of the reference's C++ core
// (pyaxisymflow/core/src/instantiate.yml:1-27), every kernel flavour and every boundary rule:
//
//   mesh -> particles (gather)   core/src/interpolation/mesh_to_particles_2D.hpp:12-175 (periodic),
//                                :186-425 (edge clipped, "unbounded"), mesh_to_particles_1D.hpp:63-108
//   particles -> mesh (scatter)  core/src/interpolation/particles_to_mesh_2D.hpp:13-148, :156-324,
//                                particles_to_mesh_1D.hpp:66-120
//   wrap                         core/src/interpolation/wrap_particles.hpp:35-120
//   weights                      particle_kernels/LinearKernel.hpp:15-18, MP4.hpp:21-39, MP6.hpp:27-50,
//                                YangSmoothThreePointKernel.hpp:31-52 through KernelWrapper.hpp:35-51
//
// One thread per particle.  The gather accumulates in the reference's order (x inside y, partial sums
// per stencil row), so with -fmad=false its results equal the serial C++ bit for bit (Yang's kernel goes
// through asin/sqrt: last-ulp differences between libm and CUDA's are possible there).  The scatter adds
// with FP64 atomics: indices and weights are exact, the sum order is not (<= 1e-14 relative).
#include "axb_common.cuh"

namespace {

// ---- weight functions: `it` is the KernelWrapper iteration index (0 .. size-1), t the scaled distance ----
template <int KID>
struct PK;
template <>
struct PK<AXB_PK_LINEAR> {
  static constexpr int S = -1, E = 1;
  __device__ static __forceinline__ double w(int, double t) { return 1.0 - t; }
};
template <>
struct PK<AXB_PK_MP4> {
  static constexpr int S = -2, E = 2;
  __device__ static __forceinline__ double w(int it, double t) {
    return (it == 1 || it == 2) ? 1.0 + t * t * (-2.5 + 1.5 * t) : 2.0 + t * (-4.0 + t * (2.5 - 0.5 * t));
  }
};
template <>
struct PK<AXB_PK_MP6> {
  static constexpr int S = -3, E = 3;
  __device__ static __forceinline__ double w(int it, double t) {
    if (it == 2 || it == 3) return 0.08333333333333333 * (1.0 - t) * (12.0 + t * (12.0 + t * (-3.0 + t * (-38.0 + 25.0 * t))));
    if (it == 1 || it == 4) return 0.041666666666666666 * (t - 1.0) * (t - 2.0) * (-48.0 + t * (153.0 + t * (-114.0 + 25.0 * t)));
    const double c = t - 3.0;
    return 0.041666666666666666 * (2.0 - t) * (5.0 * t - 8.0) * c * c * c;
  }
};
template <>
struct PK<AXB_PK_YANG> {
  static constexpr int S = -2, E = 2;
  __device__ static __forceinline__ double w(int it, double t) {
    if (it == 1 || it == 2)
      return 0.4045499823398394 + 0.25 * t * (1.0 - t) + (0.0625 - 0.125 * t) * sqrt(1.0 + 12.0 * t * (1.0 - t)) -
             0.14433756729740643 * asin(0.8660254037844386 * (2.0 * t - 1.0));
    return 1.0954500176601605 + t * (0.25 * t - 1.0833333333333333) +
           (0.04166666666666666 * t - 0.0625) * sqrt(12.0 * t * (3.0 - t) - 23.0) +
           0.04811252243246881 * asin(0.8660254037844386 * (2.0 * t - 3.0));
  }
};

// nearest-upper mesh index of a coordinate and the stencil weights (mesh_to_particles_2D.hpp:66-95)
template <int KID>
__device__ __forceinline__ int axis_weights(double pos, double delta, double* w) {
  using K = PK<KID>;
  const double p = pos / delta;
  const double fl = floor(p);
  const int hi = (int)(p >= (fl + 0.5)) + (int)fl;
#pragma unroll
  for (int i = K::S; i < K::E; ++i) w[i - K::S] = K::w(i - K::S, fabs(p - (hi + 0.5 + i)));
  return hi;
}

// stencil range of one axis: full, or clipped to the array when any side of the 2-D stencil leaves it
// (mesh_to_particles_2D.hpp:300-333; the reference clips BOTH axes as soon as one of them is near an edge,
// which gives the same ranges as clipping each axis on its own)
template <int KID>
__device__ __forceinline__ void clip(int hi, int n, int& s, int& e) {
  using K = PK<KID>;
  s = max(K::S, -hi);
  e = min(K::E, n - hi);
}

template <int KID>
__global__ void k_m2p_2d(int m0, int m1, const double* __restrict__ fx, const double* __restrict__ fy, long long np,
                         const double* __restrict__ px, const double* __restrict__ py, double* __restrict__ ox,
                         double* __restrict__ oy, double dx, double dy, bool periodic) {
  using K = PK<KID>;
  constexpr int N = K::E - K::S;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= np) return;
  double wx[N], wy[N];
  const int hx = axis_weights<KID>(px[id], dx, wx);
  const int hy = axis_weights<KID>(py[id], dy, wy);
  int s0 = K::S, e0 = K::E, s1 = K::S, e1 = K::E;
  if (!periodic) {
    clip<KID>(hx, m1, s0, e0);
    clip<KID>(hy, m0, s1, e1);
  }
  double ax = 0.0, ay = 0.0;
#pragma unroll
  for (int sy = K::S; sy < K::E; ++sy) {
    double psx = 0.0, psy = 0.0;                                   // partial sums of this stencil row
    if (sy >= s1 && sy < e1) {
      int row = hy + sy;
      if (periodic) row = (row + m0) % m0;
      const long long base = (long long)row * m1;
#pragma unroll
      for (int sx = K::S; sx < K::E; ++sx) {
        if (sx < s0 || sx >= e0) continue;
        int col = hx + sx;
        if (periodic) col = (col + m1) % m1;
        const double w = wx[sx - K::S];
        psx += w * fx[base + col];
        psy += w * fy[base + col];
      }
    }
    const double w = wy[sy - K::S];
    ax += w * psx;
    ay += w * psy;
  }
  ox[id] = ax;
  oy[id] = ay;
}

template <int KID>
__global__ void k_p2m_2d(int m0, int m1, long long np, const double* __restrict__ px, const double* __restrict__ py,
                         const double* __restrict__ val, double* mesh, double dx, double dy, bool periodic) {
  using K = PK<KID>;
  constexpr int N = K::E - K::S;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= np) return;
  double wx[N], wy[N];
  const int hx = axis_weights<KID>(px[id], dx, wx);
  const int hy = axis_weights<KID>(py[id], dy, wy);
  const double v = val[id];
  int s0 = K::S, e0 = K::E, s1 = K::S, e1 = K::E;
  if (!periodic) {
    clip<KID>(hx, m1, s0, e0);
    clip<KID>(hy, m0, s1, e1);
  }
  for (int sy = s1; sy < e1; ++sy) {
    int row = hy + sy;
    if (periodic) row = (row + m0) % m0;
    for (int sx = s0; sx < e0; ++sx) {
      int col = hx + sx;
      if (periodic) col = (col + m1) % m1;
      atomicAdd(&mesh[(long long)row * m1 + col], (wy[sy - K::S] * wx[sx - K::S]) * v);   // particles_to_mesh_2D.hpp:316-318
    }
  }
}

// 1-D MP4 (the *_with_offset_impl flavours, periodic wrap)
__global__ void k_m2p_1d(int m, const double* __restrict__ f, int np, const double* __restrict__ p, double* __restrict__ o,
                         double dx) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= np) return;
  double w[4];
  const int hi = axis_weights<AXB_PK_MP4>(p[id], dx, w);
  double acc = 0.0;
#pragma unroll
  for (int i = -2; i < 2; ++i) acc += w[i + 2] * f[(hi + i + m) % m];
  o[id] = acc;
}
__global__ void k_p2m_1d(int m, int np, const double* __restrict__ p, const double* __restrict__ v, double* mesh, double dx) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= np) return;
  double w[4];
  const int hi = axis_weights<AXB_PK_MP4>(p[id], dx, w);
  const double in = v[id];
#pragma unroll
  for (int i = -2; i < 2; ++i) atomicAdd(&mesh[(hi + i + m) % m], w[i + 2] * in);
}

// wrap_particles.hpp: "FirstN" = only the first / last 10 entries of a line are tested, "All" = every entry
__device__ __forceinline__ double wrap_lo(double p, double a, double len) { return (p < a) ? p + len : p; }
__device__ __forceinline__ double wrap_hi(double p, double b, double len) { return (p > b) ? p - len : p; }
__global__ void k_wrap_x(int n0, int n1, double* px, double a, double b) {
  // rows of px: first min(10, n1) entries wrap up, entries from max(left, n1 - 10) on wrap down
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n0) return;
  const double len = b - a;
  const int left = min(10, n1), right = max(left, n1 - 10);
  double* row = px + (long long)j * n1;
  for (int i = 0; i < left; ++i) row[i] = wrap_lo(row[i], a, len);
  for (int i = right; i < n1; ++i) row[i] = wrap_hi(row[i], b, len);
}
__global__ void k_wrap_y(int n0, int n1, double* py, double a, double b) {
  // the first / last 10 ROWS, every entry of them, both tests (a row in both sets gets both in the
  // reference's order: bottom set first)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n1) return;
  const double len = b - a;
  const int bottom = min(10, n0), top = max(bottom, n0 - 10);
  for (int j = 0; j < bottom; ++j) {
    double* q = py + (long long)j * n1 + i;
    *q = wrap_hi(wrap_lo(*q, a, len), b, len);
  }
  for (int j = top; j < n0; ++j) {
    double* q = py + (long long)j * n1 + i;
    *q = wrap_hi(wrap_lo(*q, a, len), b, len);
  }
}

}  // namespace

extern "C" {

int axb_m2p_2d(int kernel, int m0, int m1, const double* fx, const double* fy, int p0, int p1, const double* px,
               const double* py, double* ox, double* oy, double dx, double dy, int periodic, axb_stream_t s_) {
  if (!fx || !fy || !px || !py || !ox || !oy || m0 < 1 || m1 < 1 || p0 < 1 || p1 < 1) return AXB_EINVAL;
  cudaStream_t s = (cudaStream_t)s_;
  const long long np = (long long)p0 * p1;
  const unsigned blocks = (unsigned)((np + 127) / 128);
  const bool per = periodic != 0;
  switch (kernel) {
    case AXB_PK_LINEAR: k_m2p_2d<AXB_PK_LINEAR><<<blocks, 128, 0, s>>>(m0, m1, fx, fy, np, px, py, ox, oy, dx, dy, per); break;
    case AXB_PK_MP4: k_m2p_2d<AXB_PK_MP4><<<blocks, 128, 0, s>>>(m0, m1, fx, fy, np, px, py, ox, oy, dx, dy, per); break;
    case AXB_PK_MP6: k_m2p_2d<AXB_PK_MP6><<<blocks, 128, 0, s>>>(m0, m1, fx, fy, np, px, py, ox, oy, dx, dy, per); break;
    case AXB_PK_YANG: k_m2p_2d<AXB_PK_YANG><<<blocks, 128, 0, s>>>(m0, m1, fx, fy, np, px, py, ox, oy, dx, dy, per); break;
    default: return AXB_EINVAL;
  }
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_p2m_2d(int kernel, int m0, int m1, int p0, int p1, const double* px, const double* py, const double* val,
               double* mesh, double dx, double dy, int periodic, axb_stream_t s_) {
  if (!px || !py || !val || !mesh || m0 < 1 || m1 < 1 || p0 < 1 || p1 < 1) return AXB_EINVAL;
  cudaStream_t s = (cudaStream_t)s_;
  const long long np = (long long)p0 * p1;
  const unsigned blocks = (unsigned)((np + 127) / 128);
  const bool per = periodic != 0;
  cudaMemsetAsync(mesh, 0, sizeof(double) * (size_t)m0 * m1, s);
  switch (kernel) {
    case AXB_PK_LINEAR: k_p2m_2d<AXB_PK_LINEAR><<<blocks, 128, 0, s>>>(m0, m1, np, px, py, val, mesh, dx, dy, per); break;
    case AXB_PK_MP4: k_p2m_2d<AXB_PK_MP4><<<blocks, 128, 0, s>>>(m0, m1, np, px, py, val, mesh, dx, dy, per); break;
    case AXB_PK_MP6: k_p2m_2d<AXB_PK_MP6><<<blocks, 128, 0, s>>>(m0, m1, np, px, py, val, mesh, dx, dy, per); break;
    case AXB_PK_YANG: k_p2m_2d<AXB_PK_YANG><<<blocks, 128, 0, s>>>(m0, m1, np, px, py, val, mesh, dx, dy, per); break;
    default: return AXB_EINVAL;
  }
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_m2p_1d_mp4(int m, const double* field, int np, const double* pos, double* out, double dx, axb_stream_t s) {
  if (!field || !pos || !out || m < 1 || np < 1) return AXB_EINVAL;
  k_m2p_1d<<<(np + 127) / 128, 128, 0, (cudaStream_t)s>>>(m, field, np, pos, out, dx);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_p2m_1d_mp4(int m, int np, const double* pos, const double* val, double* mesh, double dx, axb_stream_t s) {
  if (!pos || !val || !mesh || m < 1 || np < 1) return AXB_EINVAL;
  cudaMemsetAsync(mesh, 0, sizeof(double) * (size_t)m, (cudaStream_t)s);
  k_p2m_1d<<<(np + 127) / 128, 128, 0, (cudaStream_t)s>>>(m, np, pos, val, mesh, dx);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_wrap_particles_2d(int n0, int n1, double* px, double* py, double x0, double x1, double y0, double y1,
                          axb_stream_t s) {
  if (n0 < 1 || n1 < 1 || (!px && !py)) return AXB_EINVAL;
  if (px) {
    k_wrap_x<<<(n0 + 127) / 128, 128, 0, (cudaStream_t)s>>>(n0, n1, px, x0, x1);
    AXB_LAUNCHED();
  }
  if (py) {
    k_wrap_y<<<(n1 + 127) / 128, 128, 0, (cudaStream_t)s>>>(n0, n1, py, y0, y1);
    AXB_LAUNCHED();
  }
  AXB_RETURN_LAST();
}

}  // extern "C"
