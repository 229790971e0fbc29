// native_core.cu -- CUDA equivalents of the reference's two C++/pybind11 core routines
// that sit on the configured per-timestep paths:
//
//   G-LS   least-squares wavefront extrapolation, order 1, 3x3 patch
//          (core/src/extrapolate_using_least_squares.hpp:136-446,
//           lstsq/ExtrapolationConfig.hpp:131-158, lstsq/least_squares.hpp:12-59)
//   G-P2M  MP4 particles-to-mesh remeshing, edge-clipped and periodic
//          (core/src/interpolation/particles_to_mesh_2D.hpp:13-148, :156-324,
//           particle_kernels/MP4.hpp:21-39, KernelWrapper.hpp:35-51)
//
// G-LS is data dependent (a do/while over wavefront sweeps).  Each sweep is three launches
// over a compacted list of pending cells: classify (integer flag arithmetic, bit-exact patch
// offsets) -> solve (one thread per candidate: 3x3 gather masked by the OLD flags, normal
// equations, Gauss elimination with partial pivoting, evaluation) -> flag.  Reads inside a
// sweep are masked by the pre-sweep flags and flags only flip afterwards, so the sweep is
// order independent and the result equals the serial reference bit for bit.  The file is
// compiled with -fmad=false and reproduces the reference's operation order, so the floating
// point results are bit-identical to the C++ reference as well (tests assert equality).
#include <initializer_list>

#include "axb_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// G-LS
// ---------------------------------------------------------------------------------------
struct LsWork {
  int* pend_a;
  int* pend_b;
  int* cand;
  short* codes;    // (start_x + 3) | (start_y + 3) << 8
  int* counters;   // [0] pending-in, [1] pending-out, [2] candidates
  long long cap;
};

__device__ __forceinline__ short ls_start(short pos_sum, short cnt) {
  // ExtrapolationConfig.hpp:131-158: temp is float, the quotient is evaluated in double and
  // truncated to short; start = (2*code - 6) / 2
  float temp = (pos_sum > 4) ? (float)(ceil(0.5 * (double)(float)pos_sum) * 2.0) : (float)pos_sum;
  short code = (short)(2.0 * (double)temp / (double)cnt);
  return (short)((short)(2 * code - 6) / (short)2);
}

__global__ void k_ls_collect(int n, const short* __restrict__ cur, const short* __restrict__ tgt, int* pend,
                             int* counters, long long cap) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  if (cur[c] ^ tgt[c]) {
    const int slot = atomicAdd(&counters[0], 1);
    if (slot < cap) pend[slot] = c;
  }
}

// pending -> (candidate with codes | still pending).  Counts live on the device (n_in pending cells, clamped to the
// list capacity); grid-stride, so the same kernel serves the host-driven loop (grid sized from the host's copy of the
// count) and the device-terminated one (fixed grid, the count may be anything including zero).
__global__ void k_ls_classify(int n1, const short* __restrict__ cur, const int* __restrict__ pend_in, int* pend_out,
                              int* cand, short* codes, const int* __restrict__ n_in, int* n_out, int* n_cand,
                              long long cap) {
  const int n = (int)min((long long)*n_in, cap);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int c = pend_in[t];
    short cnt = 0, sx = 0, sy = 0;
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
      for (int dk = -1; dk <= 1; ++dk) {
        const short f = cur[c + dj * n1 + dk];
        if (dj != 0 || dk != 0) cnt += f;
        sx += f * (short)(dk + 1);
        sy += f * (short)(dj + 1);
      }
    if (cnt) {
      const int slot = atomicAdd(n_cand, 1);
      cand[slot] = c;
      codes[slot] = (short)((ls_start(sx, cnt) + 3) | ((ls_start(sy, cnt) + 3) << 8));
    } else {
      const int slot = atomicAdd(n_out, 1);
      pend_out[slot] = c;
    }
  }
}

template <int NC>
__device__ __forceinline__ void gauss_nc(double (&A)[NC][NC + 2], double (&sol)[2][NC]) {
  // lstsq/least_squares.hpp:12-59 with NCoefficients = NC (3: order 1, 6: order 2), NComponents = 2
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    int piv = i;
    double best = fabs(A[i][i]);
#pragma unroll
    for (int k = i + 1; k < NC; ++k)
      if (fabs(A[k][i]) > best) { best = fabs(A[k][i]); piv = k; }
#pragma unroll
    for (int k = i; k < NC + 2; ++k) {
      // select-based swap keeps the arrays in registers
      const double a = A[i][k];
      double b = a;
#pragma unroll
      for (int r = 0; r < NC; ++r) if (r == piv) b = A[r][k];
#pragma unroll
      for (int r = 0; r < NC; ++r) if (r == piv) A[r][k] = a;
      A[i][k] = b;
    }
#pragma unroll
    for (int k = i + 1; k < NC; ++k) {
      const double c = -A[k][i] / A[i][i];
      A[k][i] = 0.0;
#pragma unroll
      for (int j = i + 1; j < NC + 2; ++j) A[k][j] += c * A[i][j];
    }
  }
#pragma unroll
  for (int comp = 0; comp < 2; ++comp)
#pragma unroll
    for (int i = NC - 1; i >= 0; --i) {
      sol[comp][i] = A[i][NC + comp] / A[i][i];
#pragma unroll
      for (int k = i - 1; k >= 0; --k) A[k][NC + comp] -= A[k][i] * sol[comp][i];
    }
}
// basis rows of binomial_fill_data_parallel (extrapolate_using_least_squares.hpp:43-61): each order multiplies
// the previous order's rows, [m, m x, m y, (m x) x, (m x) y, (m y) y]
template <int NC>
__device__ __forceinline__ void ls_basis(double m, double x, double y, double (&b)[NC]) {
  b[0] = m;
  b[1] = m * x;
  b[2] = m * y;
  if constexpr (NC == 6) {
    b[3] = b[1] * x;
    b[4] = b[1] * y;
    b[5] = b[2] * y;
  }
}

template <int NC>
__global__ void k_ls_solve(int n1, const short* __restrict__ cur, const int* __restrict__ cand,
                           const short* __restrict__ codes, const int* __restrict__ n_cand, double* eta_x,
                           double* eta_y, const double* __restrict__ gx, const double* __restrict__ gy) {
  const int n = *n_cand;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
  const int c = cand[t];
  const int j = c / n1, k = c - j * n1;
  const int code = codes[t];
  const int k0 = k + ((code & 0xff) - 3), j0 = j + (((code >> 8) & 0xff) - 3);
  double L[NC][9], rhs[2][9];
#pragma unroll
  for (int p = 0; p < 9; ++p) {
    const int jj = j0 + p / 3, kk = k0 + p % 3;
    const int gi = jj * n1 + kk;
    const double m = (double)cur[gi];
    double b[NC];
    ls_basis<NC>(m, gx[kk], gy[jj], b);
#pragma unroll
    for (int a = 0; a < NC; ++a) L[a][p] = b[a];
    // volatile-free 8-byte loads: a concurrent writer can only be an unflagged cell (m = 0)
    rhs[0][p] = m * eta_x[gi];
    rhs[1][p] = m * eta_y[gi];
  }
  double M[NC][NC + 2];
#pragma unroll
  for (int a = 0; a < NC; ++a) {
#pragma unroll
    for (int b = 0; b < NC; ++b) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i) s += L[a][i] * L[b][i];
      M[a][b] = s;
    }
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i) s += L[a][i] * rhs[d][i];
      M[a][NC + d] = s;
    }
  }
  double sol[2][NC];
  gauss_nc<NC>(M, sol);
  double basis[NC];
  ls_basis<NC>(1.0, gx[k], gy[j], basis);
  double e = 0.0;
#pragma unroll
  for (int i = 0; i < NC; ++i) e += sol[0][i] * basis[i];
  eta_x[c] = e;
  e = 0.0;
#pragma unroll
  for (int i = 0; i < NC; ++i) e += sol[1][i] * basis[i];
  eta_y[c] = e;
  }
}

// sweeps_done (may be null) <- sweep + 1 when this sweep added cells
__global__ void k_ls_flag(short* cur, const int* __restrict__ cand, const int* __restrict__ n_cand, int* sweeps_done,
                          int sweep) {
  const int n = *n_cand;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) cur[cand[t]] = 1;
  if (sweeps_done && n > 0 && blockIdx.x == 0 && threadIdx.x == 0) *sweeps_done = sweep + 1;
}
// device-terminated form, after the last sweep: status bit 0 = the pending list overflowed its capacity, bit 1 = the
// sweep budget ran out while cells were still being added; out = {status, sweeps that added cells}
__global__ void k_ls_finish(const int* __restrict__ ctr, int sweeps, long long cap, int* out) {
  int st = 0;
  if (ctr[4] > cap) st |= 1;
  if (ctr[4 + 2 * sweeps] > 0 && ctr[4 + 2 * (sweeps - 1) + 1] > 0) st |= 2;
  out[0] = st;
  out[1] = ctr[1];
}
// after a sweep: pending-in <- pending-out, zero the other counters (single thread)
__global__ void k_ls_roll(int* counters) {
  counters[0] = counters[1];
  counters[1] = 0;
  counters[2] = 0;
}

LsWork carve(void* work, long long cap) {
  LsWork w;
  char* p = (char*)work;
  w.counters = (int*)p; p += 64;
  w.pend_a = (int*)p; p += cap * 4;
  w.pend_b = (int*)p; p += cap * 4;
  w.cand = (int*)p; p += cap * 4;
  w.codes = (short*)p;
  w.cap = cap;
  return w;
}
inline long long ls_cap_from_bytes(long long bytes) { return (bytes - 64) / 14; }

int ls_run(int order, int n0, int n1, short* cur, const short* tgt, double* eta_x, double* eta_y, const double* gx,
           const double* gy, void* work, long long work_bytes, int max_sweeps, int* sweeps_host, cudaStream_t s) {
  const long long cap = ls_cap_from_bytes(work_bytes);
  if (cap < 1) return AXB_EWORK;
  LsWork w = carve(work, cap);
  const int n = n0 * n1;
  cudaMemsetAsync(w.counters, 0, 64, s);
  k_ls_collect<<<(n + 255) / 256, 256, 0, s>>>(n, cur, tgt, w.pend_a, w.counters, cap);
  AXB_LAUNCHED();
  int h[3] = {0, 0, 0};
  cudaMemcpyAsync(h, w.counters, sizeof(h), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return (int)e;
  if (h[0] > cap) return AXB_EWORK;
  int pending = h[0], sweeps = 0;
  int *pin = w.pend_a, *pout = w.pend_b;
  while (pending > 0 && (max_sweeps <= 0 || sweeps < max_sweeps)) {
    const int blocks = (pending + 127) / 128;
    k_ls_classify<<<blocks, 128, 0, s>>>(n1, cur, pin, pout, w.cand, w.codes, w.counters, w.counters + 1,
                                         w.counters + 2, cap);
    if (order == 2) k_ls_solve<6><<<blocks, 128, 0, s>>>(n1, cur, w.cand, w.codes, w.counters + 2, eta_x, eta_y, gx, gy);
    else k_ls_solve<3><<<blocks, 128, 0, s>>>(n1, cur, w.cand, w.codes, w.counters + 2, eta_x, eta_y, gx, gy);
    k_ls_flag<<<blocks, 128, 0, s>>>(cur, w.cand, w.counters + 2, nullptr, 0);
    g_axb_launches += 3;
    cudaMemcpyAsync(h, w.counters, sizeof(h), cudaMemcpyDeviceToHost, s);
    k_ls_roll<<<1, 1, 0, s>>>(w.counters);
    AXB_LAUNCHED();
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return (int)e;
    if (h[2] == 0) break;  // nothing could be added: the reference returns here
    ++sweeps;
    pending = h[1];
    int* t = pin; pin = pout; pout = t;
  }
  if (sweeps_host) *sweeps_host = sweeps;
  return (int)cudaGetLastError();
}

// The same sweeps without a host round trip (graph capturable): every sweep has its own counter pair
//   ctr[4 + 2 s] = cells pending before sweep s,  ctr[4 + 2 s + 1] = candidates found in sweep s   (ctr[1] = sweeps done)
// so nothing has to be rolled between sweeps; the kernels run on a fixed grid and read their counts from the device.
// A sweep that finds no candidate leaves the pending list as it is, and so do all later ones: the result is the one
// of the host-driven loop, which stops there (extrapolate_using_least_squares.hpp:372-375).
constexpr int LS_MAX_SWEEPS = 28;            // 4 + 2 (sweeps + 1) ints fit the 256-byte counter block
struct LsDevWork {
  int* ctr;
  int* pend[2];
  int* cand;
  short* codes;
  long long cap;
};
LsDevWork carve_device(void* work, long long work_bytes) {
  LsDevWork w;
  w.cap = (work_bytes - 256) / 14;
  char* p = (char*)work;
  w.ctr = (int*)p; p += 256;
  w.pend[0] = (int*)p; p += w.cap * 4;
  w.pend[1] = (int*)p; p += w.cap * 4;
  w.cand = (int*)p; p += w.cap * 4;
  w.codes = (short*)p;
  return w;
}
// `collected`: the caller has zeroed the counters and filled the first pending list (k_ls_fill<true>)
int ls_run_device(int order, int n0, int n1, short* cur, const short* tgt, double* eta_x, double* eta_y, const double* gx,
                  const double* gy, void* work, long long work_bytes, int sweeps, int* status_dev, bool collected,
                  cudaStream_t s) {
  if (sweeps < 1 || sweeps > LS_MAX_SWEEPS) return AXB_EINVAL;
  const LsDevWork w = carve_device(work, work_bytes);
  const long long cap = w.cap;
  if (cap < 1) return AXB_EWORK;
  int* ctr = w.ctr;
  int* const* pend = w.pend;
  int* cand = w.cand;
  short* codes = w.codes;
  const int n = n0 * n1;
  if (!collected) {
    cudaMemsetAsync(ctr, 0, 256, s);
    k_ls_collect<<<(n + 255) / 256, 256, 0, s>>>(n, cur, tgt, pend[0], ctr + 4, cap);
    AXB_LAUNCHED();
  }
  const int blocks = (int)min((cap + 127) / 128, 148LL * 4);
  for (int i = 0; i < sweeps; ++i) {
    int *n_in = ctr + 4 + 2 * i, *n_cand = n_in + 1, *n_out = n_in + 2;
    k_ls_classify<<<blocks, 128, 0, s>>>(n1, cur, pend[i & 1], pend[(i + 1) & 1], cand, codes, n_in, n_out, n_cand, cap);
    if (order == 2) k_ls_solve<6><<<blocks, 128, 0, s>>>(n1, cur, cand, codes, n_cand, eta_x, eta_y, gx, gy);
    else k_ls_solve<3><<<blocks, 128, 0, s>>>(n1, cur, cand, codes, n_cand, eta_x, eta_y, gx, gy);
    k_ls_flag<<<blocks, 128, 0, s>>>(cur, cand, n_cand, ctr + 1, i);
    g_axb_launches += 3;
  }
  if (status_dev) {
    k_ls_finish<<<1, 1, 0, s>>>(ctr, sweeps, cap, status_dev);
    AXB_LAUNCHED();
  }
  return (int)cudaGetLastError();
}

// fill of the doubled work arrays, elasto_kernels/extrapolate_eta_using_least_squares_unb.py:20-25
// + the flag construction of elasto_kernels/extrapolate_using_least_squares.py:33-36
// COLLECT: the pending list (cells with cur != tgt, k_ls_collect) is built here as well -- its order does not matter
// to the sweeps -- and the target flags, which only the collection reads, are not stored.
template <bool COLLECT>
__global__ void k_ls_fill(int nr, int nz, long long ld, const double* __restrict__ phi,
                          const unsigned char* __restrict__ inside, const double* __restrict__ eta1,
                          const double* __restrict__ eta2, double zone, short* cur, short* tgt, double* e1,
                          double* e2, int* pend, int* n_pend, long long cap) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (k >= nz) return;
  const long long src = (long long)j * ld + k;
  const double p = -phi[src];
  const double m = (double)inside[(long long)j * nz + k];  // the mask is a dense (nr, nz) bool array
  const double a = m * eta1[src], b = m * eta2[src];
  const short c = (short)(p < 0), t = (short)(p < zone);
  const long long up = (long long)(nr + j) * nz + k, dn = (long long)(nr - 1 - j) * nz + k;
  cur[up] = c; cur[dn] = c;
  e1[up] = a; e1[dn] = a;
  e2[up] = b; e2[dn] = -b;
  if (COLLECT) {
    if (c ^ t) {
      const int slot = atomicAdd(n_pend, 2);
      if (slot + 1 < cap) { pend[slot] = (int)up; pend[slot + 1] = (int)dn; }
    }
  } else {
    tgt[up] = t; tgt[dn] = t;
  }
}
__global__ void k_ls_unfill(int nr, int nz, long long ld, double* eta1, double* eta2, const double* __restrict__ e1,
                            const double* __restrict__ e2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (k >= nz) return;
  const long long up = (long long)(nr + j) * nz + k;
  eta1[(long long)j * ld + k] = e1[up];
  eta2[(long long)j * ld + k] = e2[up];
}

// ---------------------------------------------------------------------------------------
// G-P2M
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double mp4_w0(double t) { return 1.0 + t * t * (-2.5 + 1.5 * t); }
__device__ __forceinline__ double mp4_w1(double t) { return 2.0 + t * (-4.0 + t * (2.5 - 0.5 * t)); }

// weights + nearest-upper mesh index of one coordinate (particles_to_mesh_2D.hpp:214-260)
__device__ __forceinline__ int mp4_axis(double pos, double delta, double w[4]) {
  const double p = pos / delta;
  const double fl = floor(p);
  const int hi = (int)(p >= (fl + 0.5)) + (int)fl;
#pragma unroll
  for (int i = -2; i < 2; ++i) {
    const double t = fabs(p - (hi + 0.5 + i));
    w[i + 2] = (i == -2 || i == 1) ? mp4_w1(t) : mp4_w0(t);
  }
  return hi;
}

// scatter of one particle into a mesh of n0 x n1 whose row 0 is global row `row_off`
// of the reference's (N0 x n1) mesh; rows outside [0, n0) are dropped (mirror half).
__device__ __forceinline__ void mp4_scatter(double* mesh, int n0, int n1, long long ld, int N0, int row_off,
                                            int hx, int hy, const double wx[4], const double wy[4], double v,
                                            bool periodic) {
  int s0 = -2, e0 = 2, s1 = -2, e1 = 2;
  if (!periodic) {
    if (hx - 2 < 0 || hy - 2 < 0 || hx + 2 > n1 || hy + 2 > N0) {
      s0 = max(-2, -hx); s1 = max(-2, -hy);
      e0 = min(2, n1 - hx); e1 = min(2, N0 - hy);
    }
  }
  for (int sy = s1; sy < e1; ++sy) {
    int row = hy + sy;
    if (periodic) row = (row + N0) % N0;
    row -= row_off;
    if (row < 0 || row >= n0) continue;
    for (int sx = s0; sx < e0; ++sx) {
      int col = hx + sx;
      if (periodic) col = (col + n1) % n1;
      atomicAdd(&mesh[(long long)row * ld + col], (wy[sy + 2] * wx[sx + 2]) * v);
    }
  }
}

__global__ void k_p2m(int n0, int n1, const double* __restrict__ px, const double* __restrict__ py,
                      const double* __restrict__ val, double* mesh, double dx, double dy, bool periodic) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (k >= n1) return;
  const long long id = (long long)j * n1 + k;
  double wx[4], wy[4];
  const int hx = mp4_axis(px[id], dx, wx);
  const int hy = mp4_axis(py[id], dy, wy);
  mp4_scatter(mesh, n0, n1, n1, n0, 0, hx, hy, wx, wy, val[id], periodic);
}

// kernels/advect_particle.py:18-35 for lattice particles: doubled row jd in [0, 2nr)
__global__ void k_p2m_lattice(GridD g, double* w_out, const double* __restrict__ w_in, const double* __restrict__ u_z,
                              const double* __restrict__ u_r, const double* __restrict__ zl, const double* __restrict__ rl,
                              double dt, const double* __restrict__ dt_dev, bool periodic) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int jd = blockIdx.y;
  if (k >= g.nz) return;
  {
    const long long fo = member_field(g);
    w_out += fo; w_in += fo; u_z += fo; u_r += fo;
  }
  if (dt_dev) dt = dt_dev[member_scalar(g)];
  const bool upper = jd >= g.nr;
  const int j = upper ? jd - g.nr : g.nr - 1 - jd;
  const long long src = (long long)j * g.ld + k;
  const double uz = u_z[src];
  const double ur = upper ? u_r[src] : -u_r[src];
  const double v = upper ? w_in[src] : -w_in[src];
  const double pz = zl[k] + uz * dt;
  const double pr = rl[jd] + ur * dt;
  double wx[4], wy[4];
  const int hx = mp4_axis(pz, g.dx, wx);
  const int hy = mp4_axis(pr, g.dx, wy);
  mp4_scatter(w_out, g.nr, g.nz, g.ld, 2 * g.nr, g.nr, hx, hy, wx, wy, v, periodic);
}


// ---- gather form of the lattice remesh (no atomics) --------------------------------------------------------------
// A lattice particle that moves less than one cell (|u dt| < dx in both directions; the driver's dt keeps it far
// below that) lands with its nearest-upper index hi in {i, i + 1}, so its 4 x 4 MP4 footprint lies inside the 5 x 5
// nodes around its own lattice node, and node (R, c) only receives from the particles of rows R-2..R+2, columns
// c-2..c+2.  A warp takes 32 adjacent particle columns and marches down the particle rows: every lane evaluates its
// particle once (weights as in mp4_axis, zero-padded to 5 per axis), the 25 products (wy wx) v travel to the lanes of
// their target columns by warp shuffles and are added to five rolling row accumulators; the accumulator of row
// jd - 2 is complete after particle row jd and is stored.  Lanes 2..29 own an output column (warps overlap by 4
// columns, row chunks by 4 particle rows).  Contributions reach every node in the order the reference's sequential
// loop produces them (particle rows ascending -- the mirror half first --, columns ascending; particles_to_mesh_2D.hpp
// :273-321), with the same product expression, so the result is the reference's bit for bit -- the atomic scatter is
// not.  Particles that move a cell or more ("far") are left out here, flagged, and scattered by k_p2m_lattice_far.
constexpr int PG_OUT = 28;             // output columns per warp
constexpr int PG_WARPS = 4;
constexpr double PG_FAR = 0.99;        // |u dt| >= PG_FAR dx -> far particle (margin against rounding of pos / dx)

__device__ __forceinline__ void pad5(const double w[4], int d, double o[5]) {   // d = hi - i in {0, 1}
  o[0] = d ? 0.0 : w[0];
  o[1] = d ? w[0] : w[1];
  o[2] = d ? w[1] : w[2];
  o[3] = d ? w[2] : w[3];
  o[4] = d ? w[3] : 0.0;
}
__device__ __forceinline__ double shfl_from(double v, int delta) {   // value of lane (own lane - delta)
  return delta > 0 ? __shfl_up_sync(0xffffffffu, v, delta) : (delta < 0 ? __shfl_down_sync(0xffffffffu, v, -delta) : v);
}

__global__ void __launch_bounds__(32 * PG_WARPS, 8)
    k_p2m_lattice_gather(GridD g, int RB, double* __restrict__ w_out, const double* __restrict__ w_in,
                         const double* __restrict__ u_z, const double* __restrict__ u_r, const double* __restrict__ zl,
                         const double* __restrict__ rl, double dt, const double* __restrict__ dt_dev, int* far_flag) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * PG_WARPS + warp) * PG_OUT;       // first output column of the warp
  if (c0 >= g.nz) return;
  const int i = c0 - 2 + lane;                                   // particle column of the lane (= its output column)
  const bool col_ok = i >= 0 && i < g.nz;
  const int j0 = blockIdx.y * RB, j1 = min(j0 + RB, g.nr);
  {
    const long long fo = member_field(g);
    w_out += fo; w_in += fo; u_z += fo; u_r += fo;
  }
  if (dt_dev) dt = dt_dev[member_scalar(g)];
  const double zi = col_ok ? zl[i] : 0.0;
  const double far = PG_FAR * g.dx;
  const bool owner = lane >= 2 && lane < 2 + PG_OUT && col_ok;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};                      // output rows jd-2 .. jd+2 (doubled-grid rows)
  const int jd_first = g.nr + j0 - 2, jd_last = g.nr + j1 + 1;
  // particle row jd: source row j of the physical half, mirrored (sign -1) below the axis
  // RAW loads only: the mirror signs are applied when the values are used, one iteration later -- a negation inside the
  // fetch made the warp wait for the loads it had just issued (40 % of the stall samples in ncu's per-instruction view)
  auto fetch = [&](int jd, double& uz, double& ur, double& v) {
    uz = 0.0; ur = 0.0; v = 0.0;
    if (jd < 0 || jd >= 2 * g.nr || !col_ok) return;
    const long long src = (long long)(jd >= g.nr ? jd - g.nr : g.nr - 1 - jd) * g.ld + i;
    uz = u_z[src];
    ur = u_r[src];
    v = w_in[src];
  };
  double uz_n, ur_n, v_n;
  fetch(jd_first, uz_n, ur_n, v_n);
  bool any_far = false;
  for (int jd = jd_first; jd <= jd_last; ++jd) {
    const bool upper = jd >= g.nr;
    const double uz = uz_n, ur = upper ? ur_n : -ur_n, v = upper ? v_n : -v_n;
    fetch(jd + 1 <= jd_last ? jd + 1 : -1, uz_n, ur_n, v_n);      // next row's loads fly during this row's arithmetic
    double wxp[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, wyp[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double val = 0.0;
    const bool live = col_ok && jd >= 0 && jd < 2 * g.nr;
    if (live) {
      const double mz = uz * dt, mr = ur * dt;
      if (fabs(mz) >= far || fabs(mr) >= far) {
        // recorded once: by the lane that owns the column, for the physical-half row of the block's own rows
        if (owner && jd >= g.nr + j0 && jd < g.nr + j1) any_far = true;
      } else {
        double wx[4], wy[4];
        const int hx = mp4_axis(zi + mz, g.dx, wx);
        const int hy = mp4_axis(rl[jd] + mr, g.dx, wy);
        pad5(wx, hx - i, wxp);
        pad5(wy, hy - jd, wyp);
        val = v;
      }
    }
#pragma unroll
    for (int dr = 0; dr < 5; ++dr) {
#pragma unroll
      for (int dc = 4; dc >= 0; --dc) {                           // source columns ascending: i = c-2 .. c+2
        const double t = (wyp[dr] * wxp[dc]) * val;               // particles_to_mesh_2D.hpp:316-318
        acc[dr] = acc[dr] + shfl_from(t, dc - 2);
      }
    }
    const int R = jd - 2 - g.nr;                                  // completed physical row
    if (owner && R >= j0 && R < j1) w_out[(long long)R * g.ld + i] = acc[0];
#pragma unroll
    for (int dr = 0; dr < 4; ++dr) acc[dr] = acc[dr + 1];
    acc[4] = 0.0;
  }
  if (far_flag && any_far) *far_flag = 1;
}

// far particles (|u dt| >= PG_FAR dx) of the gather form: atomic scatter of the particle and of its mirror image.
// Skipped when the gather pass found none (far_flag given and zero).
__global__ void k_p2m_lattice_far(GridD g, double* w_out, const double* __restrict__ w_in, const double* __restrict__ u_z,
                                  const double* __restrict__ u_r, const double* __restrict__ zl, const double* __restrict__ rl,
                                  double dt, const double* __restrict__ dt_dev, const int* far_flag) {
  if (far_flag && *far_flag == 0) return;
  {
    const long long fo = member_field(g);
    w_out += fo; w_in += fo; u_z += fo; u_r += fo;
  }
  if (dt_dev) dt = dt_dev[member_scalar(g)];
  const double far = PG_FAR * g.dx;
  for (int j = blockIdx.y; j < g.nr; j += gridDim.y) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < g.nz; k += gridDim.x * blockDim.x) {
      const long long src = (long long)j * g.ld + k;
      const double uz = u_z[src], ur = u_r[src];
      if (!(fabs(uz * dt) >= far || fabs(ur * dt) >= far)) continue;
      const double v = w_in[src];
      const double pz = zl[k] + uz * dt;
      double wx[4], wy[4];
      const int hx = mp4_axis(pz, g.dx, wx);
      int hy = mp4_axis(rl[g.nr + j] + ur * dt, g.dx, wy);
      mp4_scatter(w_out, g.nr, g.nz, g.ld, 2 * g.nr, g.nr, hx, hy, wx, wy, v, false);
      hy = mp4_axis(rl[g.nr - 1 - j] + (-ur) * dt, g.dx, wy);
      mp4_scatter(w_out, g.nr, g.nz, g.ld, 2 * g.nr, g.nr, hx, hy, wx, wy, -v, false);
    }
  }
}

}  // namespace

extern "C" {

int64_t axb_ls_workspace_bytes(int n0, int n1) {
  // pending-list capacity: a quarter of the array (the reference reserves 0.15) -- the call
  // reports AXB_EWORK if a wider band is requested; doubled float/flag staging for the
  // fused wrapper is appended (20 bytes per cell).
  const long long n = (long long)n0 * n1;
  long long cap = n / 4 + 4096;
  return 512 + cap * 14 + 256 + n * 20;
}

int axb_ls_extrapolate_order1(int n0, int n1, int16_t* cur, const int16_t* tgt, double* eta_x,
                              double* eta_y, const double* gx, const double* gy, void* work,
                              int64_t work_bytes, int max_sweeps, int* sweeps_host, axb_stream_t s) {
  if (!cur || !tgt || !eta_x || !eta_y || !gx || !gy || !work || n0 < 3 || n1 < 3) return AXB_EINVAL;
  return ls_run(1, n0, n1, cur, tgt, eta_x, eta_y, gx, gy, work, work_bytes, max_sweeps, sweeps_host, (cudaStream_t)s);
}

int axb_ls_extrapolate_order2(int n0, int n1, int16_t* cur, const int16_t* tgt, double* eta_x,
                              double* eta_y, const double* gx, const double* gy, void* work,
                              int64_t work_bytes, int max_sweeps, int* sweeps_host, axb_stream_t s) {
  if (!cur || !tgt || !eta_x || !eta_y || !gx || !gy || !work || n0 < 3 || n1 < 3) return AXB_EINVAL;
  return ls_run(2, n0, n1, cur, tgt, eta_x, eta_y, gx, gy, work, work_bytes, max_sweeps, sweeps_host, (cudaStream_t)s);
}

static int ls_eta(const axb_grid_t* g, const double* ball_phi, const uint8_t* inside_solid, const double* eta1_in,
                  const double* eta2_in, double* eta1, double* eta2, double extrap_zone, const double* gx,
                  const double* gy, void* work, int64_t work_bytes, int max_sweeps, int* sweeps_host, int* status_dev,
                  bool device_loop, axb_stream_t s, int parts = 7) {
  if (!ball_phi || !inside_solid || !eta1 || !eta2 || !eta1_in || !eta2_in || !gx || !gy || !work) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  if (g->ku0 != 0 || g->ku1 != g->nz || g->nz_global != g->nz) return AXB_ENOSUP;  // single-slab only
  const int nr = g->nr, nz = g->nz;
  const long long n = 2LL * nr * nz;
  // staging: e1, e2 (double), cur, tgt (int16) on the doubled grid, 256-byte aligned
  char* p = (char*)work;
  const long long stage = n * 20;
  if (work_bytes < stage + 256 + 256 + 14) return AXB_EWORK;
  double* e1 = (double*)p;
  double* e2 = e1 + n;
  short* cur = (short*)(e2 + n);
  short* tgt = cur + n;
  char* rest = (char*)(tgt + n);
  rest = (char*)(((uintptr_t)rest + 255) & ~(uintptr_t)255);
  const long long rest_bytes = work_bytes - (rest - p);
  dim3 grd((nz + 127) / 128, nr);
  if (device_loop) {
    // parts (device loop only): 1 = fill of the doubled work arrays + first pending list, 2 = the sweeps, 4 = write-back
    const LsDevWork w = carve_device(rest, rest_bytes);
    if (w.cap < 2) return AXB_EWORK;
    if (parts & 1) {
      cudaMemsetAsync(w.ctr, 0, 256, s);
      k_ls_fill<true><<<grd, 128, 0, s>>>(nr, nz, g->ld, ball_phi, inside_solid, eta1_in, eta2_in, extrap_zone, cur, tgt,
                                          e1, e2, w.pend[0], w.ctr + 4, w.cap);
      AXB_LAUNCHED();
    }
    if (parts & 2)
      rc = ls_run_device(1, 2 * nr, nz, cur, tgt, e1, e2, gx, gy, rest, rest_bytes, max_sweeps, status_dev, true,
                         (cudaStream_t)s);
  } else {
    k_ls_fill<false><<<grd, 128, 0, s>>>(nr, nz, g->ld, ball_phi, inside_solid, eta1_in, eta2_in, extrap_zone, cur, tgt, e1,
                                         e2, nullptr, nullptr, 0);
    AXB_LAUNCHED();
    rc = ls_run(1, 2 * nr, nz, cur, tgt, e1, e2, gx, gy, rest, rest_bytes, max_sweeps, sweeps_host, (cudaStream_t)s);
  }
  if (rc) return rc;
  if (parts & 4) {
    k_ls_unfill<<<grd, 128, 0, s>>>(nr, nz, g->ld, eta1, eta2, e1, e2);
    AXB_LAUNCHED();
  }
  AXB_RETURN_LAST();
}

int axb_ls_extrapolate_eta(const axb_grid_t* g, const double* ball_phi, const uint8_t* inside_solid,
                           double* eta1, double* eta2, double extrap_zone, const double* gx,
                           const double* gy, void* work, int64_t work_bytes, int max_sweeps,
                           int* sweeps_host, axb_stream_t s) {
  return ls_eta(g, ball_phi, inside_solid, eta1, eta2, eta1, eta2, extrap_zone, gx, gy, work, work_bytes, max_sweeps,
                sweeps_host, nullptr, false, s);
}
int axb_ls_extrapolate_eta_device(const axb_grid_t* g, const double* ball_phi, const uint8_t* inside_solid,
                                  const double* eta1_in, const double* eta2_in, double* eta1_out, double* eta2_out,
                                  double extrap_zone, const double* gx, const double* gy, void* work,
                                  int64_t work_bytes, int sweeps, int32_t* status_dev, axb_stream_t s) {
  return ls_eta(g, ball_phi, inside_solid, eta1_in, eta2_in, eta1_out, eta2_out, extrap_zone, gx, gy, work, work_bytes,
                sweeps, nullptr, status_dev, true, s);
}
int axb_ls_extrapolate_eta_device_parts(const axb_grid_t* g, const double* ball_phi, const uint8_t* inside_solid,
                                        const double* eta1_in, const double* eta2_in, double* eta1_out, double* eta2_out,
                                        double extrap_zone, const double* gx, const double* gy, void* work,
                                        int64_t work_bytes, int sweeps, int32_t* status_dev, int parts, axb_stream_t s) {
  if (parts < 1 || parts > 7) return AXB_EINVAL;
  return ls_eta(g, ball_phi, inside_solid, eta1_in, eta2_in, eta1_out, eta2_out, extrap_zone, gx, gy, work, work_bytes,
                sweeps, nullptr, status_dev, true, s, parts);
}

int axb_p2m_mp4_2d(int n0, int n1, const double* px, const double* py, const double* val, double* mesh,
                   double dx, double dy, int periodic, axb_stream_t s) {
  if (!px || !py || !val || !mesh || n0 < 1 || n1 < 1) return AXB_EINVAL;
  cudaMemsetAsync(mesh, 0, sizeof(double) * (size_t)n0 * n1, s);
  k_p2m<<<dim3((n1 + 127) / 128, n0), 128, 0, s>>>(n0, n1, px, py, val, mesh, dx, dy, periodic != 0);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

static int g_p2m_atomic = 0;
int axb_set_p2m_atomic(int on) {
  g_p2m_atomic = on ? 1 : 0;
  return AXB_OK;
}

static int advect_particles(const axb_grid_t* g, double* w_out, const double* w_in, const double* u_z, const double* u_r,
                            const double* zl1d, const double* rl1d, double dt, const double* dt_dev, int periodic,
                            int32_t* far_flag, axb_stream_t s) {
  if (!w_out || !w_in || !u_z || !u_r || !zl1d || !rl1d || w_out == w_in) return AXB_EINVAL;
  int rc = axb_check_grid_batched(g);
  if (rc) return rc;
  if (g->ku0 != 0 || g->ku1 != g->nz || g->nz_global != g->nz) return AXB_ENOSUP;
  const GridD d = to_dev(g);
  if (!periodic && !g_p2m_atomic && d.nr >= 2) {
    // gather form: every node written exactly once (no memset), reference summation order
    if (far_flag) cudaMemsetAsync(far_flag, 0, sizeof(int32_t), s);
    const int rb = 64;
    const dim3 grd((d.nz + PG_OUT * PG_WARPS - 1) / (PG_OUT * PG_WARPS), (d.nr + rb - 1) / rb, d.batch);
    k_p2m_lattice_gather<<<grd, 32 * PG_WARPS, 0, s>>>(d, rb, w_out, w_in, u_z, u_r, zl1d, rl1d, dt, dt_dev, far_flag);
    AXB_LAUNCHED();
    const dim3 fg(min((d.nz + 127) / 128, 16), min(d.nr, 64), d.batch);
    k_p2m_lattice_far<<<fg, 128, 0, s>>>(d, w_out, w_in, u_z, u_r, zl1d, rl1d, dt, dt_dev, far_flag);
    AXB_LAUNCHED();
    AXB_RETURN_LAST();
  }
  if (d.batch > 1 && d.bstride == d.nz) {          // members are adjacent column blocks: one memset
    cudaMemset2DAsync(w_out, d.ld * sizeof(double), 0, (size_t)d.batch * d.nz * sizeof(double), d.nr, s);
  } else {
    for (int m = 0; m < d.batch; ++m)
      cudaMemset2DAsync(w_out + m * d.bstride, d.ld * sizeof(double), 0, d.nz * sizeof(double), d.nr, s);
  }
  k_p2m_lattice<<<dim3((d.nz + 127) / 128, 2 * d.nr, d.batch), 128, 0, s>>>(d, w_out, w_in, u_z, u_r, zl1d, rl1d, dt,
                                                                           dt_dev, periodic != 0);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_advect_vorticity_particles(const axb_grid_t* g, double* w_out, const double* w_in, const double* u_z,
                                   const double* u_r, const double* zl1d, const double* rl1d, double dt,
                                   const double* dt_dev, int periodic, axb_stream_t s) {
  return advect_particles(g, w_out, w_in, u_z, u_r, zl1d, rl1d, dt, dt_dev, periodic, nullptr, s);
}
int axb_advect_vorticity_particles_flagged(const axb_grid_t* g, double* w_out, const double* w_in, const double* u_z,
                                           const double* u_r, const double* zl1d, const double* rl1d, double dt,
                                           const double* dt_dev, int32_t* far_flag, axb_stream_t s) {
  if (!far_flag) return AXB_EINVAL;
  return advect_particles(g, w_out, w_in, u_z, u_r, zl1d, rl1d, dt, dt_dev, 0, far_flag, s);
}

}  // extern "C"
