// pde_extrap.cu -- static-PDE (Aslam-type) extrapolation of a field out of a solid, the alternative to the
// least-squares wavefront used by the periodic soft-slab driver
// (examples/PeriodicSoftSlab/bounded_static_PDE_extrapolation.py:5-235).
//
//   axb_pde_extrap_setup   :142-177 + :66-93 -- zones, upwind unit normal of the (negated) level set, its positive /
//                          negative parts, the Jacobi denominator and, inside the solid, the normal derivative of eta
//   axb_pde_extrap_jacobi  :185-218 -- Jacobi sweeps of  n . grad(soln) = rhs  over the extrapolation zone until the
//                          2-norm of an update drops to the tolerance
//
// All arrays are the reference's *bounded* arrays, dense (n0, n1); "interim" arrays are their (n0-4, n1-4) interiors.
// The reference's while-loop is data dependent.  Here it terminates on the device: every sweep kernel first reads a
// `done` flag, the last block of a sweep (atomic ticket) turns the accumulated sum of squares into the residual and
// sets the flag when it is <= tol, so the host may enqueue sweeps in batches and read the 16-byte state once per
// batch; sweeps launched after convergence return immediately.  Sweeps ping-pong between two buffers (the parity of
// the sweep count says which one is current), which is exactly Jacobi: every update reads the previous sweep only.
// -fmad=false and the reference's operation order: results equal NumPy's up to the order of the residual sum.
#include "axb_common.cuh"

namespace {

struct PdeState {
  double sumsq;
  unsigned int ticket;
  int done;
  int sweeps;
  int pad;
};

__global__ void k_pde_setup(int n0, int n1, const double* __restrict__ phi, const double* __restrict__ eta, double dx,
                            double offset, double band, double eps, double* __restrict__ nrp, double* __restrict__ nrn,
                            double* __restrict__ nzp, double* __restrict__ nzn, double* __restrict__ denom,
                            unsigned char* __restrict__ zone, double* __restrict__ gn) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= n0 || k >= n1) return;
  const long long c = (long long)j * n1 + k;
  if (j < 2 || k < 2 || j >= n0 - 2 || k >= n1 - 2) {
    gn[c] = 0.0;                                                     // grad_eta_n = bounded_eta * 0 outside the interim
    return;
  }
  const long long i = (long long)(j - 2) * (n1 - 4) + (k - 2);
  const double p = phi[c];
  const bool inside = p <= -offset;
  const bool ext = (p > -offset) && (p < band);
  zone[i] = (unsigned char)((inside ? 1 : 0) | (ext ? 2 : 0));
  const double h = 2 * dx;
  double nr = (fmax(3.0 * p - 4.0 * phi[c - n1] + phi[c - 2 * n1], 0.0) +
               fmin(-3.0 * p + 4.0 * phi[c + n1] - phi[c + 2 * n1], 0.0)) / h;
  double nz = (fmax(3.0 * p - 4.0 * phi[c - 1] + phi[c - 2], 0.0) + fmin(-3.0 * p + 4.0 * phi[c + 1] - phi[c + 2], 0.0)) / h;
  const double mag = sqrt(nr * nr + nz * nz);
  nr /= mag + eps;
  nz /= mag + eps;
  const double gr = (eta[c + n1] - eta[c - n1]) / h, gz = (eta[c + 1] - eta[c - 1]) / h;
  gn[c] = (inside ? 1.0 : 0.0) * (nr * gr + nz * gz);
  const double rp = (nr >= 0 ? 1.0 : 0.0) * nr, zp = (nz >= 0 ? 1.0 : 0.0) * nz;
  nrp[i] = rp;
  nrn[i] = nr - rp;
  nzp[i] = zp;
  nzn[i] = nz - zp;
  denom[i] = 3.0 * (fabs(nr) + fabs(nz)) + eps;
}

// original = soln[interim] * inside (bounded_static_PDE_extrapolation.py:198), and the second buffer starts as a
// copy of the first (its 2-wide rim never changes)
__global__ void k_pde_begin(int n0, int n1, const double* __restrict__ a, double* __restrict__ b,
                            const unsigned char* __restrict__ zone, double* __restrict__ original, PdeState* st) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (k == 0 && j == 0) {
    st->sumsq = 0.0;
    st->ticket = 0u;
    st->done = 0;
    st->sweeps = 0;
  }
  if (j >= n0 || k >= n1) return;
  const long long c = (long long)j * n1 + k;
  const double v = a[c];
  b[c] = v;
  if (j >= 2 && k >= 2 && j < n0 - 2 && k < n1 - 2)
    original[(long long)(j - 2) * (n1 - 4) + (k - 2)] = v * ((zone[(long long)(j - 2) * (n1 - 4) + (k - 2)] & 1) ? 1.0 : 0.0);
}

__global__ void __launch_bounds__(256)
    k_pde_sweep(int n0, int n1, double* bufA, double* bufB, const double* __restrict__ rhs,
                const double* __restrict__ original, const unsigned char* __restrict__ zone,
                const double* __restrict__ denom, const double* __restrict__ nrp, const double* __restrict__ nrn,
                const double* __restrict__ nzp, const double* __restrict__ nzn, double dx, double tol, PdeState* st) {
  if (*reinterpret_cast<volatile int*>(&st->done)) return;
  const bool odd = (*reinterpret_cast<volatile int*>(&st->sweeps)) & 1;
  const double* __restrict__ s = odd ? bufB : bufA;
  double* __restrict__ d = odd ? bufA : bufB;
  const int k = 2 + blockIdx.x * blockDim.x + threadIdx.x, j = 2 + blockIdx.y * blockDim.y + threadIdx.y;
  double sq = 0.0;
  if (j < n0 - 2 && k < n1 - 2) {
    const long long c = (long long)j * n1 + k, i = (long long)(j - 2) * (n1 - 4) + (k - 2);
    const double num_r = nrp[i] * (4.0 * s[c - n1] - s[c - 2 * n1]) - nrn[i] * (4.0 * s[c + n1] - s[c + 2 * n1]);
    const double num_z = nzp[i] * (4.0 * s[c - 1] - s[c - 2]) - nzn[i] * (4.0 * s[c + 1] - s[c + 2]);
    const double r = rhs ? rhs[c] : 0.0;
    const double z = (zone[i] & 2) ? 1.0 : 0.0;
    const double v = original[i] + z * (2 * dx * r + num_r + num_z) / denom[i];
    d[c] = v;
    const double diff = v - s[c];
    sq = diff * diff;
  }
  sq = block_sum(sq);
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    atomicAdd(&st->sumsq, sq);
    __threadfence();
    const unsigned int total = gridDim.x * gridDim.y;
    if (atomicAdd(&st->ticket, 1u) == total - 1) {                    // last block of this sweep
      __threadfence();
      const double res = sqrt(*reinterpret_cast<volatile double*>(&st->sumsq));
      st->sumsq = 0.0;
      st->ticket = 0u;
      st->sweeps = st->sweeps + 1;
      if (!(res > tol)) st->done = 1;                                 // while residual > tol
      __threadfence();
    }
  }
}

// after convergence: make bufA hold the current iterate (copy back when the sweep count is odd)
__global__ void k_pde_finish(long long n, double* bufA, const double* __restrict__ bufB, const PdeState* st) {
  if (!(st->sweeps & 1)) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bufA[i] = bufB[i];
}

}  // namespace

extern "C" {

int64_t axb_pde_extrap_workspace_bytes(int n0, int n1) {
  if (n0 < 6 || n1 < 6) return 0;
  const long long nb = (long long)n0 * n1, ni = (long long)(n0 - 4) * (n1 - 4);
  // second sweep buffer + original (interim) + state
  return (int64_t)((nb + ni) * 8 + 256);
}

int axb_pde_extrap_setup(int n0, int n1, const double* phi_b, const double* eta_b, double dx, double offset, double band,
                         double eps, double* nr_pos, double* nr_neg, double* nz_pos, double* nz_neg, double* denom,
                         uint8_t* zone, double* grad_eta_n, axb_stream_t s) {
  if (!phi_b || !eta_b || !nr_pos || !nr_neg || !nz_pos || !nz_neg || !denom || !zone || !grad_eta_n) return AXB_EINVAL;
  if (n0 < 6 || n1 < 6) return AXB_EINVAL;
  dim3 b(32, 8), g((n1 + 31) / 32, (n0 + 7) / 8);
  k_pde_setup<<<g, b, 0, (cudaStream_t)s>>>(n0, n1, phi_b, eta_b, dx, offset, band, eps, nr_pos, nr_neg, nz_pos, nz_neg,
                                            denom, zone, grad_eta_n);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_pde_extrap_jacobi(int n0, int n1, double* soln, const double* rhs, const uint8_t* zone, const double* denom,
                          const double* nr_pos, const double* nr_neg, const double* nz_pos, const double* nz_neg,
                          double dx, double tol, int max_sweeps, void* work, int64_t work_bytes, int* sweeps_host,
                          axb_stream_t s_) {
  if (!soln || !zone || !denom || !nr_pos || !nr_neg || !nz_pos || !nz_neg || !work) return AXB_EINVAL;
  if (n0 < 6 || n1 < 6) return AXB_EINVAL;
  if (work_bytes < axb_pde_extrap_workspace_bytes(n0, n1)) return AXB_EWORK;
  cudaStream_t s = (cudaStream_t)s_;
  const long long nb = (long long)n0 * n1, ni = (long long)(n0 - 4) * (n1 - 4);
  double* bufB = (double*)work;
  double* original = bufB + nb;
  PdeState* st = (PdeState*)(((uintptr_t)(original + ni) + 63) & ~(uintptr_t)63);
  dim3 b(32, 8), gb((n1 + 31) / 32, (n0 + 7) / 8), gi((n1 - 4 + 31) / 32, (n0 - 4 + 7) / 8);
  k_pde_begin<<<gb, b, 0, s>>>(n0, n1, soln, bufB, zone, original, st);
  AXB_LAUNCHED();
  PdeState h;
  h.done = 0;
  h.sweeps = 0;
  const int batch = 32;
  int launched = 0;
  while (!h.done && (max_sweeps <= 0 || launched < max_sweeps)) {
    for (int i = 0; i < batch; ++i) {
      k_pde_sweep<<<gi, b, 0, s>>>(n0, n1, soln, bufB, rhs, original, zone, denom, nr_pos, nr_neg, nz_pos, nz_neg, dx, tol,
                                   st);
    }
    g_axb_launches += batch;
    launched += batch;
    cudaMemcpyAsync(&h, st, sizeof(PdeState), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return (int)e;
  }
  k_pde_finish<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(nb, soln, bufB, st);
  AXB_LAUNCHED();
  if (sweeps_host) *sweeps_host = h.sweeps;
  AXB_RETURN_LAST();
}

}  // extern "C"
