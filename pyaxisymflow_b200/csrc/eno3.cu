// eno3.cu -- ENO3 advection passes (fusion groups G-ADV and G-REF).
//
// Reference semantics restated once, in one kernel:
//   * pyst_kernels/advection_flux.py:26-127 (conservative) / :183-296 (non-conservative):
//     four face-flux passes; only [2:-2, 2:-2] of the (doubled) array is written;
//   * pyst_kernels/advection_timestep.py:45-54: flux = 0, flux += passes(inv_dx = -dt/dx),
//     field += flux;
//   * kernels/advect_vorticity_via_eno3.py:28-42 and
//     elasto_kernels/advect_refmap_via_eno3.py:27-52: the doubled array is the physical
//     field stacked on its axis mirror.  Here the mirror is an index map
//     (row j' < 0  ->  sign * f[-j'-1]), so no doubled arrays exist: 32 B/pt for the
//     vorticity (R w,u_z,u_r + W w), 48 B/pt for both reference maps.
// Every thread owns two adjacent z columns of one row; the three z-faces it needs are
// evaluated once and shared between its two cells (front of cell k == back of cell k+1
// bit for bit), r-neighbour rows come through L1.
#include <initializer_list>

#include "axb_common.cuh"
#include "axb_march.cuh"

namespace {

__device__ __forceinline__ const double* rowp(const double* f, long long ld, int j) { return f + (long long)j * ld; }
__device__ __forceinline__ double* rowp(double* f, long long ld, int j) { return f + (long long)j * ld; }

#define C13 (1.0 / 3.0)
#define C56 (5.0 / 6.0)
#define C16 (1.0 / 6.0)

// face between cell m and cell m+1 given q[m-1], q[m], q[m+1], q[m+2] and v[m], v[m+1]
__device__ __forceinline__ double eno_face(double qm1, double q0, double qp1, double qp2, double v0, double vp1) {
  return (v0 > -vp1) ? (C13 * qp1 + C56 * q0 - C16 * qm1) : (C13 * q0 + C56 * qp1 - C16 * qp2);
}

// Row fetch with the axis mirror: physical row index jj may be -2 or -1 when MIRROR.
template <bool MIRROR>
__device__ __forceinline__ double2 ld_row(const double* f, long long ld, int jj, int k, int nz, bool vec, double sign) {
  if (MIRROR && jj < 0) {
    const double2 v = ld_pair(rowp(f, ld, -jj - 1), k, nz, vec);
    return make_double2(sign * v.x, sign * v.y);
  }
  return ld_pair(rowp(f, ld, jj), k, nz, vec);
}

// NF fields advected by the same velocity.  CONS: conservative (q = f*v) else
// non-conservative (q = f, face value times the cell-centre velocity).
// MIRROR: rows [0, nr-3] are updated with reflected rows below the axis, else rows [2, nr-3].
// FLUXONLY: out += flux (pystencils flux kernel), else out = in + flux (Euler step; cells
// outside the update set are copied).
template <int NF, bool CONS, bool MIRROR, bool FLUXONLY>
__global__ void __launch_bounds__(TBX* TBY)
    k_eno3(GridD g, double* __restrict__ out0, double* __restrict__ out1, const double* __restrict__ in0,
           const double* __restrict__ in1, const double* __restrict__ u_z, const double* __restrict__ u_r,
           double inv_dx, double dt, const double* __restrict__ dt_dev, double sign0, double sign1, bool vec) {
  const int k = 2 * (blockIdx.x * TBX + threadIdx.x);
  const int j = blockIdx.y * TBY + threadIdx.y;
  if (j >= g.nr || k >= g.ku1 || k + 1 < g.ku0) return;
  if (dt_dev) dt = *dt_dev;
  if (!FLUXONLY) inv_dx = -(dt / g.dx);
  const int nz = g.nz;
  const int j_lo = MIRROR ? 0 : 2;
  const bool row_ok = (j >= j_lo) && (j <= g.nr - 3);
  // which of the two columns are advected (global columns 2 .. nzg-3)
  const int kg = k + g.kz0;
  const bool ok0 = row_ok && (k >= g.ku0) && (kg >= 2) && (kg <= g.nzg - 3);
  const bool ok1 = row_ok && (k + 1 < g.ku1) && (kg + 1 >= 2) && (kg + 1 <= g.nzg - 3);

  double* outs[2] = {out0, out1};
  const double* ins[2] = {in0, in1};
  const double signs[2] = {sign0, sign1};

  if (!ok0 && !ok1) {
    if (!FLUXONLY) {
#pragma unroll
      for (int f = 0; f < NF; ++f)
        st_pair(rowp(outs[f], g.ld, j), k, g.ku0, g.ku1, vec, ld_pair(rowp(ins[f], g.ld, j), k, nz, vec));
    }
    return;
  }

  // ---- velocities: u_z along z (columns k-2 .. k+3), u_r along r (rows j-2 .. j+2)
  const double* uzr = rowp(u_z, g.ld, j);
  double vz[6];
  {
    const double2 c = ld_pair(uzr, k, nz, vec);
    vz[2] = c.x; vz[3] = c.y;
    const int kl = (k >= 2) ? k - 2 : k;             // clamped: only read when ok
    const int kr = (k + 2 < nz) ? k + 2 : k;
    const double2 l = ld_pair(uzr, kl, nz, vec), r = ld_pair(uzr, kr, nz, vec);
    vz[0] = l.x; vz[1] = l.y; vz[4] = r.x; vz[5] = r.y;
  }
  double2 vr[5];
#pragma unroll
  for (int m = -2; m <= 2; ++m) {
    const int jj = clampi(j + m, MIRROR ? -2 : 0, g.nr - 1);
    vr[m + 2] = ld_row<MIRROR>(u_r, g.ld, jj, k, nz, vec, -1.0);
  }

#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const double* fin = ins[f];
    const double* fr = rowp(fin, g.ld, j);
    // z samples k-2 .. k+3
    double qz[6];
    {
      const double2 c = ld_pair(fr, k, nz, vec);
      qz[2] = c.x; qz[3] = c.y;
      const int kl = (k >= 2) ? k - 2 : k;
      const int kr = (k + 2 < nz) ? k + 2 : k;
      const double2 l = ld_pair(fr, kl, nz, vec), r = ld_pair(fr, kr, nz, vec);
      qz[0] = l.x; qz[1] = l.y; qz[4] = r.x; qz[5] = r.y;
    }
    const double c0 = qz[2], c1 = qz[3];
    // r samples j-2 .. j+2 (pairs)
    double2 qr[5];
#pragma unroll
    for (int m = -2; m <= 2; ++m) {
      if (m == 0) { qr[2] = make_double2(c0, c1); continue; }
      const int jj = clampi(j + m, MIRROR ? -2 : 0, g.nr - 1);
      qr[m + 2] = ld_row<MIRROR>(fin, g.ld, jj, k, nz, vec, signs[f]);
    }
    if (CONS) {
#pragma unroll
      for (int i = 0; i < 6; ++i) qz[i] *= vz[i];
#pragma unroll
      for (int i = 0; i < 5; ++i) { qr[i].x *= vr[i].x; qr[i].y *= vr[i].y; }
    }
    // z faces: F[0] = face(k-1|k), F[1] = face(k|k+1), F[2] = face(k+1|k+2)
    const double Fz0 = eno_face(qz[0], qz[1], qz[2], qz[3], vz[1], vz[2]);
    const double Fz1 = eno_face(qz[1], qz[2], qz[3], qz[4], vz[2], vz[3]);
    const double Fz2 = eno_face(qz[2], qz[3], qz[4], qz[5], vz[3], vz[4]);
    // r faces per column: back = face(j-1|j), front = face(j|j+1)
    const double Frb0 = eno_face(qr[0].x, qr[1].x, qr[2].x, qr[3].x, vr[1].x, vr[2].x);
    const double Frf0 = eno_face(qr[1].x, qr[2].x, qr[3].x, qr[4].x, vr[2].x, vr[3].x);
    const double Frb1 = eno_face(qr[0].y, qr[1].y, qr[2].y, qr[3].y, vr[1].y, vr[2].y);
    const double Frf1 = eno_face(qr[1].y, qr[2].y, qr[3].y, qr[4].y, vr[2].y, vr[3].y);

    double2 o;
    if (FLUXONLY) o = ld_pair(rowp(outs[f], g.ld, j), k, nz, vec);
    else o = make_double2(0.0, 0.0);
    double a0 = o.x, a1 = o.y;
    if (CONS) {
      a0 = a0 + inv_dx * Fz1; a0 = a0 - inv_dx * Fz0; a0 = a0 + inv_dx * Frf0; a0 = a0 - inv_dx * Frb0;
      a1 = a1 + inv_dx * Fz2; a1 = a1 - inv_dx * Fz1; a1 = a1 + inv_dx * Frf1; a1 = a1 - inv_dx * Frb1;
    } else {
      const double z0 = vz[2], z1 = vz[3], r0 = vr[2].x, r1 = vr[2].y;
      a0 = a0 + inv_dx * Fz1 * z0; a0 = a0 - inv_dx * Fz0 * z0; a0 = a0 + inv_dx * Frf0 * r0; a0 = a0 - inv_dx * Frb0 * r0;
      a1 = a1 + inv_dx * Fz2 * z1; a1 = a1 - inv_dx * Fz1 * z1; a1 = a1 + inv_dx * Frf1 * r1; a1 = a1 - inv_dx * Frb1 * r1;
    }
    if (FLUXONLY) {
      o.x = ok0 ? a0 : o.x;
      o.y = ok1 ? a1 : o.y;
    } else {
      o.x = ok0 ? (c0 + a0) : c0;
      o.y = ok1 ? (c1 + a1) : c1;
    }
    st_pair(rowp(outs[f], g.ld, j), k, g.ku0, g.ku1, vec, o);
  }
}

inline bool vec_ok(const GridD& g, std::initializer_list<const void*> ptrs) {
  if (g.ld & 1) return false;
  for (const void* p : ptrs)
    if (p && !axb_al16(p)) return false;
  return true;
}

}  // namespace

extern "C" {

int axb_advect_vorticity_eno3(const axb_grid_t* g, double* w_out, const double* w_in, const double* u_z,
                              const double* u_r, double dt, const double* dt_dev, axb_stream_t s) {
  if (!w_out || !w_in || !u_z || !u_r || w_out == w_in) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  if (d.nr < 5 || d.nzg < 5) return AXB_EINVAL;
  const bool vec = vec_ok(d, {w_out, w_in, u_z, u_r});
  if (!g_axb_legacy_stencils) {
    rc = march_eno3(1, true, true, false, d, w_out, nullptr, w_in, nullptr, u_z, u_r, 0.0, dt, dt_dev, -1.0, 0.0, vec, s);
    AXB_LAUNCHED();
    return rc;
  }
  k_eno3<1, true, true, false><<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, w_out, nullptr, w_in, nullptr, u_z, u_r, 0.0,
                                                                    dt, dt_dev, -1.0, 0.0, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_advect_refmap_eno3(const axb_grid_t* g, double* eta1_out, double* eta2_out, const double* eta1,
                           const double* eta2, const double* u_z, const double* u_r, double dt,
                           const double* dt_dev, axb_stream_t s) {
  if (!eta1_out || !eta2_out || !eta1 || !eta2 || !u_z || !u_r) return AXB_EINVAL;
  if (eta1_out == eta1 || eta2_out == eta2 || eta1_out == eta2 || eta2_out == eta1) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  if (d.nr < 5 || d.nzg < 5) return AXB_EINVAL;
  const bool vec = vec_ok(d, {eta1_out, eta2_out, eta1, eta2, u_z, u_r});
  if (!g_axb_legacy_stencils) {
    rc = march_eno3(2, false, true, false, d, eta1_out, eta2_out, eta1, eta2, u_z, u_r, 0.0, dt, dt_dev, +1.0, -1.0,
                    vec, s);
    AXB_LAUNCHED();
    return rc;
  }
  k_eno3<2, false, true, false><<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, eta1_out, eta2_out, eta1, eta2, u_z, u_r,
                                                                     0.0, dt, dt_dev, +1.0, -1.0, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_eno3_flux(const axb_grid_t* g, double* flux, const double* field, const double* vel0,
                  const double* vel1, double inv_dx, int conservative, axb_stream_t s) {
  if (!flux || !field || !vel0 || !vel1 || flux == field) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  if (d.nr < 5 || d.nzg < 5) return AXB_OK;  // empty interior: nothing to do (pystencils loops are empty)
  const bool vec = vec_ok(d, {flux, field, vel0, vel1});
  if (!g_axb_legacy_stencils) {
    rc = march_eno3(1, conservative != 0, false, true, d, flux, nullptr, field, nullptr, vel0, vel1, inv_dx, 0.0,
                    nullptr, 1.0, 1.0, vec, s);
    AXB_LAUNCHED();
    return rc;
  }
  if (conservative)
    k_eno3<1, true, false, true><<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, flux, nullptr, field, nullptr, vel0, vel1,
                                                                      inv_dx, 0.0, nullptr, 1.0, 1.0, vec);
  else
    k_eno3<1, false, false, true><<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, flux, nullptr, field, nullptr, vel0, vel1,
                                                                       inv_dx, 0.0, nullptr, 1.0, 1.0, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_eno3_euler_step(const axb_grid_t* g, double* field_out, const double* field_in, const double* vel0,
                        const double* vel1, double dt_by_dx, int conservative, axb_stream_t s) {
  if (!field_out || !field_in || !vel0 || !vel1 || field_out == field_in) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  GridD d = to_dev(g);
  const bool vec = vec_ok(d, {field_out, field_in, vel0, vel1});
  // the kernel forms inv_dx = -(dt / dx); feed dt = dt_by_dx with dx = 1
  d.dx = 1.0;
  if (d.nr < 5 || d.nzg < 5) {
    // no interior: the step is a plain copy (the tiled kernel handles that case)
  } else if (!g_axb_legacy_stencils) {
    rc = march_eno3(1, conservative != 0, false, false, d, field_out, nullptr, field_in, nullptr, vel0, vel1, 0.0,
                    dt_by_dx, nullptr, 1.0, 1.0, vec, s);
    AXB_LAUNCHED();
    return rc;
  }
  if (conservative)
    k_eno3<1, true, false, false><<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, field_out, nullptr, field_in, nullptr, vel0,
                                                                       vel1, 0.0, dt_by_dx, nullptr, 1.0, 1.0, vec);
  else
    k_eno3<1, false, false, false><<<grid2d(d), dim3(TBX, TBY), 0, s>>>(d, field_out, nullptr, field_in, nullptr,
                                                                        vel0, vel1, 0.0, dt_by_dx, nullptr, 1.0, 1.0,
                                                                        vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

}  // extern "C"
