// axb_march.cuh -- launchers of the row-marching stencil kernels (stencils_march.cu)
#pragma once
#include "axb_common.cuh"

int march_velocity(const GridD& d, double* u_z, double* u_r, const double* psi, const double* r1d, double uz_add,
                   double ur_add, const double* add_dev, double* umax_out, bool vec, cudaStream_t s);
int march_penalise(const GridD& d, double* u_z, double* u_r, double* w, const double* uzu, const double* uru,
                   const double* chi, double lam, double dt, const double* dt_dev, double U_z, double U_r,
                   const double* U_dev, const double* r1d, double* sum_out, bool vec, cudaStream_t s);
int march_diffusion(int stage, const GridD& d, double* out, const double* in, const double* src2, const double* r1d,
                    double nu, const double* nu_dev, double dt, const double* dt_dev, bool vec, cudaStream_t s);
int march_diffusion_fused(const GridD& d, double* out, const double* in, const double* r1d, double nu, double dt,
                          const double* dt_dev, bool vec, cudaStream_t s);
int march_eno3(int nf, bool cons, bool mirror, bool fluxonly, const GridD& d, double* out0, double* out1,
               const double* in0, const double* in1, const double* u_z, const double* u_r, double inv_dx, double dt,
               const double* dt_dev, double sign0, double sign1, bool vec, cudaStream_t s);
extern int g_axb_legacy_stencils;   // 1: use the 2-D tiled kernels of stencils.cu / eno3.cu instead
