// probe_dmma.cu -- register-resident throughput probe of the FP64 pipes on sm_100a:
// DFMA and every mma.sync f64 shape.  Prints TFLOP/s for 148 x {1,2,4} CTAs of 256 threads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_dmma tools/probe_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_dfma(double* out, double a, double b) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_884(double* out, double a, double b) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_1684(double* out, double a, double b) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_1688(double* out, double a, double b) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_16816(double* out, double a, double b) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
          "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
          : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
          : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F kern, double flops_per_thread_iter, double* out) {
  for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
    const int grid = 148 * ctas_per_sm, block = 256;
    kern<<<grid, block>>>(out, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) kern<<<grid, block>>>(out, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 5.0 * grid * block * (double)ITERS * flops_per_thread_iter;
    printf("%-10s ctas/SM=%d  %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, ctas_per_sm, ms / 5, flops / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 4 * 256);
  // per thread per iteration: DFMA 16*2; mma: (#mma * 2*M*N*K) / 32 lanes
  run("dfma", k_dfma, 16 * 2.0, out);
  run("m8n8k4", k_884, 16 * 2.0 * 8 * 8 * 4 / 32, out);
  run("m16n8k4", k_1684, 8 * 2.0 * 16 * 8 * 4 / 32, out);
  run("m16n8k8", k_1688, 8 * 2.0 * 16 * 8 * 8 / 32, out);
  run("m16n8k16", k_16816, 8 * 2.0 * 16 * 8 * 16 / 32, out);
  return 0;
}
