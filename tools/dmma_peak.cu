// Microbenchmark: peak FP64 DMMA.8x8x4 and DFMA issue rate on one B200 (denominator for the FP64 roofline).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void dmma_loop(double* out, int iters, double a0, double b0) {
  double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void dfma_loop(double* out, int iters, double a0, double b0) {
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void fill(double* p, size_t n, double v) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2;
  size_t stride = (size_t)gridDim.x * blockDim.x * 2;
  for (; i + 1 < n; i += stride) *reinterpret_cast<double2*>(p + i) = make_double2(v, v);
}
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("device %s SMs %d clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 32 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  for (int wps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      dmma_loop<8><<<148, wps * 32>>>(out, iters, 1.0, 0.5);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 256 * 8 * (double)iters * wps * 148;
      if (rep) printf("DMMA.8x8x4 warps/SM=%2d: %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
    }
  }
  for (int wps : {8, 16, 32}) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      dfma_loop<8><<<148, wps * 32>>>(out, iters, 1.0000001, 0.5);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 32 * 8 * (double)iters * wps * 148;
      if (rep) printf("DFMA warps/SM=%2d: %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
    }
  }
  size_t n = (size_t)21000 * 21000; double* buf; cudaMalloc(&buf, n * 8);
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); fill<<<148 * 8, 512>>>(buf, n, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("fill 8*N^2 (N=21000) %.3f ms  %.1f GB/s\n", ms, n * 8.0 / ms * 1e-6);
  }
  cudaError_t err = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(err));
  return 0;
}
