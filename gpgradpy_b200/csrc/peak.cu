// Roofline denominator measured in the run: issue peak of the fp64 tensor-core instruction every O(N^3) kernel of
// this library is built from (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).  MEASURED_PEAKS.json carries no fp64 entry,
// so bench.py calls gegp_dmma_peak() on the GPU it is timing instead of quoting a constant.
#include "common.cuh"
#include "../../include/gegp.h"

namespace gegp {
namespace {
template <int NACC>
__global__ void __launch_bounds__(512) dmma_issue_kernel(double* out, int iters, double a0, double b0) {
  const double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace
}  // namespace gegp

extern "C" int gegp_dmma_peak(double* scratch, size_t scratch_doubles, int reps, double* tflops_out, void* stream) {
  using namespace gegp;
  if (!scratch) return -1;
  if (reps < 1) return -3;
  if (!tflops_out) return -4;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return -1001;
  constexpr int WARPS = 16, NACC = 8, ITERS = 20000;   // 16 warps per SM: the issue rate has saturated (8 already does)
  if (scratch_doubles < (size_t)sms * WARPS * 32) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1002;
  double best = 0.0;
  for (int r = 0; r <= reps; r++) {   // r = 0 is a warm-up
    cudaEventRecord(e0, st);
    dmma_issue_kernel<NACC><<<sms, WARPS * 32, 0, st>>>(scratch, ITERS, 1.0, 0.5);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return -1003; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 256 * NACC * (double)ITERS * WARPS * sms / ms * 1e-9;   // m8n8k4: 256 multiply-adds per warp
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return 0;
}
