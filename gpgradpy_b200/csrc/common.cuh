// Common device helpers for the gradient-enhanced GP kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>

// ---- optional instrumentation (bench.py roofline pass; off by default, no cost when off) ----
namespace gegp {
struct Prof {
  bool on = false;
  std::atomic<long> launches{0};   // every kernel launch of the library since the last reset
  double gemm_flops = 0.0;  // useful flops of the DMMA GEMM launches (2*M*N*K with triangular clipping)
  long gemm_launches = 0;
};
Prof& prof();
void prof_gemm_begin(cudaStream_t s);
void prof_gemm_end(cudaStream_t s, double flops);
// Debug timeline (env GEGP_TIMELINE=<file>): an event pair around a launch on its stream; dumped by gegp_profile_end
// as "label stream ready_us done_us" relative to the first mark.  No cost when the variable is unset.
bool timeline_on();
void timeline_begin(cudaStream_t s, const char* label, int a = 0, int b = 0, int c = 0, double mflop = 0.0);
void timeline_end(cudaStream_t s);
// look-ahead (multi-stream) factorisation switch: GEGP_OPT_LOOKAHEAD / env GEGP_NO_LOOKAHEAD
int& lookahead_enabled();
}  // namespace gegp

namespace gegp {

constexpr int GEGP_MAX_DYN_SMEM = 227 * 1024;  // largest shared memory (static + dynamic) a CTA can opt in to on sm_100
constexpr int LEAF = 128;  // blocking quantum of the recursive factorisation / inverse

// Launch context: every kernel of one C-ABI call goes to this stream; `batch` independent problems
// (multi-start candidates) are laid out with a fixed element stride and mapped to blockIdx.z.
struct Ctx {
  cudaStream_t stream;
  int batch;
};

#define GEGP_CHECK_LAUNCH()                                                                      \
  do {                                                                                           \
    ::gegp::prof().launches++;                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess) {                                                                    \
      fprintf(stderr, "[gegp] launch error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__,   \
              __LINE__);                                                                         \
      return -1000 - (int)e__;                                                                   \
    }                                                                                            \
  } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory, once per (kernel instantiation, device): the attribute
// is per device, so the flag is a per-device bit, not a process-wide bool (thread-safe: the worst case is a repeat).
#define GEGP_SET_SMEM(kern, bytes)                                                               \
  do {                                                                                           \
    static std::atomic<unsigned long long> done__{0ull};                                         \
    int dev__ = 0;                                                                               \
    cudaGetDevice(&dev__);                                                                       \
    const unsigned long long bit__ = 1ull << (dev__ & 63);                                       \
    if (!(done__.load(std::memory_order_acquire) & bit__)) {                                     \
      int want__ = (int)(bytes);                                                                 \
      if (want__ >= ::gegp::GEGP_MAX_DYN_SMEM) {   /* "as much as there is": leave room for the static part */ \
        cudaFuncAttributes fa__;                                                                 \
        if (cudaFuncGetAttributes(&fa__, kern) == cudaSuccess) want__ = ::gegp::GEGP_MAX_DYN_SMEM - (int)fa__.sharedSizeBytes; \
      }                                                                                          \
      cudaError_t e__ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, want__); \
      if (e__ != cudaSuccess) {                                                                  \
        (void)cudaGetLastError();   /* do not leave the error for the next launch check to find */ \
        fprintf(stderr, "[gegp] cannot set %d bytes of dynamic shared memory: %s (%s:%d)\n", (int)(bytes), \
                cudaGetErrorString(e__), __FILE__, __LINE__);                                    \
        return -1000 - (int)e__;                                                                 \
      }                                                                                          \
      done__.fetch_or(bit__, std::memory_order_release);                                         \
    }                                                                                            \
  } while (0)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4) * B(4x8), fp64 tensor-core MMA (SASS: DMMA.8x8x4).
// a: row = lane>>2, k = lane&3 ; b: k = lane&3, n = lane>>2 ; c: row = lane>>2, cols = 2*(lane&3)+{0,1}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in every thread. `sh` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0;
  for (int i = 0; i < nw; i++) t += sh[i];  // fixed order: deterministic
  return t;
}

}  // namespace gegp

