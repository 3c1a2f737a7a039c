// Probe: do green contexts (SM partitions) keep a latency-bound chain of small kernels away from the bulk GEMM CTAs?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/greenctx_probe tools/greenctx_probe.cu
// Measures the time of a chain of 40 dependent "leaf" kernels (1 CTA, 136 KB smem, ~20 us each) while a hog grid of
// long CTAs fills the machine on another stream: (a) priority streams in the primary context, (b) chain in an 8-SM
// green context and hog in the remainder, (c) the same captured into one CUDA graph and replayed.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s = nullptr; pcuGetErrorString(e, &s); printf("driver error %d (%s) at line %d\n", (int)e, s ? s : "?", __LINE__); return 2; } } while (0)

template <class F> static F entry(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    printf("no driver entry point %s\n", name);
    return nullptr;
  }
  return reinterpret_cast<F>(p);
}

__global__ void hog_kernel(long long cycles, unsigned* smids) {
  extern __shared__ double sm[];
  const long long t0 = clock64();
  double acc = threadIdx.x;
  while (clock64() - t0 < cycles) acc = acc * 1.0000001 + 1e-9;
  sm[threadIdx.x] = acc;
  if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smids[blockIdx.x] = s; }
}
__global__ void leaf_kernel(long long cycles, unsigned* smid_out, int idx) {
  extern __shared__ double sm[];
  const long long t0 = clock64();
  double acc = threadIdx.x;
  while (clock64() - t0 < cycles) acc = acc * 1.0000001 + 1e-9;
  sm[threadIdx.x] = acc;
  if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smid_out[idx] = s; }
}
__global__ void __cluster_dims__(4, 1, 1) cluster_kernel(long long cycles, unsigned* smid_out, int idx) {
  extern __shared__ double sm[];
  const long long t0 = clock64();
  double acc = threadIdx.x;
  while (clock64() - t0 < cycles) acc = acc * 1.0000001 + 1e-9;
  sm[threadIdx.x] = acc;
  if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smid_out[idx * 4 + blockIdx.x] = s; }
}

static unsigned *d_hog_smid, *d_leaf_smid, *d_cl_smid;
static const int NLEAF = 40;
static const int LEAF_SMEM = 136 * 1024, HOG_SMEM = 72 * 1024;

// one "factorisation": hog grid on `bulk`, chain of leaf + cluster kernels on `chain`; forked from / joined to `main`
static void enqueue(cudaStream_t main, cudaStream_t chain, cudaStream_t bulk, cudaEvent_t fork, cudaEvent_t j1, cudaEvent_t j2,
                    int hog_ctas) {
  CK(cudaEventRecord(fork, main));
  CK(cudaStreamWaitEvent(chain, fork, 0));
  CK(cudaStreamWaitEvent(bulk, fork, 0));
  hog_kernel<<<hog_ctas, 128, HOG_SMEM, bulk>>>(100000, d_hog_smid);   // ~50 us per CTA
  for (int i = 0; i < NLEAF; i++) {
    leaf_kernel<<<1, 128, LEAF_SMEM, chain>>>(40000, d_leaf_smid, i);      // ~20 us
    cluster_kernel<<<4, 128, LEAF_SMEM, chain>>>(20000, d_cl_smid, i);    // ~10 us
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(j1, chain));
  CK(cudaEventRecord(j2, bulk));
  CK(cudaStreamWaitEvent(main, j1, 0));
  CK(cudaStreamWaitEvent(main, j2, 0));
}

static void report(const char* what, float ms_chain, float ms_total) {
  std::vector<unsigned> ls(NLEAF), cs(NLEAF * 4), hs(148 * 16);
  CK(cudaMemcpy(ls.data(), d_leaf_smid, NLEAF * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cs.data(), d_cl_smid, NLEAF * 16, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hs.data(), d_hog_smid, 148 * 16 * 4, cudaMemcpyDeviceToHost));
  bool used[256] = {};
  bool hused[256] = {};
  for (unsigned s : ls) used[s & 255] = true;
  for (unsigned s : cs) used[s & 255] = true;
  for (unsigned s : hs) hused[s & 255] = true;
  int nu = 0, nh = 0, both = 0;
  for (int i = 0; i < 256; i++) { nu += used[i]; nh += hused[i]; both += used[i] && hused[i]; }
  printf("%-58s chain %.3f ms (%.1f us per step)  total %.3f ms   chain SMs %d, hog SMs %d, shared %d\n", what, ms_chain,
         ms_chain * 1e3 / NLEAF, ms_total, nu, nh, both);
}

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  CK(cudaMalloc(&d_hog_smid, 148 * 16 * 4));
  CK(cudaMalloc(&d_leaf_smid, NLEAF * 4));
  CK(cudaMalloc(&d_cl_smid, NLEAF * 16));
  CK(cudaFuncSetAttribute(hog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HOG_SMEM));
  CK(cudaFuncSetAttribute(leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
  CK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
  const int hog_ctas = 148 * 16;

  auto pcuGetErrorString = entry<CUresult (*)(CUresult, const char**)>("cuGetErrorString");
  auto pDeviceGetDevResource = entry<CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType)>("cuDeviceGetDevResource");
  auto pSplit = entry<CUresult (*)(CUdevResource*, unsigned*, const CUdevResource*, CUdevResource*, unsigned, unsigned)>("cuDevSmResourceSplitByCount");
  auto pGenDesc = entry<CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned)>("cuDevResourceGenerateDesc");
  auto pGreenCreate = entry<CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned)>("cuGreenCtxCreate");
  auto pGreenStream = entry<CUresult (*)(CUstream*, CUgreenCtx, unsigned, int)>("cuGreenCtxStreamCreate");
  if (!pcuGetErrorString || !pDeviceGetDevResource || !pSplit || !pGenDesc || !pGreenCreate || !pGreenStream) return 3;

  cudaStream_t mainS;
  CK(cudaStreamCreateWithFlags(&mainS, cudaStreamNonBlocking));
  cudaEvent_t fork, j1, j2, t0, t1, tc;
  CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&j1, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&j2, cudaEventDisableTiming));
  CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1)); CK(cudaEventCreate(&tc));

  // (a) priority streams, primary context
  int lo, hi;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  cudaStream_t chainP, bulkP;
  CK(cudaStreamCreateWithPriority(&chainP, cudaStreamNonBlocking, hi));
  CK(cudaStreamCreateWithPriority(&bulkP, cudaStreamNonBlocking, lo));
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(t0, mainS));
    enqueue(mainS, chainP, bulkP, fork, j1, j2, hog_ctas);
    CK(cudaEventRecord(t1, mainS));
    CK(cudaStreamSynchronize(mainS));
  }
  float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
  // the chain's own time: run it once more with a timing event on the chain stream
  CK(cudaEventRecord(t0, mainS));
  CK(cudaEventRecord(fork, mainS));
  CK(cudaStreamWaitEvent(chainP, fork, 0)); CK(cudaStreamWaitEvent(bulkP, fork, 0));
  hog_kernel<<<hog_ctas, 128, HOG_SMEM, bulkP>>>(100000, d_hog_smid);
  for (int i = 0; i < NLEAF; i++) {
    leaf_kernel<<<1, 128, LEAF_SMEM, chainP>>>(40000, d_leaf_smid, i);
    cluster_kernel<<<4, 128, LEAF_SMEM, chainP>>>(20000, d_cl_smid, i);
  }
  CK(cudaEventRecord(tc, chainP));
  CK(cudaDeviceSynchronize());
  float msc; CK(cudaEventElapsedTime(&msc, t0, tc));
  report("(a) priority streams, primary context, eager:", msc, ms);

  // chain alone (no hog)
  CK(cudaEventRecord(t0, chainP));
  for (int i = 0; i < NLEAF; i++) {
    leaf_kernel<<<1, 128, LEAF_SMEM, chainP>>>(40000, d_leaf_smid, i);
    cluster_kernel<<<4, 128, LEAF_SMEM, chainP>>>(20000, d_cl_smid, i);
  }
  CK(cudaEventRecord(tc, chainP));
  CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(&msc, t0, tc));
  report("(0) chain alone, eager:", msc, msc);

  // (b) green contexts
  CUdevice dev = 0;
  CUdevResource all, grp[2], rem;
  CU(pDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SM resource: %u SMs\n", all.sm.smCount);
  unsigned ng = 1;
  CU(pSplit(grp, &ng, &all, &rem, 0, 8));
  printf("split: %u group(s) of %u SMs, remainder %u SMs\n", ng, grp[0].sm.smCount, rem.sm.smCount);
  CUdevResourceDesc descA, descB;
  CU(pGenDesc(&descA, &grp[0], 1));
  CU(pGenDesc(&descB, &rem, 1));
  CUgreenCtx gA, gB;
  CU(pGreenCreate(&gA, descA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CU(pGreenCreate(&gB, descB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream chainG, bulkG;
  CU(pGreenStream(&chainG, gA, CU_STREAM_NON_BLOCKING, hi));
  CU(pGreenStream(&bulkG, gB, CU_STREAM_NON_BLOCKING, lo));
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(t0, mainS));
    enqueue(mainS, chainG, bulkG, fork, j1, j2, hog_ctas);
    CK(cudaEventRecord(t1, mainS));
    CK(cudaStreamSynchronize(mainS));
  }
  CK(cudaEventElapsedTime(&ms, t0, t1));
  CK(cudaEventRecord(t0, mainS));
  CK(cudaEventRecord(fork, mainS));
  CK(cudaStreamWaitEvent(chainG, fork, 0)); CK(cudaStreamWaitEvent(bulkG, fork, 0));
  hog_kernel<<<hog_ctas, 128, HOG_SMEM, bulkG>>>(100000, d_hog_smid);
  for (int i = 0; i < NLEAF; i++) {
    leaf_kernel<<<1, 128, LEAF_SMEM, chainG>>>(40000, d_leaf_smid, i);
    cluster_kernel<<<4, 128, LEAF_SMEM, chainG>>>(20000, d_cl_smid, i);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(tc, chainG));
  CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(&msc, t0, tc));
  report("(b) chain in 8-SM green ctx, hog in the remainder, eager:", msc, ms);

  // (c) the same captured into a graph
  cudaGraph_t graph; cudaGraphExec_t exec;
  cudaError_t e = cudaStreamBeginCapture(mainS, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { printf("begin capture: %s\n", cudaGetErrorString(e)); return 4; }
  enqueue(mainS, chainG, bulkG, fork, j1, j2, hog_ctas);
  e = cudaStreamEndCapture(mainS, &graph);
  if (e != cudaSuccess) { printf("end capture: %s\n", cudaGetErrorString(e)); return 4; }
  e = cudaGraphInstantiate(&exec, graph, 0);
  if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); return 4; }
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaMemsetAsync(d_leaf_smid, 0xff, NLEAF * 4, mainS));
    CK(cudaEventRecord(t0, mainS));
    CK(cudaGraphLaunch(exec, mainS));
    CK(cudaEventRecord(t1, mainS));
    CK(cudaStreamSynchronize(mainS));
  }
  CK(cudaEventElapsedTime(&ms, t0, t1));
  report("(c) green-ctx streams captured into one graph, replay:", ms, ms);

  // (d) graph of the priority-stream version for comparison
  e = cudaStreamBeginCapture(mainS, cudaStreamCaptureModeThreadLocal);
  enqueue(mainS, chainP, bulkP, fork, j1, j2, hog_ctas);
  e = cudaStreamEndCapture(mainS, &graph);
  if (e != cudaSuccess) { printf("end capture (d): %s\n", cudaGetErrorString(e)); return 4; }
  CK(cudaGraphInstantiate(&exec, graph, 0));
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(t0, mainS));
    CK(cudaGraphLaunch(exec, mainS));
    CK(cudaEventRecord(t1, mainS));
    CK(cudaStreamSynchronize(mainS));
  }
  CK(cudaEventElapsedTime(&ms, t0, t1));
  report("(d) priority streams captured into one graph, replay:", ms, ms);
  // hog alone for reference
  CK(cudaEventRecord(t0, mainS));
  hog_kernel<<<hog_ctas, 128, HOG_SMEM, mainS>>>(100000, d_hog_smid);
  CK(cudaEventRecord(t1, mainS));
  CK(cudaStreamSynchronize(mainS));
  CK(cudaEventElapsedTime(&ms, t0, t1));
  printf("hog alone on all SMs: %.3f ms\n", ms);
  CK(cudaEventRecord(t0, mainS));
  CK(cudaEventRecord(fork, mainS)); CK(cudaStreamWaitEvent(bulkG, fork, 0));
  hog_kernel<<<hog_ctas, 128, HOG_SMEM, bulkG>>>(100000, d_hog_smid);
  CK(cudaEventRecord(j2, bulkG)); CK(cudaStreamWaitEvent(mainS, j2, 0));
  CK(cudaEventRecord(t1, mainS));
  CK(cudaStreamSynchronize(mainS));
  CK(cudaEventElapsedTime(&ms, t0, t1));
  printf("hog alone in the remainder green ctx: %.3f ms\n", ms);
  return 0;
}
