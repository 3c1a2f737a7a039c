// Latency-bound leaves of the blocked factorisation.
//
// potf2_inv_kernel factors one LEAF x LEAF (128) diagonal block AND inverts the factor inside one CTA.
// The block lives in shared memory.  Each 32-column panel is factored with one thread per row (the row sits in
// registers, the pivot column is broadcast through shared memory: one barrier per column), the trailing update
// and the assembly of the 128 x 128 inverse from the 32 x 32 diagonal inverses run on the fp64 tensor cores
// (DMMA.8x8x4) straight from shared memory.  The inverse-transposed block U_bb = L_bb^-T is what the blocked
// algorithms multiply with: leaf solves use its 32 x 32 diagonal blocks (with a refinement step), the explicit
// inverse K^-1 starts from the whole block.  There is no serial substitution kernel on the O(N^3) path.
#include "linalg.h"

namespace gegp {

namespace {

constexpr int NB = LEAF;        // 128
constexpr int PLD = NB + 4;     // 132 == 4 (mod 16): every DMMA fragment load below is bank-conflict free
constexpr int LT = 256;         // threads per CTA (8 warps)
constexpr int NW = LT / 32;
constexpr unsigned FULL = 0xffffffffu;

// Shared-memory tile T[NB][PLD]:  L(r,c) = T[r][c] (c <= r) ;  U(i,j) = L^-T (i <= j) = T[i][j+1].
// The one-column shift keeps both diagonals (L_ii and 1/L_ii).

constexpr int SLD = 36;  // == 4 (mod 16): stride of the per-warp 8-row scratch strips

// Factor the 32-column panel at j0 (rows j0 .. kk-1): thread t owns row t in registers (all threads of the CTA
// run the loop so that its barriers are plain __syncthreads).  The panel is processed in micro-blocks of MB columns
// with two barriers each instead of one barrier per column:
//   1. the rows of the diagonal block publish their MB entries of the micro-block columns;
//   2. every thread factors the MB x MB diagonal micro-block itself (registers; the rsqrt chain is the only serial
//      part), solves its own row against it, and the diagonal-block rows publish their solved entries;
//   3. every thread applies the rank-MB update to the remaining columns of its row.
// `cb` is 32 x MB doubles, `xb` 32 x MB doubles, `rsd[j0 + j]` receives 1 / L_jj, `bad` (shared, initialised to
// INT_MAX) the first non-positive pivot (1-based, leaf-local).
constexpr int MB = 8;
constexpr int UBLD = 34;  // row stride of the 32 x 32 U_jj staging block (even: double2 rows)

template <int m0>
__device__ __forceinline__ void panel_micro_block(double (&a)[32], double* cb, double* xb, double* rsd, int* bad,
                                                  int j0, int r, bool diag, bool warp_active) {
  if (diag) {
#pragma unroll
    for (int i = 0; i < MB; i += 2)
      *reinterpret_cast<double2*>(cb + (r - j0) * MB + i) = make_double2(a[m0 + i], a[m0 + i + 1]);
  }
  __syncthreads();
  double x[MB];
#pragma unroll
  for (int i = 0; i < MB; i++) x[i] = 0.0;
  if (warp_active) {
    // factor the micro-block (rows m0 .. m0+MB-1 of cb): lm[i][k] (k < i) and rinv[i]
    double lm[MB][MB], rinv[MB];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
      for (int k = 0; k <= i; k++) lm[i][k] = cb[(m0 + i) * MB + k];
#pragma unroll
    for (int j = 0; j < MB; j++) {
      const double piv = lm[j][j];
      if (!(piv > 0.0) && r == j0 + m0 + j) atomicMin(bad, j0 + m0 + j + 1);   // also catches NaN
      const double rs = rsqrt(piv);
      rinv[j] = rs;
      if (r == j0 + m0 + j) rsd[r] = rs;
#pragma unroll
      for (int i = j + 1; i < MB; i++) lm[i][j] *= rs;
#pragma unroll
      for (int i = j + 1; i < MB; i++)
#pragma unroll
        for (int k = j + 1; k <= i; k++) lm[i][k] -= lm[i][j] * lm[k][j];
    }
    // own row against the micro-block
#pragma unroll
    for (int i = 0; i < MB; i++) {
      double v = a[m0 + i];
#pragma unroll
      for (int k = 0; k < i; k++) v -= x[k] * lm[i][k];
      x[i] = v * rinv[i];
      a[m0 + i] = x[i];
    }
    if (diag && m0 + MB < 32) {
#pragma unroll
      for (int i = 0; i < MB; i += 2)
        *reinterpret_cast<double2*>(xb + (r - j0) * MB + i) = make_double2(x[i], x[i + 1]);
    }
  }
  if (m0 + MB < 32) {
    __syncthreads();
    if (warp_active) {
#pragma unroll
      for (int c = m0 + MB; c < 32; c++) {
        const double2* xc = reinterpret_cast<const double2*>(xb + c * MB);
        double v0 = a[c], v1 = 0.0;
#pragma unroll
        for (int i = 0; i < MB; i += 4) {
          const double2 p = xc[i / 2], q = xc[i / 2 + 1];
          v0 -= x[i] * p.x;
          v1 -= x[i + 1] * p.y;
          v0 -= x[i + 2] * q.x;
          v1 -= x[i + 3] * q.y;
        }
        a[c] = v0 + v1;
      }
    }
  }
}

// Rows: thread t < 128 owns row t of the leaf.  Threads 128..159 own 32 VIRTUAL rows e_i^T (identity): the panel
// solve turns them into e_i^T L_jj^-T, i.e. row i of U_jj = L_jj^-T, so the 32 x 32 diagonal inverse the leaf solves
// need comes out of the same substitution at no extra cost on the critical path.  Finished columns go straight
// from registers to global memory (L into A, U_jj into the diagonal 32-blocks of Dinv).
__device__ __forceinline__ void panel_factor32(double* s, double* cb, double* xb, double* rsd, int* bad, int j0, int kk,
                                               int tid, double* ub) {
  const int r = tid;
  const bool virt = (tid >= NB) && (tid < NB + 32);
  const int vi = tid - NB;                                   // virtual row index
  const bool active = ((r >= j0) && (r < kk)) || virt;
  const bool diag = (r >= j0) && (r < j0 + 32) && (r < kk);
  const bool warp_active = virt || (((tid | 31) >= j0) && ((tid & ~31) < kk));   // some row of this warp takes part
  double a[32];
  const double* row = s + (virt ? 0 : r) * PLD + j0;
#pragma unroll
  for (int c = 0; c < 32; c++) {
    double v = 0.0;
    if (virt) v = (c == vi) ? 1.0 : 0.0;
    else if (active && j0 + c <= r) v = row[c];
    a[c] = v;
  }
  panel_micro_block<0>(a, cb, xb, rsd, bad, j0, r, diag, warp_active);
  panel_micro_block<8>(a, cb, xb, rsd, bad, j0, r, diag, warp_active);
  panel_micro_block<16>(a, cb, xb, rsd, bad, j0, r, diag, warp_active);
  panel_micro_block<24>(a, cb, xb, rsd, bad, j0, r, diag, warp_active);
  static_assert(MB == 8, "four micro-blocks per 32-column panel");
  if (virt) {
    double* urow = ub + vi * UBLD;                           // U_jj(vi, c), c >= vi (zero below)
#pragma unroll
    for (int c = 0; c < 32; c += 2) *reinterpret_cast<double2*>(urow + c) = make_double2(a[c], a[c + 1]);
  } else if (active) {
    double* wrow = s + r * PLD + j0;
#pragma unroll
    for (int c = 0; c < 32; c++)
      if (j0 + c <= r) wrow[c] = a[c];
  }
}

// Coalesced write-out of a finished panel (after the barrier that follows panel_factor32): the 32 columns of L at
// j0 for rows j0..k-1 go to A, the 32 x 32 block U_jj goes to the diagonal 32-block of Dinv.  Fire and forget: the
// stores drain while the trailing update runs.
__device__ __forceinline__ void store_panel(const double* s, const double* ub, double* __restrict__ A, int64_t lda,
                                            double* __restrict__ Dinv, int j0, int k, int tid) {
  const int cpair = (tid & 15) * 2;                           // column pair inside the panel
  for (int r = j0 + (tid >> 4); r < k; r += LT / 16) {
    const int c = j0 + cpair;
    const double* sp = s + r * PLD + c;
    double* gp = A + (int64_t)r * lda + c;
    if (c + 1 <= r) *reinterpret_cast<double2*>(gp) = *reinterpret_cast<const double2*>(sp);
    else if (c <= r) gp[0] = sp[0];
  }
  for (int e = tid; e < 32 * 16; e += LT) {
    const int i = e >> 4, c = (e & 15) * 2;
    const double2 v = *reinterpret_cast<const double2*>(ub + i * UBLD + c);
    double* gp = Dinv + (j0 + i) * NB + j0 + c;
    if (c >= i) *reinterpret_cast<double2*>(gp) = v;
    else if (c + 1 >= i) gp[1] = v.y;
  }
}

// Trailing update of the lower triangle of rows/cols [base, kk) with the 32-wide panel at j0:  C -= X X^T.
// Work unit: one row of 8 x 8 tiles times a group of up to four column tiles (shared A fragments, four independent
// accumulator chains); units are dealt round-robin to the warps.
__device__ __forceinline__ void trailing_update32(double* s, int j0, int base, int kk, int warp, int lane) {
  const int lr = lane >> 2, lk = lane & 3;
  const int nt = (kk - base) >> 3;
  int g = warp;                                   // index of this warp's next unit
  int first = 0;                                  // index of the first unit of row tile rt
  for (int rt = 0; rt < nt; rt++) {
    const int ng = (rt >> 2) + 1;                 // column groups of row tile rt: tiles 0..rt
    for (; g < first + ng; g += NW) {
      const int cg = g - first;
      const int r0 = base + rt * 8;
      double af[8];
#pragma unroll
      for (int kq = 0; kq < 8; kq++) af[kq] = -s[(r0 + lr) * PLD + j0 + 4 * kq + lk];
      double acc[4][2];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int ct = min(cg * 4 + u, rt);
        const double* cp = s + (r0 + lr) * PLD + base + ct * 8 + 2 * lk;
        acc[u][0] = cp[0];
        acc[u][1] = cp[1];
      }
#pragma unroll
      for (int kq = 0; kq < 8; kq++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int ct = min(cg * 4 + u, rt);
          const double b = s[(base + ct * 8 + lr) * PLD + j0 + 4 * kq + lk];
          dmma884(acc[u][0], acc[u][1], af[kq], b);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int ct = cg * 4 + u;
        if (ct > rt) continue;
        double* cp = s + (r0 + lr) * PLD + base + ct * 8 + 2 * lk;
        const int r = r0 + lr, c = base + ct * 8 + 2 * lk;
        if (c <= r) cp[0] = acc[u][0];        // entries right of the diagonal are not part of L: never touched
        if (c + 1 <= r) cp[1] = acc[u][1];
      }
    }
    first += ng;
  }
}

// Inverse assembly for the block pair a = [a0, a0+sza), b = [b0, b0+szb), a before b:
//   U_ab = -U_aa * L_ba^T * U_bb
// phase 1:  P = U_aa * L_ba^T  -> stored where U_ab will live ;  phase 2:  U_ab = -P * U_bb (in place).
__device__ __forceinline__ void pair_phase1(double* s, int a0, int sza, int b0, int szb, int w, int nw,
                                            int lane) {
  const int lr = lane >> 2, lk = lane & 3;
  const int nit = sza >> 3, njg = szb >> 5;   // row tiles of 8, column groups of 4 tiles (szb is a multiple of 32)
  for (int t = w; t < nit * njg; t += nw) {
    const int it = t / njg, jg = t - it * njg;
    const int i = it * 8 + lr;
    double acc[4][2];
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u][0] = acc[u][1] = 0.0;
    for (int kq = 2 * it; kq < (sza >> 2); kq++) {
      const int k = 4 * kq + lk;
      const double a = (k >= i) ? s[(a0 + i) * PLD + a0 + k + 1] : 0.0;       // U_aa(i,k)
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const double b = s[(b0 + (jg * 4 + u) * 8 + lr) * PLD + a0 + k];      // L(b0+j, a0+k)
        dmma884(acc[u][0], acc[u][1], a, b);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      double* o = s + (a0 + i) * PLD + b0 + (jg * 4 + u) * 8 + 2 * lk + 1;
      o[0] = acc[u][0];
      o[1] = acc[u][1];
    }
  }
}

template <int MAXKQ>
__device__ __forceinline__ void pair_phase2(double* s, int a0, int sza, int b0, int szb, int w, int nw,
                                            int lane) {
  const int lr = lane >> 2, lk = lane & 3;
  const int nst = sza >> 3, njt = szb >> 3, nkq = szb >> 2;
  for (int st = w; st < nst; st += nw) {
    const int i = a0 + st * 8 + lr;
    double af[MAXKQ];
#pragma unroll
    for (int kq = 0; kq < MAXKQ; kq++) af[kq] = (kq < nkq) ? s[i * PLD + b0 + 4 * kq + lk + 1] : 0.0;
    double acc[MAXKQ / 2][2];
#pragma unroll
    for (int jt = 0; jt < MAXKQ / 2; jt++) acc[jt][0] = acc[jt][1] = 0.0;
    // k outer, column tiles inner: consecutive DMMAs hit independent accumulators
#pragma unroll
    for (int kq = 0; kq < MAXKQ; kq++) {
      if (kq < nkq) {
        const int k = 4 * kq + lk;
#pragma unroll
        for (int jt = kq / 2; jt < MAXKQ / 2; jt++) {
          if (jt < njt) {
            const int j = jt * 8 + lr;
            const double b = (k <= j) ? s[(b0 + k) * PLD + b0 + j + 1] : 0.0;   // U_bb(k,j)
            dmma884(acc[jt][0], acc[jt][1], af[kq], b);
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int jt = 0; jt < MAXKQ / 2; jt++) {
      if (jt < njt) {
        double* o = s + i * PLD + b0 + jt * 8 + 2 * lk + 1;
        o[0] = -acc[jt][0];
        o[1] = -acc[jt][1];
      }
    }
  }
}

}  // namespace

#ifdef GEGP_LEAF_CLOCKS
__device__ long long g_leaf_clk[32];
#define LEAF_CLK(i) do { if (threadIdx.x == 0) g_leaf_clk[i] = clock64(); } while (0)
#define PREP_CLK(i) do { if (threadIdx.x == 0 && rank == 0) g_leaf_clk[16 + (i)] = clock64(); } while (0)
// wall-clock (globaltimer, ns) begin / end of every chain kernel, in launch order: [kind (0 factor, 1 chain step), t0, t1]
__device__ unsigned long long g_chain_ts[3 * 1024];
__device__ unsigned int g_chain_n;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CHAIN_TS_BEGIN(kind, lead)                                                         \
  unsigned int ts_slot__ = 0xffffffffu;                                                    \
  if ((lead) && threadIdx.x == 0 && blockIdx.z == 0) {                                     \
    ts_slot__ = atomicAdd(&g_chain_n, 1u);                                                 \
    if (ts_slot__ < 1024) { g_chain_ts[3 * ts_slot__] = (kind); g_chain_ts[3 * ts_slot__ + 1] = gtimer(); } \
  }
#define CHAIN_TS_END()                                                                     \
  do { if (ts_slot__ < 1024) g_chain_ts[3 * ts_slot__ + 2] = gtimer(); } while (0)
#else
#define LEAF_CLK(i) do { } while (0)
#define PREP_CLK(i) do { } while (0)
#define CHAIN_TS_BEGIN(kind, lead) do { } while (0)
#define CHAIN_TS_END() do { } while (0)
#endif

// One CTA per problem: A (k x k lower, k <= 128) -> L in place; Dinv (128 x 128, ld 128) <- L^-T (upper
// triangular, explicit zeros elsewhere).  The first non-positive pivot is recorded in info[z]
// (1-based global index row0 + i + 1) if info[z] was 0.
// Stage the lower triangle of a leaf block (k x k, identity padded to kk) into the shared tile with cp.async.
__device__ __forceinline__ void stage_lower(double* s, const double* __restrict__ A, int64_t lda, int k, int kk, int tid,
                                            int nthreads, bool wait = true) {
  // 16-byte chunks, all in flight at once (A is 16-byte aligned, lda even).  The chunk that starts ON the diagonal is
  // an 8-byte copy, so nothing is ever written into the shifted upper part (other copies may fill that concurrently).
  for (int e = tid; e < kk * (NB / 2); e += nthreads) {
    const int r = e >> 6, c = (e & 63) * 2;
    if (c > r) continue;
    if (r < k) {
      if (c == r) cp_async8(s + r * PLD + c, A + (int64_t)r * lda + c);
      else cp_async16(s + r * PLD + c, A + (int64_t)r * lda + c, (c + 1 < k) ? 16 : 8);
    } else {   // identity padding
      s[r * PLD + c] = (r == c) ? 1.0 : 0.0;
      if (c + 1 <= r) s[r * PLD + c + 1] = (r == c + 1) ? 1.0 : 0.0;
    }
  }
  if (wait) {
    cp_async_commit();
    cp_async_wait<0>();
  }
}

__global__ void __launch_bounds__(LT, 1)
potf2_inv_kernel(double* __restrict__ A, int64_t lda, int64_t strideA, int k, int row0, int* __restrict__ info,
                 double* __restrict__ Dinv, int64_t strideD) {
  extern __shared__ __align__(16) double s[];  // NB x PLD tile
  __shared__ __align__(16) double cb[32 * MB];
  __shared__ __align__(16) double xb[32 * MB];
  __shared__ __align__(16) double ub[32 * UBLD];
  __shared__ double rsd[NB];
  __shared__ int bad_sh;
  A += (int64_t)blockIdx.z * strideA;
  Dinv += (int64_t)blockIdx.z * strideD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kk = (k + 31) & ~31;  // padded with an identity block up to a multiple of 32
  if (tid == 0) bad_sh = 0x7fffffff;
  CHAIN_TS_BEGIN(0, true);
  LEAF_CLK(0);
  stage_lower(s, A, lda, k, kk, tid, LT);
  __syncthreads();
  LEAF_CLK(1);
  for (int j0 = 0; j0 < kk; j0 += 32) {
    panel_factor32(s, cb, xb, rsd, &bad_sh, j0, kk, tid, ub);
    __syncthreads();
    LEAF_CLK(2 + (j0 >> 5) * 2);
    store_panel(s, ub, A, lda, Dinv, j0, k, tid);
    if (j0 + 32 < kk) {
      trailing_update32(s, j0, j0 + 32, kk, warp, lane);
      __syncthreads();
    }
    LEAF_CLK(3 + (j0 >> 5) * 2);
  }
  if (tid == 0 && bad_sh <= k) atomicCAS(info + blockIdx.z, 0, row0 + bad_sh);
  LEAF_CLK(10);
  CHAIN_TS_END();
}

// Complete the Dinv blocks of a factor: the leaf kernel leaves only the 32 x 32 diagonal inverses there; this
// kernel (one CTA per 128-block, all blocks in parallel, off the factorisation's critical path) assembles the full
// 128 x 128 inverse-transposed block U_bb = L_bb^-T by block doubling on the tensor cores and writes it back with
// explicit zeros below the diagonal.
__global__ void __launch_bounds__(LT, 1)
dinv_assemble_kernel(const double* __restrict__ L, int64_t ldl, int64_t strideL, double* __restrict__ Dinv,
                     int64_t strideD, int N) {
  extern __shared__ __align__(16) double s[];  // NB x PLD tile
  const int b0 = blockIdx.x * NB;
  const int k = min(NB, N - b0);
  const int kk = (k + 31) & ~31;
  L += (int64_t)blockIdx.z * strideL + (int64_t)b0 * ldl + b0;
  Dinv += (int64_t)blockIdx.z * strideD + (int64_t)blockIdx.x * NB * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  stage_lower(s, L, ldl, k, kk, tid, LT);
  __syncthreads();
  for (int e = tid; e < kk * 32; e += LT) {      // diagonal 32-blocks of U, shifted one column right
    const int i = e >> 5, c = (i & ~31) + (e & 31);
    if (c >= i) cp_async8(s + i * PLD + c + 1, Dinv + i * NB + c);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (kk >= 64) {   // level 1: pairs of 32-blocks -> 64-blocks
    const int npairs = (kk >= 128) ? 2 : 1;
    const int nw = NW / npairs, pr = warp / nw, w = warp - pr * nw;
    pair_phase1(s, pr * 64, 32, pr * 64 + 32, 32, w, nw, lane);
    __syncthreads();
    pair_phase2<8>(s, pr * 64, 32, pr * 64 + 32, 32, w, nw, lane);
    __syncthreads();
  }
  if (kk > 64) {    // level 2: [0,64) with [64,kk)
    pair_phase1(s, 0, 64, 64, kk - 64, warp, NW, lane);
    __syncthreads();
    pair_phase2<16>(s, 0, 64, 64, kk - 64, warp, NW, lane);
    __syncthreads();
  }
  // The upper parts of the diagonal 32-blocks are final since the leaf factorisation and may be being read by solves
  // that run beside this kernel (it is issued while the factorisation goes on): they are not stored again.
  for (int i = warp; i < NB; i += NW) {
    const double* srow = s + i * PLD + 1;
    double* grow = Dinv + i * NB;
#pragma unroll
    for (int j = 2 * lane; j < NB; j += 64) {
      double2 v = make_double2(0.0, 0.0);
      if (i < kk && j < kk) {
        if (j >= i) v.x = srow[j];
        if (j + 1 >= i) v.y = srow[j + 1];
      }
      const bool same32 = i < kk && (i >> 5) == (j >> 5);   // j is even: j and j + 1 lie in the same 32-block
      // (rows of the identity padding behind a ragged last leaf were never stored: they get their zeros below)
      if (same32 && j >= i) continue;                // both entries already final
      if (same32 && j + 1 >= i) grow[j] = v.x;       // (j + 1 == i: the diagonal entry is final, the one left of it is 0)
      else *reinterpret_cast<double2*>(grow + j) = v;
    }
  }
}

int leaf_dinv_assemble(const Ctx& ctx, const double* L, int64_t ldl, int64_t strideL, double* Dinv, int64_t strideD,
                       int N) {
  if (N <= 0) return 0;
  const int smem = NB * PLD * (int)sizeof(double);
  GEGP_SET_SMEM(dinv_assemble_kernel, smem);
  timeline_begin(ctx.stream, "dinvasm", N);
  dinv_assemble_kernel<<<dim3((N + NB - 1) / NB, 1, ctx.batch), LT, smem, ctx.stream>>>(L, ldl, strideL, Dinv, strideD, N);
  timeline_end(ctx.stream);
  GEGP_CHECK_LAUNCH();
  return 0;
}

#ifdef GEGP_LEAF_CLOCKS
extern "C" int gegp_debug_leaf_clocks(long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, g_leaf_clk, sizeof(long long) * 32);
}
// copies the chain time stamps (n records of 3 uint64) and resets the counter; returns n
extern "C" int gegp_debug_chain_ts(unsigned long long* out, int max_records) {
  unsigned int n = 0;
  cudaMemcpyFromSymbol(&n, g_chain_n, sizeof(n));
  if (n > 1024u) n = 1024u;
  if ((int)n > max_records) n = (unsigned)max_records;
  cudaMemcpyFromSymbol(out, g_chain_ts, sizeof(unsigned long long) * 3 * n);
  const unsigned int zero = 0;
  cudaMemcpyToSymbol(g_chain_n, &zero, sizeof(zero));
  return (int)n;
}
#endif

int leaf_potf2_inv(const Ctx& ctx, double* A, int64_t lda, int64_t strideA, int k, int row0, int* info, double* Dinv,
                   int64_t strideD) {
  if (k <= 0) return 0;
  if (k > NB) return -901;
  const int smem = NB * PLD * (int)sizeof(double);
  GEGP_SET_SMEM(potf2_inv_kernel, smem);
  timeline_begin(ctx.stream, "potf2", row0, k);
  potf2_inv_kernel<<<dim3(1, 1, ctx.batch), LT, smem, ctx.stream>>>(A, lda, strideA, k, row0, info, Dinv, strideD);
  timeline_end(ctx.stream);
  GEGP_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Leaf solve  B (r x k) <- B * L^-T  for one factored LEAF block (k <= 128), fused in one kernel.
// Block substitution over the four 32-column blocks j:
//     C_j = B_j - sum_{i<j} X_i L_ji^T                      (DMMA, X_i kept in registers as A fragments)
//     X_j = C_j L_jj^-T  via  X1 = C_j U_jj ; R = C_j - X1 L_jj^T ; X_j = X1 + R U_jj
// i.e. only the 32 x 32 diagonal inverses U_jj (the diagonal blocks of Dinv) are multiplied with, and each of
// those products is followed by one refinement step, so the solve is as accurate as a substitution while
// running entirely on the fp64 tensor cores.  A warp owns an 8-row strip; a CTA of 16 warps owns 128 rows and
// stages the L block and the U_jj blocks in the same shifted shared-memory tile the factor kernel uses.
// ------------------------------------------------------------------------------------------------
// TW: warps per CTA of the leaf solve (each warp owns an 8-row strip): 16 for tall panels, 8 when that still gives
// every CTA an SM of its own (half the tensor-pipe time per CTA on the latency-critical mid-size problems).

__device__ __forceinline__ void frag_c_to_a(double* scr, const double (&c)[4][2], double (&a)[8], int lr,
                                            int lk, double sign) {
  __syncwarp();
#pragma unroll
  for (int ct = 0; ct < 4; ct++) {
    scr[lr * SLD + ct * 8 + 2 * lk] = c[ct][0];
    scr[lr * SLD + ct * 8 + 2 * lk + 1] = c[ct][1];
  }
  __syncwarp();
#pragma unroll
  for (int kq = 0; kq < 8; kq++) a[kq] = sign * scr[lr * SLD + 4 * kq + lk];
}

// One 8-row strip of the leaf solve.  `s` is the shared tile with L (lower) and the shifted 32 x 32 diagonal blocks of
// U = L^-T, `scr` this warp's 8 x SLD scratch strip.  On entry x[j][ct] holds the strip's right-hand side as C fragments
// (32-column block j, 8-column tile ct); on exit it holds the solution.  Blocks j >= nb32 are left untouched.
__device__ __forceinline__ void solve_strip(const double* s, double* scr, double (&x)[4][4][2], int nb32, int lr, int lk) {
  double xneg[3][8];  // -X_i as A fragments, i = 0..2
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (j < nb32) {
      const int j0 = j * 32;
      double (&acc)[4][2] = x[j];
      // C_j = B_j - sum_{i<j} X_i L_ji^T
#pragma unroll
      for (int i = 0; i < 3; i++) {
        if (i < j) {
#pragma unroll
          for (int ct = 0; ct < 4; ct++) {
            const double* lrow = s + (j0 + ct * 8 + lr) * PLD + i * 32 + lk;
#pragma unroll
            for (int kq = 0; kq < 8; kq++) dmma884(acc[ct][0], acc[ct][1], xneg[i][kq], lrow[4 * kq]);
          }
        }
      }
      // X1 = C_j U_jj
      double af[8], x1[4][2];
      frag_c_to_a(scr, acc, af, lr, lk, 1.0);
#pragma unroll
      for (int ct = 0; ct < 4; ct++) {
        x1[ct][0] = x1[ct][1] = 0.0;
        const int c = ct * 8 + lr;
#pragma unroll
        for (int kq = 0; kq <= 2 * ct + 1; kq++) {
          const int kx = 4 * kq + lk;
          const double b = (kx <= c) ? s[(j0 + kx) * PLD + j0 + c + 1] : 0.0;
          dmma884(x1[ct][0], x1[ct][1], af[kq], b);
        }
      }
      // R = C_j - X1 L_jj^T
      frag_c_to_a(scr, x1, af, lr, lk, -1.0);
#pragma unroll
      for (int ct = 0; ct < 4; ct++) {
        const int nn = ct * 8 + lr;
#pragma unroll
        for (int kq = 0; kq <= 2 * ct + 1; kq++) {
          const int kx = 4 * kq + lk;
          const double b = (kx <= nn) ? s[(j0 + nn) * PLD + j0 + kx] : 0.0;
          dmma884(acc[ct][0], acc[ct][1], af[kq], b);
        }
      }
      // X_j = X1 + R U_jj
      frag_c_to_a(scr, acc, af, lr, lk, 1.0);
#pragma unroll
      for (int ct = 0; ct < 4; ct++) {
        const int c = ct * 8 + lr;
#pragma unroll
        for (int kq = 0; kq <= 2 * ct + 1; kq++) {
          const int kx = 4 * kq + lk;
          const double b = (kx <= c) ? s[(j0 + kx) * PLD + j0 + c + 1] : 0.0;
          dmma884(x1[ct][0], x1[ct][1], af[kq], b);
        }
      }
#pragma unroll
      for (int ct = 0; ct < 4; ct++) { acc[ct][0] = x1[ct][0]; acc[ct][1] = x1[ct][1]; }
      if (j < 3) frag_c_to_a(scr, x1, xneg[j], lr, lk, -1.0);
    }
  }
}

// One 8-row strip of the leaf solve, streamed: each 32-column block of the strip is read from / written back to global
// memory (row pointer brow) as it is reached, so only one block of the strip is ever held in registers.  Performs
// exactly the arithmetic of solve_strip (same DMMA sequence): the two give bit-identical results.
__device__ __forceinline__ void solve_strip_streamed(const double* s, double* scr, double* brow, bool rok, int k, int nb32,
                                                     bool vec, int lr, int lk) {
  double xneg[3][8];  // -X_i as A fragments, i = 0..2
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (j < nb32) {
      const int j0 = j * 32;
      double acc[4][2];
#pragma unroll
      for (int ct = 0; ct < 4; ct++) {
        const int c = j0 + ct * 8 + 2 * lk;
        acc[ct][0] = acc[ct][1] = 0.0;
        if (rok) {
          if (vec) {
            if (c < k) {
              const double2 v = *reinterpret_cast<const double2*>(brow + c);
              acc[ct][0] = v.x;
              acc[ct][1] = v.y;
            }
          } else {
            if (c < k) acc[ct][0] = brow[c];
            if (c + 1 < k) acc[ct][1] = brow[c + 1];
          }
        }
      }
      // C_j = B_j - sum_{i<j} X_i L_ji^T
#pragma unroll
      for (int i = 0; i < 3; i++) {
        if (i < j) {
#pragma unroll
          for (int ct = 0; ct < 4; ct++) {
            const double* lrow = s + (j0 + ct * 8 + lr) * PLD + i * 32 + lk;
#pragma unroll
            for (int kq = 0; kq < 8; kq++) dmma884(acc[ct][0], acc[ct][1], xneg[i][kq], lrow[4 * kq]);
          }
        }
      }
      // X1 = C_j U_jj
      double af[8], x1[4][2];
      frag_c_to_a(scr, acc, af, lr, lk, 1.0);
#pragma unroll
      for (int ct = 0; ct < 4; ct++) {
        x1[ct][0] = x1[ct][1] = 0.0;
        const int c = ct * 8 + lr;
#pragma unroll
        for (int kq = 0; kq <= 2 * ct + 1; kq++) {
          const int kx = 4 * kq + lk;
          const double b = (kx <= c) ? s[(j0 + kx) * PLD + j0 + c + 1] : 0.0;
          dmma884(x1[ct][0], x1[ct][1], af[kq], b);
        }
      }
      // R = C_j - X1 L_jj^T
      frag_c_to_a(scr, x1, af, lr, lk, -1.0);
#pragma unroll
      for (int ct = 0; ct < 4; ct++) {
        const int nn = ct * 8 + lr;
#pragma unroll
        for (int kq = 0; kq <= 2 * ct + 1; kq++) {
          const int kx = 4 * kq + lk;
          const double b = (kx <= nn) ? s[(j0 + nn) * PLD + j0 + kx] : 0.0;
          dmma884(acc[ct][0], acc[ct][1], af[kq], b);
        }
      }
      // X_j = X1 + R U_jj
      frag_c_to_a(scr, acc, af, lr, lk, 1.0);
#pragma unroll
      for (int ct = 0; ct < 4; ct++) {
        const int c = ct * 8 + lr;
#pragma unroll
        for (int kq = 0; kq <= 2 * ct + 1; kq++) {
          const int kx = 4 * kq + lk;
          const double b = (kx <= c) ? s[(j0 + kx) * PLD + j0 + c + 1] : 0.0;
          dmma884(x1[ct][0], x1[ct][1], af[kq], b);
        }
      }
      if (rok) {
#pragma unroll
        for (int ct = 0; ct < 4; ct++) {
          const int c = j0 + ct * 8 + 2 * lk;
          if (vec) {
            if (c < k) *reinterpret_cast<double2*>(brow + c) = make_double2(x1[ct][0], x1[ct][1]);
          } else {
            if (c < k) brow[c] = x1[ct][0];
            if (c + 1 < k) brow[c + 1] = x1[ct][1];
          }
        }
      }
      if (j < 3) frag_c_to_a(scr, x1, xneg[j], lr, lk, -1.0);
    }
  }
}

// Strip rows <-> global memory (C-fragment layout): all loads of a strip are issued up front so that only one L2
// round trip is exposed per strip.
__device__ __forceinline__ void load_strip(const double* brow, bool rok, int k, bool vec, int lk, double (&x)[4][4][2]) {
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int ct = 0; ct < 4; ct++) {
      const int c = j * 32 + ct * 8 + 2 * lk;
      x[j][ct][0] = x[j][ct][1] = 0.0;
      if (rok) {
        if (vec) {
          if (c < k) {
            const double2 v = *reinterpret_cast<const double2*>(brow + c);
            x[j][ct][0] = v.x;
            x[j][ct][1] = v.y;
          }
        } else {
          if (c < k) x[j][ct][0] = brow[c];
          if (c + 1 < k) x[j][ct][1] = brow[c + 1];
        }
      }
    }
}
__device__ __forceinline__ void store_strip(double* brow, bool rok, int k, bool vec, int lk, const double (&x)[4][4][2]) {
  if (!rok) return;
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int ct = 0; ct < 4; ct++) {
      const int c = j * 32 + ct * 8 + 2 * lk;
      if (vec) {
        if (c < k) *reinterpret_cast<double2*>(brow + c) = make_double2(x[j][ct][0], x[j][ct][1]);
      } else {
        if (c < k) brow[c] = x[j][ct][0];
        if (c + 1 < k) brow[c + 1] = x[j][ct][1];
      }
    }
}

// Stage a factored leaf for solves: L (lower, identity padded to kk) and the 32 x 32 diagonal blocks of U = L^-T
// (shifted one column right) into the shared tile.
__device__ __forceinline__ void stage_factor(double* s, const double* __restrict__ L, int64_t ldl,
                                             const double* __restrict__ Dinv, int k, int kk, int tid, int nthreads) {
  stage_lower(s, L, ldl, k, kk, tid, nthreads, false);   // both sets of copies in flight together: one wait
  for (int e = tid; e < kk * 32; e += nthreads) {   // 8-byte copies: odd offsets
    const int i = e >> 5, c = (i & ~31) + (e & 31);
    if (c >= i) cp_async8(s + i * PLD + c + 1, Dinv + i * NB + c);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
}

template <int TW>
__global__ void __launch_bounds__(TW * 32, 1)
leaf_trsm_kernel(const double* __restrict__ L, int64_t ldl, int64_t strideL, const double* __restrict__ Dinv,
                 int64_t strideD, double* __restrict__ B, int64_t ldb, int64_t strideB, int r, int k) {
  extern __shared__ __align__(16) double s[];  // NB x PLD tile, then TW x 8 x SLD scratch strips
  L += (int64_t)blockIdx.z * strideL;
  Dinv += (int64_t)blockIdx.z * strideD;
  B += (int64_t)blockIdx.z * strideB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  const int kk = (k + 31) & ~31;
  stage_factor(s, L, ldl, Dinv, k, kk, tid, TW * 32);
  double* scr = s + NB * PLD + warp * 8 * SLD;
  const int nb32 = kk >> 5;
  const bool vec = ((k & 1) == 0);
  for (int st = blockIdx.x * TW + warp; st * 8 < r; st += gridDim.x * TW) {
    const int row = st * 8 + lr;
    const bool rok = row < r;
    double* brow = B + (int64_t)row * ldb;
    solve_strip_streamed(s, scr, brow, rok, k, nb32, vec, lr, lk);
  }
}

// ------------------------------------------------------------------------------------------------
// Chain step, first half (everything between two leaf factorisations that the NEXT leaf factorisation waits for):
//   X  = B L_p^-T          the 128 rows right below the previous leaf p (the next leaf's block row), solved in place
//   E -= X X^T             the next leaf's diagonal block (lower triangle), K = 128
// One thread-block cluster of CS CTAs: the 16 row strips of the solve are dealt out over the CTAs, every CTA then
// receives all of X in its shared tile (distributed shared memory stores), and the 8 x 8 output tiles of the update
// are dealt out over all CS x 8 warps.  The rows further down are solved off the critical path by leaf_trsm_kernel.
// Same arithmetic per strip / per output element for every CS, so results do not depend on the cluster size.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f64x2(const void* local_ptr, uint32_t rank, double a, double b) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local_ptr);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(ra), "d"(a), "d"(b) : "memory");
}

constexpr int SYRK_UNITS = 72;   // (row tile rt, pair of column tiles cp <= rt / 2) units of the 128 x 128 lower update

template <int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(LT, 1)
chain_prep_kernel(const double* __restrict__ Lp, int64_t lda, int64_t strideA, const double* __restrict__ Dinv,
                  int64_t strideD, double* __restrict__ B, double* __restrict__ E, int kc) {
  extern __shared__ __align__(16) double s[];  // NB x PLD tile, then NW x 8 x SLD scratch strips
  Lp += (int64_t)blockIdx.z * strideA;
  B += (int64_t)blockIdx.z * strideA;
  E += (int64_t)blockIdx.z * strideA;
  Dinv += (int64_t)blockIdx.z * strideD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  const uint32_t rank = (CS > 1) ? cluster_rank() : 0u;
  constexpr int SPC = 16 / CS;                 // strips per CTA
  constexpr int WTOT = CS * NW;                // warps of the whole cluster
  double* scr = s + NB * PLD + warp * 8 * SLD;
  CHAIN_TS_BEGIN(1, rank == 0);

  if (CS == 1) {
    // ---- one CTA per problem (batches that fill the machine): streamed solve, then X comes back from L2
    stage_factor(s, Lp, lda, Dinv, NB, NB, tid, LT);
    for (int st = warp; st * 8 < kc; st += NW) {
      const int row = st * 8 + lr;
      solve_strip_streamed(s, scr, B + (int64_t)row * lda, row < kc, NB, 4, true, lr, lk);
    }
    __syncthreads();
    for (int e = tid; e < NB * (NB / 2); e += LT) {
      const int r = e >> 6, c = (e & 63) * 2;
      cp_async16(s + r * PLD + c, B + (int64_t)(r < kc ? r : 0) * lda + c, r < kc ? 16 : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  } else {
    // ---- solve: strip rank * SPC + warp of the block row (one per warp), kept in registers
    static_assert(CS == 1 || SPC <= NW, "one strip per warp");
    const int st = (int)rank * SPC + warp;
    const bool mine = (warp < SPC) && st * 8 < kc;            // warp-uniform
    const int row = st * 8 + lr;
    double x[4][4][2];
    PREP_CLK(0);
    load_strip(B + (int64_t)row * lda, mine && row < kc, NB, true, lk, x);
    stage_factor(s, Lp, lda, Dinv, NB, NB, tid, LT);
    PREP_CLK(1);
    if (mine) {
      solve_strip(s, scr, x, 4, lr, lk);
      store_strip(B + (int64_t)row * lda, row < kc, NB, true, lk, x);
    }
    PREP_CLK(2);
    // ---- every CTA gets all of X (rows beyond kc are zero) in its tile, plain layout s[r][c]
    cluster_sync();                                            // everybody is done reading L_p
    PREP_CLK(3);
    if (warp < SPC) {
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int ct = 0; ct < 4; ct++) {
          const double* dst = s + row * PLD + j * 32 + ct * 8 + 2 * lk;
#pragma unroll
          for (int t = 0; t < CS; t++) st_cluster_f64x2(dst, (uint32_t)t, x[j][ct][0], x[j][ct][1]);
        }
    }
    PREP_CLK(4);
    cluster_sync();
    PREP_CLK(5);
  }
  // ---- E -= X X^T (accumulate X X^T from zero, then subtract: the arithmetic of the GEMM engine with alpha = -1,
  //      beta = 1).  Work unit: a row tile times a PAIR of column tiles; the 72 units of the lower triangle are dealt
  //      round-robin over the CTAs first, then over the warps, so that every SM of the cluster carries the same load.
#pragma unroll 1
  for (int u = (int)rank + CS * warp; u < SYRK_UNITS; u += WTOT) {
    // unit u -> (rt, cp): row tiles 2g and 2g+1 have g + 1 column pairs each
    int g = 0, first = 0;
    while (u >= first + 2 * (g + 1)) { first += 2 * (g + 1); g++; }
    const int rt = 2 * g + (u - first) / (g + 1), cp = (u - first) % (g + 1);
    if (rt * 8 >= kc) continue;                // warp-uniform
    const int r = rt * 8 + lr;
    double e[2][2], acc[2][2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const int ct = cp * 2 + t, c = ct * 8 + 2 * lk;
      e[t][0] = e[t][1] = acc[t][0] = acc[t][1] = 0.0;
      if (ct <= rt && r < kc) {
        if (c <= r) e[t][0] = E[(int64_t)r * lda + c];
        if (c + 1 <= r) e[t][1] = E[(int64_t)r * lda + c + 1];
      }
    }
    const double* arow = s + r * PLD + lk;
    const double* b0 = s + (min(cp * 2, rt) * 8 + lr) * PLD + lk;
    const double* b1 = s + (min(cp * 2 + 1, rt) * 8 + lr) * PLD + lk;
#pragma unroll 16
    for (int kq = 0; kq < 32; kq++) {
      const double a = arow[4 * kq];
      dmma884(acc[0][0], acc[0][1], a, b0[4 * kq]);
      dmma884(acc[1][0], acc[1][1], a, b1[4 * kq]);
    }
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const int ct = cp * 2 + t, c = ct * 8 + 2 * lk;
      if (ct <= rt && r < kc) {
        if (c <= r) E[(int64_t)r * lda + c] = e[t][0] - acc[t][0];
        if (c + 1 <= r) E[(int64_t)r * lda + c + 1] = e[t][1] - acc[t][1];
      }
    }
  }
  PREP_CLK(6);
  PREP_CLK(7);
  CHAIN_TS_END();
}

template <int TW>
static int launch_leaf_trsm(const Ctx& ctx, const double* L, int64_t ldl, int64_t strideL, const double* Dinv,
                            int64_t strideD, double* B, int64_t ldb, int64_t strideB, int r, int k) {
  const int smem = (NB * PLD + TW * 8 * SLD) * (int)sizeof(double);
  GEGP_SET_SMEM(leaf_trsm_kernel<TW>, smem);
  const int nctas = (r + TW * 8 - 1) / (TW * 8);
  timeline_begin(ctx.stream, "ltrsm", r, k);
  leaf_trsm_kernel<TW><<<dim3(nctas, 1, ctx.batch), TW * 32, smem, ctx.stream>>>(L, ldl, strideL, Dinv, strideD, B, ldb,
                                                                                strideB, r, k);
  timeline_end(ctx.stream);
  GEGP_CHECK_LAUNCH();
  return 0;
}

template <int CS>
static int launch_chain_prep(const Ctx& ctx, const double* Lp, int64_t lda, int64_t strideA, const double* Dinv,
                             int64_t strideD, double* B, double* E, int kc) {
  const int smem = (NB * PLD + NW * 8 * SLD) * (int)sizeof(double);
  GEGP_SET_SMEM(chain_prep_kernel<CS>, smem);
  timeline_begin(ctx.stream, "cprep", kc, CS);
  chain_prep_kernel<CS><<<dim3(CS, 1, ctx.batch), LT, smem, ctx.stream>>>(Lp, lda, strideA, Dinv, strideD, B, E, kc);
  timeline_end(ctx.stream);
  GEGP_CHECK_LAUNCH();
  return 0;
}

static int g_chain_cluster = -1;
int& chain_cluster() {
  if (g_chain_cluster < 0) g_chain_cluster = getenv("GEGP_CHAIN_CLUSTER") ? atoi(getenv("GEGP_CHAIN_CLUSTER")) : 0;
  return g_chain_cluster;
}

int leaf_chain_prep(const Ctx& ctx, const double* Lp, int64_t lda, int64_t strideA, const double* Dinv, int64_t strideD,
                    double* B, double* E, int kc) {
  if (kc <= 0) return 0;
  if (kc > NB) return -905;
  // a lone problem is latency-bound: spread the chain step over a cluster of four SMs; a batch that fills the machine
  // with one CTA per problem runs the same arithmetic in a single CTA each (bit-identical either way)
  int cs = chain_cluster();
  if (cs != 1 && cs != 2 && cs != 4) cs = (ctx.batch * 4 <= 148) ? 4 : (ctx.batch * 2 <= 148 ? 2 : 1);
  if (cs == 4) return launch_chain_prep<4>(ctx, Lp, lda, strideA, Dinv, strideD, B, E, kc);
  if (cs == 2) return launch_chain_prep<2>(ctx, Lp, lda, strideA, Dinv, strideD, B, E, kc);
  return launch_chain_prep<1>(ctx, Lp, lda, strideA, Dinv, strideD, B, E, kc);
}

int leaf_trsm(const Ctx& ctx, const double* L, int64_t ldl, int64_t strideL, const double* Dinv, int64_t strideD,
              double* B, int64_t ldb, int64_t strideB, int r, int k) {
  if (r <= 0 || k <= 0) return 0;
  if (k > NB) return -902;
  // (both variants do the same arithmetic per 8-row strip: results are bit-identical whichever is chosen)
  if (ctx.batch == 1 && r <= 148 * 64) return launch_leaf_trsm<8>(ctx, L, ldl, strideL, Dinv, strideD, B, ldb, strideB, r, k);
  return launch_leaf_trsm<16>(ctx, L, ldl, strideL, Dinv, strideD, B, ldb, strideB, r, k);
}

// ------------------------------------------------------------------------------------------------
// U diagonal blocks <- Dinv blocks and their transposes (start of the explicit inverse; the rest of U is produced
// by GEMMs).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_dinv_kernel(const double* __restrict__ Dinv, int64_t strideD, double* __restrict__ U, int64_t ldu,
                    int64_t strideU, double* __restrict__ W, int64_t ldw, int64_t strideW, int N) {
  Dinv += (int64_t)blockIdx.z * strideD + (int64_t)blockIdx.x * NB * NB;
  U += (int64_t)blockIdx.z * strideU;
  W += (int64_t)blockIdx.z * strideW;
  const int b0 = blockIdx.x * NB;
  const int k = min(NB, N - b0);
  for (int e = threadIdx.x; e < NB * NB; e += 256) {
    const int i = e >> 7, j = e & (NB - 1);
    if (i < k && j < k) {
      U[(int64_t)(b0 + i) * ldu + b0 + j] = Dinv[e];             // L_bb^-T: upper triangular, zeros below
      W[(int64_t)(b0 + i) * ldw + b0 + j] = Dinv[j * NB + i];    // L_bb^-1: lower triangular, zeros above
    }
  }
}

int leaf_scatter_dinv(const Ctx& ctx, const double* Dinv, int64_t strideD, double* U, int64_t ldu, int64_t strideU,
                      double* W, int64_t ldw, int64_t strideW, int N) {
  if (N <= 0) return 0;
  scatter_dinv_kernel<<<dim3((N + NB - 1) / NB, 1, ctx.batch), 256, 0, ctx.stream>>>(Dinv, strideD, U, ldu, strideU, W, ldw,
                                                                                  strideW, N);
  GEGP_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Back substitution x <- L^-T x for nrhs (<= 4) vectors, blocked by LEAF, diagonal blocks applied as
// x_j <- U_jj x_j.  Step j (descending): every CTA c < j applies x_c -= L[j,c]^T x_j ; CTA c == j-1 then
// multiplies its block with U so that x_{j-1} is final for the next step.  One launch per block row.
// (Only used when alpha is wanted without the gradient; the gradient path multiplies with L^-T directly.)
// ------------------------------------------------------------------------------------------------
constexpr int VMAXRHS = 4;

__global__ void __launch_bounds__(256)
trsv_lt_step_kernel(const double* __restrict__ L, int64_t ldl, int64_t strideL, const double* __restrict__ Dinv,
                    int64_t strideD, double* __restrict__ x, int64_t ldx, int64_t strideX, int N, int nrhs, int j) {
  __shared__ double xj[VMAXRHS * NB];
  __shared__ double xs[VMAXRHS * NB];
  L += (int64_t)blockIdx.z * strideL;
  x += (int64_t)blockIdx.z * strideX;
  Dinv += (int64_t)blockIdx.z * strideD;
  const int nblk = (N + NB - 1) / NB;
  const int tid = threadIdx.x;
  const int c = blockIdx.x;  // column block updated by this CTA
  const int i = tid & (NB - 1), half = tid >> 7;  // 2 halves split the reduction range
  if (j < nblk) {
    const int j0 = j * NB, kj = min(NB, N - j0), c0 = c * NB;
    for (int e = tid; e < nrhs * NB; e += 256) {
      const int rr = e / NB, ii = e % NB;
      xj[e] = (ii < kj) ? x[(int64_t)rr * ldx + j0 + ii] : 0.0;
    }
    __syncthreads();
    // x_c[i] -= sum_r L[j0+r][c0+i] * x_j[r]   (c0 block is always full: c < j)
    double acc[VMAXRHS] = {0, 0, 0, 0};
    for (int r = half; r < kj; r += 2) {
      const double l = L[(int64_t)(j0 + r) * ldl + c0 + i];
#pragma unroll
      for (int rr = 0; rr < VMAXRHS; rr++)
        if (rr < nrhs) acc[rr] += l * xj[rr * NB + r];
    }
    __syncthreads();
    if (half == 1)
#pragma unroll
      for (int rr = 0; rr < VMAXRHS; rr++)
        if (rr < nrhs) xs[rr * NB + i] = acc[rr];
    __syncthreads();
    if (half == 0)
#pragma unroll
      for (int rr = 0; rr < VMAXRHS; rr++)
        if (rr < nrhs) x[(int64_t)rr * ldx + c0 + i] -= acc[rr] + xs[rr * NB + i];
    __syncthreads();
  }
  if (c == j - 1) {
    // x_c <- U_cc x_c ,  U_cc[i][r] (r >= i) at Dinv[c][i*NB + r]
    const int b0 = c * NB, k = min(NB, N - b0);
    const double* D = Dinv + (int64_t)c * NB * NB;
    for (int e = tid; e < nrhs * NB; e += 256) {
      const int rr = e / NB, ii = e % NB;
      xj[e] = (ii < k) ? x[(int64_t)rr * ldx + b0 + ii] : 0.0;
    }
    __syncthreads();
    double acc[VMAXRHS] = {0, 0, 0, 0};
    // thread (i, half) sums r = i + half, i + half + 2, ... ; rows of D are read with stride NB (L2-resident, tiny)
    for (int r = i + half; r < k; r += 2) {
      const double u = D[i * NB + r];
#pragma unroll
      for (int rr = 0; rr < VMAXRHS; rr++)
        if (rr < nrhs) acc[rr] += u * xj[rr * NB + r];
    }
    if (half == 1)
#pragma unroll
      for (int rr = 0; rr < VMAXRHS; rr++)
        if (rr < nrhs) xs[rr * NB + i] = acc[rr];
    __syncthreads();
    if (half == 0 && i < k)
#pragma unroll
      for (int rr = 0; rr < VMAXRHS; rr++)
        if (rr < nrhs) x[(int64_t)rr * ldx + b0 + i] = acc[rr] + xs[rr * NB + i];
  }
}

int trsv_lower_trans(const Ctx& ctx, const double* L, int64_t ldl, int64_t strideL, const double* Dinv, int64_t strideD,
                     double* x, int64_t ldx, int64_t strideX, int N, int nrhs) {
  if (N <= 0 || nrhs <= 0) return 0;
  if (nrhs > VMAXRHS) return -903;
  const int nblk = (N + NB - 1) / NB;
  // step j = nblk: only the diagonal multiply of the last block; then j = nblk-1 .. 1
  for (int j = nblk; j >= 1; j--) {
    trsv_lt_step_kernel<<<dim3(j, 1, ctx.batch), 256, 0, ctx.stream>>>(L, ldl, strideL, Dinv, strideD, x, ldx, strideX,
                                                                      N, nrhs, j);
    GEGP_CHECK_LAUNCH();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// y = U x for the upper-triangular U = L^-T (N x N, row-major): one warp per row, x staged through L2.
// Gives alpha = L^-T (L^-1 r) on the gradient path without any substitution.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
trmv_upper_kernel(const double* __restrict__ U, int64_t ldu, int64_t strideU, const double* __restrict__ x,
                  int64_t strideX, double* __restrict__ y, int64_t strideY, int N) {
  U += (int64_t)blockIdx.z * strideU;
  x += (int64_t)blockIdx.z * strideX;
  y += (int64_t)blockIdx.z * strideY;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const double* u = U + (int64_t)row * ldu;
  double a0 = 0.0, a1 = 0.0;
  int j = (row & ~1) + 2 * lane;  // even start: 16-byte aligned pairs
  for (; j + 1 < N; j += 64) {
    const double2 uv = *reinterpret_cast<const double2*>(u + j);
    const double2 xv = *reinterpret_cast<const double2*>(x + j);
    if (j >= row) a0 += uv.x * xv.x;
    a1 += uv.y * xv.y;
  }
  if (j < N && j >= row) a0 += u[j] * x[j];
  const double t = warp_sum(a0 + a1);
  if (lane == 0) y[row] = t;
}

int trmv_upper(const Ctx& ctx, const double* U, int64_t ldu, int64_t strideU, const double* x, int64_t strideX,
               double* y, int64_t strideY, int N) {
  if (N <= 0) return 0;
  timeline_begin(ctx.stream, "trmv", N);
  trmv_upper_kernel<<<dim3((N + 7) / 8, 1, ctx.batch), 256, 0, ctx.stream>>>(U, ldu, strideU, x, strideX, y, strideY, N);
  timeline_end(ctx.stream);
  GEGP_CHECK_LAUNCH();
  return 0;
}

}  // namespace gegp
