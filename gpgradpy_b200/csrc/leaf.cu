// Latency-bound leaves of the blocked factorisation: diagonal-block Cholesky, substitution-based
// triangular solves (no explicit inverses on the solve path), diagonal-block triangular inverse,
// and the blocked back-substitution for a few right-hand sides.
#include "linalg.h"

namespace gegp {

// ------------------------------------------------------------------------------------------------
// potf2: k x k (k <= 128) lower Cholesky in shared memory, one CTA per problem.
// ------------------------------------------------------------------------------------------------
constexpr int PNB = LEAF;       // 128
constexpr int PLD = PNB + 1;    // odd stride: column reads hit distinct banks

__global__ void __launch_bounds__(256) potf2_kernel(double* A, int64_t lda, int64_t strideA, int k, int row0,
                                                     int* info) {
  extern __shared__ double s[];  // PNB x PLD
  __shared__ int bad;
  A += (int64_t)blockIdx.z * strideA;
  const int tid = threadIdx.x;
  if (tid == 0) bad = 0;
  for (int e = tid; e < PNB * PNB; e += 256) {
    const int r = e / PNB, c = e % PNB;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < k && c <= r) v = A[(int64_t)r * lda + c];
    s[r * PLD + c] = v;
  }
  __syncthreads();
  const int tr = tid >> 4, tc = tid & 15;
  for (int j = 0; j < k; j++) {
    __syncthreads();  // trailing update of step j-1 is complete
    const double piv = s[j * PLD + j];
    if (!(piv > 0.0)) {  // also catches NaN
      if (tid == 0 && bad == 0) bad = j + 1;
    }
    const double dj = sqrt(piv);
    const double inv = 1.0 / dj;
    __syncthreads();  // everyone has read the pivot
    for (int r = j + 1 + tid; r < k; r += 256) s[r * PLD + j] *= inv;
    if (tid == 0) s[j * PLD + j] = dj;
    __syncthreads();
    // rank-1 update of the trailing lower triangle: s[r][c] -= s[r][j]*s[c][j], j < c <= r < k
    const int p0 = (j + 1 - tr + 15) >> 4, q0 = (j + 1 - tc + 15) >> 4;
    for (int p = max(p0, 0); tr + 16 * p < k; p++) {
      const int r = tr + 16 * p;
      const double lrj = s[r * PLD + j];
      for (int q = max(q0, 0); tc + 16 * q <= r; q++) {
        const int c = tc + 16 * q;
        s[r * PLD + c] -= lrj * s[c * PLD + j];
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < k * k; e += 256) {
    const int r = e / k, c = e % k;
    if (c <= r) A[(int64_t)r * lda + c] = s[r * PLD + c];
  }
  if (tid == 0 && bad) atomicCAS(info + blockIdx.z, 0, row0 + bad);
}

int leaf_potf2(const Ctx& ctx, double* A, int64_t lda, int64_t strideA, int k, int row0, int* info) {
  if (k <= 0) return 0;
  if (k > PNB) return -901;
  static bool attr = false;
  const int smem = PNB * PLD * (int)sizeof(double);
  if (!attr) { GEGP_SET_SMEM(potf2_kernel, smem); attr = true; }
  potf2_kernel<<<dim3(1, 1, ctx.batch), 256, smem, ctx.stream>>>(A, lda, strideA, k, row0, info);
  GEGP_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// trsm (right, lower, transposed): B <- B * L^-T, k <= 64, by forward substitution per row.
// One thread per row; the row lives in shared memory, 8 columns at a time in registers.
// ------------------------------------------------------------------------------------------------
constexpr int TKB = 64;         // max triangle width of the leaf
constexpr int TROWS = 128;      // rows per CTA
constexpr int TXLD = TKB + 1;

__global__ void __launch_bounds__(TROWS) trsm_right_leaf_kernel(const double* L, int64_t ldl, int64_t strideL,
                                                                 double* B, int64_t ldb, int64_t strideB, int r,
                                                                 int k) {
  extern __shared__ __align__(16) double tsm[];
  double* Lt = tsm;               // Lt[kk*TKB + c] = L[c][kk]  (zero above the diagonal)
  double* xs = tsm + TKB * TKB;   // TROWS x TXLD
  L += (int64_t)blockIdx.z * strideL;
  B += (int64_t)blockIdx.z * strideB;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * TROWS;
  for (int e = tid; e < TKB * TKB; e += TROWS) {
    const int c = e / TKB, kk = e % TKB;  // read L row-major (coalesced), store transposed
    double v = (c == kk) ? 1.0 : 0.0;
    if (c < k && kk <= c) v = L[(int64_t)c * ldl + kk];
    else if (c != kk) v = 0.0;
    Lt[kk * TKB + c] = v;
  }
  const int nrows = min(TROWS, r - row0);
  for (int e = tid; e < TROWS * TKB; e += TROWS) {
    const int rr = e / TKB, c = e % TKB;
    xs[rr * TXLD + c] = (rr < nrows && c < k) ? B[(int64_t)(row0 + rr) * ldb + c] : 0.0;
  }
  __syncthreads();
  double* x = xs + tid * TXLD;
  for (int c0 = 0; c0 < k; c0 += 8) {
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = x[c0 + c];
    for (int kk = 0; kk < c0; kk++) {
      const double xk = x[kk];
      const double2* lp = reinterpret_cast<const double2*>(Lt + kk * TKB + c0);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double2 l2 = lp[c];
        acc[2 * c] -= l2.x * xk;
        acc[2 * c + 1] -= l2.y * xk;
      }
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
      double v = acc[c];
#pragma unroll
      for (int kk = 0; kk < c; kk++) v -= Lt[(c0 + kk) * TKB + c0 + c] * acc[kk];
      v /= Lt[(c0 + c) * TKB + c0 + c];
      acc[c] = v;
      x[c0 + c] = v;
    }
  }
  __syncthreads();
  for (int e = tid; e < TROWS * TKB; e += TROWS) {
    const int rr = e / TKB, c = e % TKB;
    if (rr < nrows && c < k) B[(int64_t)(row0 + rr) * ldb + c] = xs[rr * TXLD + c];
  }
}

int leaf_trsm_right(const Ctx& ctx, const double* L, int64_t ldl, int64_t strideL, double* B, int64_t ldb,
                    int64_t strideB, int r, int k) {
  if (r <= 0 || k <= 0) return 0;
  if (k > TKB) return -902;
  static bool attr = false;
  const int smem = (TKB * TKB + TROWS * TXLD) * (int)sizeof(double);
  if (!attr) { GEGP_SET_SMEM(trsm_right_leaf_kernel, smem); attr = true; }
  trsm_right_leaf_kernel<<<dim3((r + TROWS - 1) / TROWS, 1, ctx.batch), TROWS, smem, ctx.stream>>>(L, ldl, strideL, B,
                                                                                             ldb, strideB, r, k);
  GEGP_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// trtri (transposed output): U_blk = (L_blk^-1)^T for every LEAF x LEAF diagonal block.
// One CTA per block, one thread per column of L_blk^-1 (forward substitution on e_j).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PNB) trtri_t_kernel(const double* L, int64_t ldl, int64_t strideL, double* U,
                                                       int64_t ldu, int64_t strideU, int N) {
  extern __shared__ double sm[];
  double* Ls = sm;               // PNB x PLD, row-major lower block
  double* Ms = sm + PNB * PLD;   // packed lower: Ms[r*(r+1)/2 + j] = (L^-1)[r][j], j <= r
  L += (int64_t)blockIdx.z * strideL;
  U += (int64_t)blockIdx.z * strideU;
  const int b0 = blockIdx.x * PNB;
  const int k = min(PNB, N - b0);
  const int tid = threadIdx.x;
  for (int e = tid; e < PNB * PNB; e += PNB) {
    const int r = e / PNB, c = e % PNB;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < k && c <= r) v = L[(int64_t)(b0 + r) * ldl + b0 + c];
    Ls[r * PLD + c] = v;
  }
  __syncthreads();
  const int j = tid;  // column of the inverse
  for (int r = 0; r < k; r++) {
    double acc = (r == j) ? 1.0 : 0.0;
    if (r > j) {
      double a0 = 0, a1 = 0;
      int kk = j;
      for (; kk + 1 < r; kk += 2) {
        a0 += Ls[r * PLD + kk] * Ms[kk * (kk + 1) / 2 + j];
        a1 += Ls[r * PLD + kk + 1] * Ms[(kk + 1) * (kk + 2) / 2 + j];
      }
      if (kk < r) a0 += Ls[r * PLD + kk] * Ms[kk * (kk + 1) / 2 + j];
      acc -= a0 + a1;
    }
    if (r >= j) Ms[r * (r + 1) / 2 + j] = acc / Ls[r * PLD + r];
  }
  __syncthreads();
  // U[b0+j][b0+r] = Minv[r][j], r >= j  (upper triangular); coalesced along r
  for (int e = tid; e < k * k; e += PNB) {
    const int jj = e / k, r = e % k;
    if (r >= jj) U[(int64_t)(b0 + jj) * ldu + b0 + r] = Ms[r * (r + 1) / 2 + jj];
  }
}

int leaf_trtri_t(const Ctx& ctx, const double* L, int64_t ldl, int64_t strideL, double* U, int64_t ldu,
                 int64_t strideU, int N) {
  if (N <= 0) return 0;
  static bool attr = false;
  const int smem = (PNB * PLD + PNB * (PNB + 1) / 2) * (int)sizeof(double);
  if (!attr) { GEGP_SET_SMEM(trtri_t_kernel, smem); attr = true; }
  trtri_t_kernel<<<dim3((N + PNB - 1) / PNB, 1, ctx.batch), PNB, smem, ctx.stream>>>(L, ldl, strideL, U, ldu, strideU, N);
  GEGP_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Back substitution x <- L^-T x for nrhs (<= 4) vectors, blocked by LEAF.
// Step j (descending): every CTA c < j applies x_c -= L[j,c]^T x_j ; CTA c == j-1 then solves its
// diagonal block so that x_{j-1} is final for the next step.  One launch per block row.
// ------------------------------------------------------------------------------------------------
constexpr int VNB = LEAF;
constexpr int VMAXRHS = 4;

__device__ void trsv_diag_solve_t(const double* L, int64_t ldl, int b0, int k, double* x, int64_t ldx, int nrhs,
                                  double* xs /*[VMAXRHS][VNB]*/, double* Ls /*[VNB][PLD]*/) {
  // solve L_bb^T z = x_b in place; the block is staged in shared memory, warp `w` handles rhs w
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int e = tid; e < k * k; e += blockDim.x) {
    const int r = e / k, c = e % k;
    if (c <= r) Ls[r * PLD + c] = L[(int64_t)(b0 + r) * ldl + b0 + c];
  }
  for (int e = tid; e < nrhs * VNB; e += blockDim.x) {
    const int rr = e / VNB, i = e % VNB;
    xs[rr * VNB + i] = (i < k) ? x[(int64_t)rr * ldx + b0 + i] : 0.0;
  }
  __syncthreads();
  if (w < nrhs) {
    double* z = xs + w * VNB;
    for (int i = k - 1; i >= 0; i--) {
      // z_i = (z_i - sum_{r>i} L[r][i] z_r) / L[i][i]
      double part = 0.0;
      for (int r = i + 1 + lane; r < k; r += 32) part += Ls[r * PLD + i] * z[r];
      part = warp_sum(part);
      if (lane == 0) z[i] = (z[i] - part) / Ls[i * PLD + i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < nrhs * VNB; e += blockDim.x) {
    const int rr = e / VNB, i = e % VNB;
    if (i < k) x[(int64_t)rr * ldx + b0 + i] = xs[rr * VNB + i];
  }
}

__global__ void __launch_bounds__(256) trsv_lt_step_kernel(const double* L, int64_t ldl, int64_t strideL, double* x,
                                                            int64_t ldx, int64_t strideX, int N, int nrhs, int j) {
  extern __shared__ double Ls_dyn[];  // VNB x PLD, used by the CTA that solves a diagonal block
  __shared__ double xj[VMAXRHS * VNB];
  __shared__ double xs[VMAXRHS * VNB];
  L += (int64_t)blockIdx.z * strideL;
  x += (int64_t)blockIdx.z * strideX;
  const int nblk = (N + VNB - 1) / VNB;
  const int tid = threadIdx.x;
  const int c = blockIdx.x;  // column block updated by this CTA
  if (j < nblk) {
    const int j0 = j * VNB, kj = min(VNB, N - j0), c0 = c * VNB;
    for (int e = tid; e < nrhs * VNB; e += 256) {
      const int rr = e / VNB, i = e % VNB;
      xj[e] = (i < kj) ? x[(int64_t)rr * ldx + j0 + i] : 0.0;
    }
    __syncthreads();
    // x_c[i] -= sum_r L[j0+r][c0+i] * x_j[r]   (c0 block is always full: c < j)
    const int i = tid & (VNB - 1), half = tid >> 7;  // 2 halves split the r range
    double acc[VMAXRHS] = {0, 0, 0, 0};
    for (int r = half; r < kj; r += 2) {
      const double l = L[(int64_t)(j0 + r) * ldl + c0 + i];
#pragma unroll
      for (int rr = 0; rr < VMAXRHS; rr++)
        if (rr < nrhs) acc[rr] += l * xj[rr * VNB + r];
    }
    __syncthreads();
    if (half == 1)
      for (int rr = 0; rr < nrhs; rr++) xs[rr * VNB + i] = acc[rr];
    __syncthreads();
    if (half == 0)
      for (int rr = 0; rr < nrhs; rr++) x[(int64_t)rr * ldx + c0 + i] -= acc[rr] + xs[rr * VNB + i];
    __threadfence_block();
    __syncthreads();
  }
  if (c == j - 1) {
    const int b0 = c * VNB;
    trsv_diag_solve_t(L, ldl, b0, min(VNB, N - b0), x, ldx, nrhs, xs, Ls_dyn);
  }
}

int trsv_lower_trans(const Ctx& ctx, const double* L, int64_t ldl, int64_t strideL, double* x, int64_t ldx,
                     int64_t strideX, int N, int nrhs) {
  if (N <= 0 || nrhs <= 0) return 0;
  if (nrhs > VMAXRHS) return -903;
  const int nblk = (N + VNB - 1) / VNB;
  static bool attr = false;
  const int smem = VNB * PLD * (int)sizeof(double);
  if (!attr) { GEGP_SET_SMEM(trsv_lt_step_kernel, smem); attr = true; }
  // step j = nblk: only the diagonal solve of the last block; then j = nblk-1 .. 1
  for (int j = nblk; j >= 1; j--) {
    trsv_lt_step_kernel<<<dim3(j, 1, ctx.batch), 256, smem, ctx.stream>>>(L, ldl, strideL, x, ldx, strideX, N, nrhs, j);
    GEGP_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace gegp
