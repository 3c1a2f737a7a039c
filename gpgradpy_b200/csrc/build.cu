// K1: fused Gaussian-kernel covariance builder for the gradient-enhanced GP.
//
// Writes the N x N matrix (N = n + n_g*d, dimension-major order) in ONE pass: value block, dK/dx
// blocks, d2K/dxdx' blocks, observation noise, diagonal preconditioner P^-1 . P^-1 and nugget.
// Replaces sq_exp_calc_KernGrad (kernel/KernelSqExp.py:322-410), calc_Rtensor (base/CommonFun.py:58-84)
// and the dense diag-matrix products of calc_all_K_w_chofac (kernel/Kernel.py:213-237, 268-277).
//
// Each CTA owns a tile of point pairs (TA x TB); the squared distance and exp are evaluated once per
// pair, then all (d+1)^2 block entries of the pair are emitted.  Threads run along b, so every store
// instruction of a warp writes one contiguous 256-byte row segment.  HBM-write bound: 8*N^2 bytes.
#include "kernels.h"

namespace gegp {

constexpr int TA = 16;   // a-points per CTA
constexpr int TB = 64;   // b-points per CTA (2 warps wide)
constexpr int BUILD_THREADS = 256;

// p[row] = sqrt(diag(K) + noise[row]), pinv = 1/p.  diag(K) = 1 (value rows), 2*theta_i (gradient rows).
__device__ __forceinline__ double noise_at(const NoiseSpec& ns, int z, int row) {
  const double v = ns.noise[z * ns.stride + row];
  if (!ns.divide) return v;
  return v / (ns.varK_all ? ns.varK_all[z] : ns.varK);
}

__global__ void prep_p_kernel(Geom gm, const double* __restrict__ theta, int64_t strideTheta, NoiseSpec ns, int mode,
                              double* __restrict__ p, double* __restrict__ pinv, int64_t strideP) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= gm.N) return;
  const int z = blockIdx.z;
  double pv = 1.0, pi = 1.0;
  if (mode == GEGP_MODE_PRECON) {
    const double* th = theta + z * strideTheta;
    double dg = 1.0;
    if (row >= gm.n) dg = 2.0 * th[(row - gm.n) / gm.ng];
    if (ns.noise) dg += noise_at(ns, z, row);
    pv = sqrt(dg);
    pi = 1.0 / pv;
  }
  if (p) p[z * strideP + row] = pv;
  pinv[z * strideP + row] = pi;
}

// mode: GEGP_MODE_BASE   out = varK * (K + diag(noise) + eta I)
//       GEGP_MODE_PRECON out = varK * (P^-1 (K + diag(noise)) P^-1 + eta I)
//       GEGP_MODE_PRECON_COV out = varK * (K + diag(noise) + eta * diag(K + noise))   (= P Ktilde P)
__global__ void __launch_bounds__(BUILD_THREADS)
build_cov_kernel(Geom gm, const double* __restrict__ theta_all, int64_t strideTheta, NoiseSpec ns,
                 const double* __restrict__ pinv_all, int64_t strideP, int mode, double eta,
                 double* __restrict__ out_all, int64_t ld, int64_t strideOut, int lower_only) {
  extern __shared__ double sm[];
  const int d = gm.d, n = gm.n, ng = gm.ng;
  double* xa = sm;                 // [TA][d]
  double* xb = xa + TA * d;        // [d][TB]  (transposed: conflict-free along b)
  double* th = xb + d * TB;        // [d]
  double* sg = th + d;             // [d] gradient-row scale 1/sqrt(2 theta_i) (precon, noise-free) or 1
  int* slot_a = reinterpret_cast<int*>(sg + d);  // [TA]
  int* slot_b = slot_a + TA;                     // [TB]

  const int tid = threadIdx.x;
  const int a0 = blockIdx.y * TA, b0 = blockIdx.x * TB;
  const int z = blockIdx.z;
  const double* theta = theta_all + z * strideTheta;
  const double* pinv = pinv_all ? pinv_all + z * strideP : nullptr;
  double* out = out_all + z * strideOut;
  const bool noise = (ns.noise != nullptr);
  const double varK = ns.varK_all ? ns.varK_all[z] : ns.varK;

  for (int e = tid; e < TA * d; e += BUILD_THREADS) {
    const int a = e / d, i = e % d;
    xa[e] = (a0 + a < n) ? gm.X[(int64_t)(a0 + a) * d + i] : 0.0;
  }
  for (int e = tid; e < TB * d; e += BUILD_THREADS) {
    const int b = e / d, i = e % d;
    xb[i * TB + b] = (b0 + b < n) ? gm.X[(int64_t)(b0 + b) * d + i] : 0.0;
  }
  const bool precon = (mode == GEGP_MODE_PRECON);
  for (int e = tid; e < d; e += BUILD_THREADS) {
    th[e] = theta[e];
    sg[e] = precon ? 1.0 / sqrt(2.0 * theta[e]) : 1.0;
  }
  for (int e = tid; e < TA; e += BUILD_THREADS) slot_a[e] = (a0 + e < n) ? (gm.slot ? gm.slot[a0 + e] : a0 + e) : -1;
  for (int e = tid; e < TB; e += BUILD_THREADS) slot_b[e] = (b0 + e < n) ? (gm.slot ? gm.slot[b0 + e] : b0 + e) : -1;
  __syncthreads();

  const int tb = tid & (TB - 1);        // b within tile
  const int ta0 = tid / TB;             // 0..3 ; thread owns a = ta0 + 4*q, q = 0..3
  const int b = b0 + tb;
  if (b >= n) return;
  const int sb = slot_b[tb];
  const bool use_pvec = precon && noise;  // per-row scales only differ from sg[] with noise

#pragma unroll 1
  for (int q = 0; q < TA / 4; q++) {
    const int al = ta0 + 4 * q, a = a0 + al;
    if (a >= n) continue;
    const int sa = slot_a[al];
    const double* xav = xa + al * d;
    double e = 0.0;
    for (int i = 0; i < d; i++) {
      const double r = xav[i] - xb[i * TB + tb];
      e -= th[i] * (r * r);
    }
    const double k = exp(e);
    const bool same = (a == b);
    // ---- value-value entry
    {
      const double sr = use_pvec ? pinv[a] : 1.0, sc = use_pvec ? pinv[b] : 1.0;
      double v = k;
      if (same && noise) v += noise_at(ns, z, a);
      double dadd = 0.0;
      if (same) dadd = (mode == GEGP_MODE_PRECON_COV) ? eta * v : eta;
      v = (v * sr) * sc + dadd;
      if (!lower_only || a >= b) out[(int64_t)a * ld + b] = varK * v;
    }
    // ---- value row, gradient columns: K_0j = +2 th_j r_j k   (upper part: skipped when lower_only)
    if (sb >= 0 && !lower_only) {
      const double sr = use_pvec ? pinv[a] : 1.0;
      for (int j = 0; j < d; j++) {
        const int col = n + j * ng + sb;
        const double r = xav[j] - xb[j * TB + tb];
        const double sc = use_pvec ? pinv[col] : sg[j];
        out[(int64_t)a * ld + col] = varK * (((2.0 * th[j] * r * k) * sr) * sc);
      }
    }
    if (sa < 0) continue;
    // ---- gradient rows
    for (int i = 0; i < d; i++) {
      const int row = n + i * ng + sa;
      const double ri = xav[i] - xb[i * TB + tb];
      const double sr = use_pvec ? pinv[row] : sg[i];
      const double ui = th[i] * ri;
      double* orow = out + (int64_t)row * ld;
      // gradient-value: K_i0 = -2 th_i r_i k
      {
        const double sc = use_pvec ? pinv[b] : 1.0;
        orow[b] = varK * (((-2.0 * ui * k) * sr) * sc);
      }
      if (sb < 0) continue;
      const int jmax = lower_only ? i : d - 1;
      for (int j = 0; j <= jmax; j++) {
        if (lower_only && j == i && a < b) continue;
        const int col = n + j * ng + sb;
        const double rj = xav[j] - xb[j * TB + tb];
        double v;
        if (j == i) v = (2.0 * th[i] - 4.0 * (ui * ui)) * k;   // (2 th_i - 4 th_i^2 r_i^2) k
        else v = -4.0 * (ui * (th[j] * rj)) * k;              // -4 th_i th_j r_i r_j k
        const double sc = use_pvec ? pinv[col] : sg[j];
        double dadd = 0.0;
        if (same && j == i) {
          if (noise) v += noise_at(ns, z, row);
          dadd = (mode == GEGP_MODE_PRECON_COV) ? eta * v : eta;
        }
        orow[col] = varK * ((v * sr) * sc + dadd);
      }
    }
  }
}

// Cross covariance rows for prediction: Kx[x][col] = K(x*_x ; training datum col) * pinv[col]
// value columns: k ; gradient column (j, slot): -2 th_j r_j k with r = x_train - x_test
// (eval/GpEvalModel.py:133-139 builds K(X, X*) and keeps its value columns; this is its transpose).
__global__ void __launch_bounds__(256)
cross_cov_kernel(Geom gm, const double* __restrict__ theta, const double* __restrict__ pinv,
                 const double* __restrict__ Xs, int nx, double* __restrict__ out, int64_t ld) {
  extern __shared__ double sm[];
  const int d = gm.d, n = gm.n, ng = gm.ng;
  double* xt = sm;            // [CX][d] test points of this CTA
  double* th = xt + 8 * d;    // [d]
  constexpr int CX = 8;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.y * CX;
  for (int e = tid; e < CX * d; e += 256) {
    const int x = e / d, i = e % d;
    xt[e] = (x0 + x < nx) ? Xs[(int64_t)(x0 + x) * d + i] : 0.0;
  }
  for (int e = tid; e < d; e += 256) th[e] = theta[e];
  __syncthreads();
  const int b = blockIdx.x * 256 + tid;  // training point
  if (b >= n) return;
  const int sb = gm.slot ? gm.slot[b] : b;
  const double* xb = gm.X + (int64_t)b * d;
  for (int x = 0; x < CX && x0 + x < nx; x++) {
    const double* xv = xt + x * d;
    double e = 0.0;
    for (int i = 0; i < d; i++) {
      const double r = xb[i] - xv[i];
      e -= th[i] * (r * r);
    }
    const double k = exp(e);
    double* orow = out + (int64_t)(x0 + x) * ld;
    orow[b] = k * (pinv ? pinv[b] : 1.0);
    if (sb >= 0) {
      for (int j = 0; j < d; j++) {
        const int col = n + j * ng + sb;
        const double r = xb[j] - xv[j];
        orow[col] = (-2.0 * th[j] * r * k) * (pinv ? pinv[col] : 1.0);
      }
    }
  }
}

// rows[0] = pinv .* y ; rows[1] = pinv .* H  (H = [1_n ; 0], eval/GpMeanFun.py:172-191)
__global__ void append_rhs_kernel(int N, int n, const double* __restrict__ y, const double* __restrict__ pinv,
                                  int64_t strideP, double* __restrict__ rows, int64_t ld, int64_t strideRows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double s = pinv ? pinv[blockIdx.z * strideP + i] : 1.0;
  double* r = rows + blockIdx.z * strideRows;
  r[i] = y[i] * s;
  r[ld + i] = (i < n) ? s : 0.0;
}

// row = pinv .* (y - beta H)
__global__ void append_res_kernel(int N, int n, const double* __restrict__ y, double beta,
                                  const double* __restrict__ pinv, double* __restrict__ row) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  row[i] = (y[i] - (i < n ? beta : 0.0)) * (pinv ? pinv[i] : 1.0);
}

int launch_append_res(const Ctx& ctx, int N, int n, const double* y, double beta, const double* pinv, double* row) {
  append_res_kernel<<<(N + 255) / 256, 256, 0, ctx.stream>>>(N, n, y, beta, pinv, row);
  GEGP_CHECK_LAUNCH();
  return 0;
}

int launch_prep_p(const Ctx& ctx, const Geom& gm, const double* theta, int64_t strideTheta, NoiseSpec ns, int mode,
                  double* p, double* pinv, int64_t strideP) {
  prep_p_kernel<<<dim3((gm.N + 255) / 256, 1, ctx.batch), 256, 0, ctx.stream>>>(gm, theta, strideTheta, ns, mode, p,
                                                                             pinv, strideP);
  GEGP_CHECK_LAUNCH();
  return 0;
}

int launch_build_cov(const Ctx& ctx, const Geom& gm, const double* theta, int64_t strideTheta, NoiseSpec ns,
                     const double* pinv, int64_t strideP, int mode, double eta, double* out, int64_t ld,
                     int64_t strideOut, int lower_only) {
  const size_t smem = (size_t)(TA * gm.d + gm.d * TB + 2 * gm.d) * sizeof(double) + (TA + TB) * sizeof(int);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    GEGP_SET_SMEM(build_cov_kernel, smem);
    smem_set = smem;
  }
  dim3 grid((gm.n + TB - 1) / TB, (gm.n + TA - 1) / TA, ctx.batch);
  build_cov_kernel<<<grid, BUILD_THREADS, smem, ctx.stream>>>(gm, theta, strideTheta, ns, pinv, strideP, mode, eta, out,
                                                              ld, strideOut, lower_only);
  GEGP_CHECK_LAUNCH();
  return 0;
}

int launch_cross_cov(const Ctx& ctx, const Geom& gm, const double* theta, const double* pinv, const double* Xs, int nx,
                     double* out, int64_t ld) {
  const size_t smem = (size_t)(8 * gm.d + gm.d) * sizeof(double);
  dim3 grid((gm.n + 255) / 256, (nx + 7) / 8, 1);
  cross_cov_kernel<<<grid, 256, smem, ctx.stream>>>(gm, theta, pinv, Xs, nx, out, ld);
  GEGP_CHECK_LAUNCH();
  return 0;
}

int launch_append_rhs(const Ctx& ctx, int N, int n, const double* y, const double* pinv, int64_t strideP, double* rows,
                      int64_t ld, int64_t strideRows) {
  append_rhs_kernel<<<dim3((N + 255) / 256, 1, ctx.batch), 256, 0, ctx.stream>>>(N, n, y, pinv, strideP, rows, ld,
                                                                              strideRows);
  GEGP_CHECK_LAUNCH();
  return 0;
}

}  // namespace gegp
