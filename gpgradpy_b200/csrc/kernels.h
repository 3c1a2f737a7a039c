// Internal kernel launchers of the gradient-enhanced GP path (covariance build, LML reductions, predict).
#pragma once
#include "common.cuh"
#include "../../include/gegp.h"

namespace gegp {

// Problem geometry: n points in d dimensions, n_g of them carry gradients.
// Matrix/data order (dimension-major): value of point a -> a ; d/dx_i at gradient slot g -> n + i*ng + g.
struct Geom {
  int n, ng, d, N;
  const double* X;   // [n, d] row-major, device
  const int* slot;   // [n] gradient slot of each point or -1 ; nullptr when every point has a gradient
};

// Observation noise as the kernels see it: noise[z*stride + row], optionally divided by varK (per problem).
struct NoiseSpec {
  const double* noise;     // nullptr: noise-free
  int64_t stride;          // 0: one vector shared by every problem of the batch
  const double* varK_all;  // per-problem varK (device) or nullptr -> use the scalar
  double varK;             // scalar varK when varK_all == nullptr
  int divide;              // 1: use noise / varK (kernel/Kernel.py:218)
};

int launch_prep_p(const Ctx& ctx, const Geom& gm, const double* theta, int64_t strideTheta, NoiseSpec ns, int mode,
                  double* p, double* pinv, int64_t strideP);
int launch_build_cov(const Ctx& ctx, const Geom& gm, const double* theta, int64_t strideTheta, NoiseSpec ns,
                     const double* pinv, int64_t strideP, int mode, double eta, double* out, int64_t ld,
                     int64_t strideOut, int lower_only);
int launch_append_res(const Ctx& ctx, int N, int n, const double* y, double beta, const double* pinv, double* row);
int launch_scale_vec(const Ctx& ctx, int N, const double* a, int64_t strideA, const double* pinv, int64_t strideP,
                     double* out, int64_t strideOut);
int launch_cross_cov(const Ctx& ctx, const Geom& gm, const double* theta, const double* pinv, const double* Xs, int nx,
                     double* out, int64_t ld);
int launch_cross_cov_dx(const Ctx& ctx, const Geom& gm, const double* theta, const double* pinv, const double* Xs, int nx,
                        double* out, int64_t ld);
int launch_predict_grad_rows(const Ctx& ctx, int N, int d, const double* Z, int64_t ldz, int nx, const double* w,
                             double beta, double varK, double* mu, double* sig, double* sig2, double* dmu, double* dsig,
                             int* n_negative);
int launch_predict_hess(const Ctx& ctx, const Geom& gm, const double* theta, const double* xs, const double* a,
                        const double* b, const double* Z, int64_t ldz, double* out);
int launch_append_rhs(const Ctx& ctx, int N, int n, const double* y, const double* pinv, int64_t strideP, double* rows,
                      int64_t ld, int64_t strideRows);

// out layout per problem (GEGP_OUT_* in gegp.h)
int launch_lml_finalize(const Ctx& ctx, int N, const double* A, int64_t lda, int64_t strideA, const double* pinv,
                        int64_t strideP, int noisy, const double* varK, double* w, int64_t strideW, double* out,
                        int64_t strideOut, const int* info, int zero_grad_d);
int launch_lml_grad(const Ctx& ctx, const Geom& gm, const double* theta, int64_t strideTheta, const double* Kinv,
                    int64_t ldk, int64_t strideK, const double* alpha_t, int64_t strideAlpha, const double* pinv,
                    int64_t strideP, int mode, double eta, int noisy, const double* varK, double pnlt_grad, double* partial,
                    int64_t stridePartial, double* out, int64_t strideOut, int quad = 0);
int launch_predict_rows(const Ctx& ctx, int N, const double* Z, int64_t ldz, int nx, const double* w, double beta,
                        double varK, double* mu, double* sig, double* sig2, int* n_negative);

size_t lml_grad_partial_doubles(int n, int d);

}  // namespace gegp
