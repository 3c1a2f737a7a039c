"""Generate tests/golden/*.npz by running the REFERENCE (marchildon/gpgradpy at /root/reference).

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (the GPU box has no /root/reference);
the fixtures it writes are committed and are what pins oracle/gegp_oracle.py and the CUDA path.

    python oracle/make_golden.py            # writes tests/golden/*.npz and golden_report.json

The reference imports `smt` at module scope (optz/GpHparaX0.py:12); oracle/ref_shim provides a stand-in.
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from gpgradpy.src.GaussianProcess import GaussianProcess  # noqa: E402  (the reference)
from oracle import gegp_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
report = {}


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(1e-300, np.abs(b))))


def make_gp(x, f, g, mode, mask=None, std_f=0.0, std_g=0.0):
    n, d = x.shape
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.set_data(x, f, std_f * np.ones(n), g, std_g * np.ones(g.shape), mask)
    return GP


def case_noise_free(name, n, d, mode, seed=0, lo=-2.0, hi=2.0, theta=None, mask=None, store_mats=True,
                    nx=0, calc_grad=True):
    x, f, g = O.synthetic_problem(n, d, seed, lo, hi)
    if mask is not None:
        g_in = g[mask]
    else:
        g_in = g
    th = O.bench_theta(d) if theta is None else np.asarray(theta, float)
    GP = make_gp(x, f, g_in, mode, mask)
    hp = GP.make_hp_class(theta=th)
    t0 = time.time()
    info, ok = GP.calc_lkd_all(hp, calc_grad=calc_grad)
    t_lkd = time.time() - t0
    xs, Rt = GP.get_scl_x_w_dist()
    fs, _, gs, _ = GP.get_scl_eval_data()
    d_out = dict(x=x, fval=f, grad=g_in, theta=th, mode=mode, eta=GP._etaK, ok=ok,
                 x_scl=xs, fval_scl=fs, grad_scl=gs,
                 mask=np.zeros(0, bool) if mask is None else mask)
    if ok:
        d_out.update(ln_lkd=info.ln_lkd, hp_varK=info.hp_varK, hp_beta=info.hp_beta, ln_det=info.ln_det_Kmat)
        if calc_grad:
            d_out.update(ln_lkd_grad=info.ln_lkd_grad)
    if store_mats:
        Kern, Kcor, Kcov, fac, _, eta, _ = GP.calc_all_K_w_chofac(Rt, hp, varK=1)
        d_out.update(Kern=Kern, Kcov=Kcov)
        if Kcor is not None:
            d_out.update(Kcor=Kcor)
        if fac is not None:
            L = np.tril(fac[0]) if fac[1] else np.triu(fac[0]).T
            d_out.update(chol_lower=L)
    else:
        # sampled rows of the kernel matrix keep the fixture small
        Kern = GP.calc_Kern(Rt, th, None, GP.bvec_use_grad, GP.bvec_use_grad)
        rows = np.unique(np.linspace(0, Kern.shape[0] - 1, 12).astype(int))
        d_out.update(Kern_rows_idx=rows, Kern_rows=Kern[rows], Kern_fro=np.linalg.norm(Kern))
    if nx and ok:
        rng = np.random.default_rng(100 + seed)
        xt = rng.uniform(lo, hi, (nx, d))
        xt[: min(3, n)] = x[: min(3, n)] + 1e-3      # near training points: sigma ~ 0
        hp2 = GP.make_hp_class(theta=th, varK=info.hp_varK, beta=info.hp_beta)
        GP.set_hpara("set", 1, hp2)
        mu, sig = GP.eval_model(xt)[:2]
        d_out.update(x_test=xt, mu=mu, sig=sig)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d_out)
    # oracle-vs-reference deviation (recorded, not asserted here)
    if ok and mask is None:
        o = O.lkd_wo_noise(xs, fs, gs, th, mode, GP._etaK, calc_grad=calc_grad)
        r = dict(t_ref_s=t_lkd, lml=rel(o.ln_lkd, info.ln_lkd), varK=rel(o.hp_varK, info.hp_varK))
        if calc_grad:
            r["grad"] = rel(o.ln_lkd_grad, info.ln_lkd_grad)
        report[name] = r
    print(name, "ok" if ok else "CHOL FAIL", f"{t_lkd:.2f}s", report.get(name, ""), flush=True)


def case_noisy(name, n, d, mode, std_f, std_g, varK, seed=0):
    x, f, g = O.synthetic_problem(n, d, seed)
    th = O.bench_theta(d)
    GP = make_gp(x, f, g, mode, None, std_f, std_g)
    hp = GP.make_hp_class(theta=th, varK=varK)
    info, ok = GP.calc_lkd_all(hp, calc_grad=True)
    xs, Rt = GP.get_scl_x_w_dist()
    noise = GP.calc_noise_vec(hp)
    Kern, Kcor, Kcov, fac, _, eta, _ = GP.calc_all_K_w_chofac(Rt, hp)
    d_out = dict(x=x, fval=f, grad=g, theta=th, mode=mode, eta=GP._etaK, ok=ok, varK=varK, noise_vec=noise,
                 std_f=std_f, std_g=std_g, Kern=Kern, Kcov=Kcov,
                 ln_lkd=info.ln_lkd, ln_lkd_grad=info.ln_lkd_grad, hp_beta=info.hp_beta, ln_det=info.ln_det_Kmat)
    if Kcor is not None:
        d_out.update(Kcor=Kcor)
    rng = np.random.default_rng(7)
    xt = rng.uniform(-2, 2, (16, d))
    hp2 = GP.make_hp_class(theta=th, varK=varK, beta=info.hp_beta)
    GP.set_hpara("set", 1, hp2)
    mu, sig = GP.eval_model(xt)[:2]
    d_out.update(x_test=xt, mu=mu, sig=sig)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d_out)
    o = O.lkd_w_noise(x, f, g, th, varK, noise, mode, GP._etaK)
    report[name] = dict(lml=rel(o.ln_lkd, info.ln_lkd), grad=rel(o.ln_lkd_grad, info.ln_lkd_grad))
    print(name, report[name], flush=True)


def case_candidates(name, n, d, B, seed=0):
    """Batch point A (optz/GpHparaX0.py:39-45): LML at B candidate rows, sequential reference loop."""
    x, f, g = O.synthetic_problem(n, d, seed)
    GP = make_gp(x, f, g, "precon")
    rng = np.random.default_rng(seed)
    log_th = rng.uniform(-5.0, 1.0, (B, d))
    lml = np.full(B, np.nan)
    grad = np.full((B, d), np.nan)
    varK = np.full(B, np.nan)
    for i in range(B):
        hp = GP.hp_vec2dataclass(GP.hp_info_optz_lkd, log_th[i])
        info, ok = GP.calc_lkd_all(hp, calc_grad=True)
        if ok:
            lml[i], grad[i], varK[i] = info.ln_lkd, info.ln_lkd_grad, info.hp_varK
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, fval=f, grad=g, log10_theta=log_th,
                        ln_lkd=lml, ln_lkd_grad=grad, hp_varK=varK, eta=GP._etaK)
    print(name, "n_ok", int(np.sum(np.isfinite(lml))), flush=True)


if __name__ == "__main__":
    # config 1: 2-D Rosenbrock, n=20, points in [0.9, 1.1]^2 (plt/plt_cond.py:108-118), three modes
    for mode in ("precon", "base", "rescale_origin"):
        case_noise_free(f"c1_d2_n20_{mode}", 20, 2, mode, lo=0.9, hi=1.1, theta=[2.0, 8.0], nx=24)
    case_noise_free("d2_n20_wide_precon", 20, 2, "precon", nx=24)
    case_noise_free("d4_n37_precon", 37, 4, "precon", nx=24)
    case_noise_free("d3_n30_base", 30, 3, "base", nx=24)
    case_noise_free("d3_n25_rescale_origin", 25, 3, "rescale_origin", nx=24)
    case_noise_free("d5_n64_precon_seed1", 64, 5, "precon", seed=1, nx=24, store_mats=False)
    case_noise_free("d1_n9_precon", 9, 1, "precon", nx=8)
    # partial gradients: prefix mask (reference dK/dtheta is right) and a scattered mask (builder only)
    m = np.zeros(18, bool); m[:11] = True
    case_noise_free("d3_n18_mask_prefix", 18, 3, "precon", mask=m)
    m2 = np.array([1, 0, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0], bool)
    case_noise_free("d2_n12_mask_scatter", 12, 2, "precon", mask=m2, calc_grad=False)
    # noisy data (known std): varK is an explicit hyper-parameter
    case_noisy("d3_n24_noisy_precon", 24, 3, "precon", 1e-2, 5e-2, 350.0)
    case_noisy("d3_n24_noisy_base", 24, 3, "base", 1e-2, 5e-2, 350.0)
    # candidate batch (config 4 shape, small B)
    case_candidates("c4_d5_n200_cand8", 200, 5, 8)
    # config 4 single and config 2 (scalars + sampled kernel rows only)
    case_noise_free("c4_d5_n200_precon", 200, 5, "precon", store_mats=False, nx=32)
    case_noise_free("c2_d10_n500_precon", 500, 10, "precon", store_mats=False, nx=32)
    json.dump(report, open(os.path.join(OUT, "golden_report.json"), "w"), indent=1)
