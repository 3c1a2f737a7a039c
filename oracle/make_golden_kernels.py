"""Generate tests/golden/kern_*.npz by running the REFERENCE with its Matern-5/2 ('Ma5f2') and rational-quadratic
('RatQu') kernels (kernel/KernelMatern5f2.py, kernel/KernelRatQuad.py) in all three conditioning modes.

TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference).  Each fixture holds the kernel matrices, the
noise-free LML with its gradient w.r.t. [theta.., alpha (RatQu)], the posterior with x-gradients at a few test points
and the surrogate Hessian at one point -- all produced by the reference's own GaussianProcess methods.

    python oracle/make_golden_kernels.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from gpgradpy.src.GaussianProcess import GaussianProcess  # noqa: E402  (the reference)
from oracle import gegp_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
report = {}


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(1e-300, float(np.max(np.abs(b)))))


def case(name, kernel, hp_kernel, n, d, mode, seed=0, mask=None, nx=12, theta=None, post=True, mats=True):
    x, f, g = O.synthetic_problem(n, d, seed)
    g_in = g if mask is None else g[mask]
    th = O.bench_theta(d) * 4.0 if theta is None else np.asarray(theta, float)
    GP = GaussianProcess(d, True, kernel, mode)
    GP.set_data(x, f, np.zeros(n), g_in, np.zeros(g_in.shape), mask)
    hp = GP.make_hp_class(theta=th, kernel=hp_kernel)
    info, ok = GP.calc_lkd_all(hp, calc_grad=True)
    assert ok, name
    xs, Rt = GP.get_scl_x_w_dist()
    fs, _, gs, _ = GP.get_scl_eval_data()
    Kern, Kcor, Kcov, fac, _, eta, _ = GP.calc_all_K_w_chofac(Rt, hp, varK=1)
    out = dict(kernel=kernel, hp_kernel=np.nan if hp_kernel is None else float(hp_kernel), x=x, fval=f, grad=g_in, theta=th,
               mode=mode, eta=GP._etaK, x_scl=xs, fval_scl=fs, grad_scl=gs,
               mask=np.zeros(0, bool) if mask is None else mask, Kern=Kern, Kcov=Kcov,
               ln_lkd=info.ln_lkd, ln_lkd_grad=info.ln_lkd_grad, hp_varK=info.hp_varK, hp_beta=info.hp_beta,
               ln_det=info.ln_det_Kmat, n_hp=GP.hp_info_optz_lkd.n_hp)
    if Kcor is not None:
        out.update(Kcor=Kcor)
    if not mats:      # larger cases: scalars, gradient and posterior only (keeps the fixtures small)
        for key in ("Kern", "Kcov", "Kcor"):
            out.pop(key, None)
    # posterior with x-gradients, and the Hessians at one point (the reference takes one point per call there)
    rng = np.random.default_rng(50 + seed)
    xt = rng.uniform(-2, 2, (nx, d))
    xt[:2] = x[:2] + 1e-3
    if post:
        GP.set_hpara("set", 1, GP.make_hp_class(theta=th, kernel=hp_kernel, varK=info.hp_varK, beta=info.hp_beta))
        mu, sig, dmu, dsig = GP.eval_model(xt, calc_grad=True)[:4]
        out.update(x_test=xt, mu=mu, sig=sig, dmudx=dmu, dsigdx=dsig)
    if mask is None:
        h = GP.eval_model(xt[3:4], calc_grad=True, calc_hess=True)
        out.update(d2mudx2=h[4], d2sigdx2=h[5])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    # oracle vs reference (recorded; asserted in tests/test_oracle_golden.py)
    kern = (kernel, hp_kernel)
    m = "precon" if mode == "precon" else "base"
    o = O.lkd_wo_noise(xs, fs, gs, th, m, GP._etaK, mask, kernel=kern)
    report[name] = dict(lml=rel(o.ln_lkd, info.ln_lkd), grad=rel(o.ln_lkd_grad, info.ln_lkd_grad),
                        Kern=rel(O.kern_grad(xs, xs, th, mask, mask, kernel=kern), Kern), eta=float(GP._etaK))
    print(name, report[name], flush=True)


def case_noisy(name, kernel, hp_kernel, n, d, mode, std_f, std_g, varK, seed=0):
    x, f, g = O.synthetic_problem(n, d, seed)
    th = O.bench_theta(d) * 4.0
    GP = GaussianProcess(d, True, kernel, mode)
    GP.set_data(x, f, std_f * np.ones(n), g, std_g * np.ones(g.shape))
    hp = GP.make_hp_class(theta=th, kernel=hp_kernel, varK=varK)
    info, ok = GP.calc_lkd_all(hp, calc_grad=True)
    assert ok, name
    xs, Rt = GP.get_scl_x_w_dist()
    noise = GP.calc_noise_vec(hp)
    Kern, Kcor, Kcov, fac, _, eta, _ = GP.calc_all_K_w_chofac(Rt, hp)
    out = dict(kernel=kernel, hp_kernel=np.nan if hp_kernel is None else float(hp_kernel), x=x, fval=f, grad=g, theta=th,
               mode=mode, eta=GP._etaK, varK=varK, noise_vec=noise, std_f=std_f, std_g=std_g, Kcov=Kcov,
               ln_lkd=info.ln_lkd, ln_lkd_grad=info.ln_lkd_grad, hp_beta=info.hp_beta, ln_det=info.ln_det_Kmat)
    rng = np.random.default_rng(7)
    xt = rng.uniform(-2, 2, (10, d))
    GP.set_hpara("set", 1, GP.make_hp_class(theta=th, kernel=hp_kernel, varK=varK, beta=info.hp_beta))
    mu, sig = GP.eval_model(xt)[:2]
    out.update(x_test=xt, mu=mu, sig=sig)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    o = O.lkd_w_noise(x, f, g, th, varK, noise, mode, GP._etaK, kernel=(kernel, hp_kernel))
    report[name] = dict(lml=rel(o.ln_lkd, info.ln_lkd), grad=rel(o.ln_lkd_grad, info.ln_lkd_grad))
    print(name, report[name], flush=True)


if __name__ == "__main__":
    for kernel, hp, tag in (("Ma5f2", None, "ma5f2"), ("RatQu", 2.0, "ratqu_a2"), ("RatQu", 0.7, "ratqu_a07")):
        for mode in ("precon", "base", "rescale_origin"):
            case(f"kern_{tag}_d3_n22_{mode}", kernel, hp, 22, 3, mode)
    # partial gradients.  Two reference defects shape these cases: matern_5f2_calc_KernGrad_grad_th is only right for
    # prefix masks (like the Gaussian kernel's), and rat_quad_calc_KernGrad raises for
    # (bvec_use_grad1 = mask, bvec_use_grad2 = None), i.e. eval_model cannot run with a RatQu mask at all.
    mp = np.zeros(14, bool); mp[:9] = True
    case("kern_ma5f2_d2_n14_mask_prefix", "Ma5f2", None, 14, 2, "precon", mask=mp, seed=1)
    ms = np.array([1, 0, 1, 1, 0, 1, 1, 1, 0, 1, 1, 0, 1, 1], bool)
    case("kern_ratqu_d2_n14_mask_scatter", "RatQu", 1.3, 14, 2, "precon", mask=ms, seed=1, post=False)
    case("kern_ma5f2_d5_n60_precon", "Ma5f2", None, 60, 5, "precon", seed=2, mats=False)
    case("kern_ratqu_d5_n60_precon", "RatQu", 3.0, 60, 5, "precon", seed=2, mats=False)
    case_noisy("kern_ma5f2_d3_n20_noisy_precon", "Ma5f2", None, 20, 3, "precon", 1e-2, 5e-2, 350.0)
    case_noisy("kern_ratqu_d3_n20_noisy_precon", "RatQu", 1.5, 20, 3, "precon", 1e-2, 5e-2, 350.0)
    case_noisy("kern_ratqu_d3_n20_noisy_base", "RatQu", 1.5, 20, 3, "base", 1e-2, 5e-2, 350.0)
    json.dump(report, open(os.path.join(OUT, "golden_report_kernels.json"), "w"), indent=1)
