"""Generate tests/golden/cond_*.npz by running the REFERENCE: 2-norm condition number of the factored matrix and its
hyper-parameter gradient (optz/GpHparaCon.py:161-235 through calc_lkd_all(calc_cond=True, calc_grad=True)), and
one hyper-parameter fit per conditioning mode with the condition-number constraint (optz/OptzLkd.py:185-333).

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).

    python oracle/make_golden_cond.py
"""
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


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(1e-300, float(np.max(np.abs(b)))))


def case_cond(name, n, d, mode, theta, seed=0, lo=-2.0, hi=2.0, std_f=0.0, std_g=0.0, varK=None):
    x, f, g = O.synthetic_problem(n, d, seed, lo, hi)
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.set_data(x, f, std_f * np.ones(n), g, std_g * np.ones(g.shape))
    th = np.asarray(theta, float)
    hp = GP.make_hp_class(theta=th, varK=varK)
    info, ok = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=True)
    cond_nograd = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=False)[0].cond
    xs = GP.get_scl_x_w_dist()[0]
    fs, _, gs, _ = GP.get_scl_eval_data()
    out = dict(x=x, fval=f, grad=g, theta=th, mode=mode, eta=GP._etaK, ok=ok, cond=info.cond, cond_nograd=cond_nograd, x_scl=xs, fval_scl=fs,
               grad_scl=gs, std_f=std_f, std_g=std_g, varK=np.nan if varK is None else varK,
               ln_lkd=info.ln_lkd, ln_lkd_grad=info.ln_lkd_grad)
    if info.cond_grad is not None:
        out["cond_grad"] = info.cond_grad
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    if varK is None:
        c, cg = O.cond_wo_noise(xs, th, mode, GP._etaK)
    else:
        c, cg = O.cond_w_noise(xs, th, varK, GP.calc_noise_vec(hp), mode, GP._etaK)
    c0 = (O.cond_wo_noise(xs, th, mode, GP._etaK, calc_grad=False)[0] if varK is None else
          O.cond_w_noise(xs, th, varK, GP.calc_noise_vec(hp), mode, GP._etaK, calc_grad=False)[0])
    print(name, f"cond {info.cond:.6e} oracle rel {rel(c, info.cond):.2e} nograd {cond_nograd:.6e} rel {rel(c0, cond_nograd):.2e}",
          "" if info.cond_grad is None else f"grad rel {rel(cg, info.cond_grad):.2e}", flush=True)


def case_eta_vary(name, n, d, theta, seed=0):
    """wellcond_mtd='rescale_eta_vary': nugget from the Gershgorin row sums (kernel/Kernel.py:269-274), v_min = 1 data."""
    x, f, g = O.synthetic_problem(n, d, seed)
    GP = GaussianProcess(d, True, "SqExp", "rescale_eta_vary")
    GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
    th = np.asarray(theta, float)
    hp = GP.make_hp_class(theta=th)
    info, ok = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=True)
    xs, Rt = GP.get_scl_x_w_dist()
    eta, idx = GP.calc_all_K_w_chofac(Rt, hp, varK=1)[5:7]
    rng = np.random.default_rng(11)
    xt = rng.uniform(-2, 2, (12, d))
    GP.set_hpara("set", 1, GP.make_hp_class(theta=th, varK=info.hp_varK, beta=info.hp_beta))
    mu, sig = GP.eval_model(xt)[:2]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, fval=f, grad=g, theta=th, eta=eta, idx=idx, ok=ok,
                        ln_lkd=info.ln_lkd, ln_lkd_grad=info.ln_lkd_grad, hp_varK=info.hp_varK, hp_beta=info.hp_beta,
                        cond=info.cond, cond_grad=info.cond_grad, x_test=xt, mu=mu, sig=sig, x_scl=xs)
    print(name, "eta", eta, "idx", idx, "lml", info.ln_lkd, "cond", info.cond, flush=True)


def case_cond_fro(name, n, d, theta, seed=0, std_f=0.0, std_g=0.0, varK=None):
    """cond_norm = 'fro' (optz/GpHparaCon.py:237-261) through calc_lkd_all(calc_cond=True, calc_grad=True), base mode."""
    x, f, g = O.synthetic_problem(n, d, seed)
    GP = GaussianProcess(d, True, "SqExp", "base")
    GP.cond_norm = "fro"
    GP.set_data(x, f, std_f * np.ones(n), g, std_g * np.ones(g.shape))
    th = np.asarray(theta, float)
    hp = GP.make_hp_class(theta=th, varK=varK)
    info, ok = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=True)
    c0 = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=False)[0].cond
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, fval=f, grad=g, theta=th, eta=GP._etaK, cond=info.cond,
                        cond_grad=info.cond_grad, cond_nograd=c0, std_f=std_f, std_g=std_g,
                        varK=np.nan if varK is None else varK, ln_lkd=info.ln_lkd)
    if varK is None:
        ka = O.all_K_w_chofac(x, th, "base", GP._etaK, None, 1.0, None, calc_chofac=False)
        c, cg = O.cond_fro_w_grad(ka.Kcov, O.kerngrad_hp(x, th, "base", GP._etaK))
    else:
        nv = GP.calc_noise_vec(hp)
        ka = O.all_K_w_chofac(x, th, "base", GP._etaK, nv, varK, None, calc_chofac=False)
        c, cg = O.cond_fro_w_grad(ka.Kcov, O.kcov_grad_hp_noisy(x, th, ka.Kern, "base", GP._etaK, varK, False, False))
    print(name, f"cond {info.cond:.6e} oracle rel {rel(c, info.cond):.2e} grad rel {rel(cg, info.cond_grad):.2e}", flush=True)


def case_surr_grad(name, n, d, mode, seed=0, mask=None):
    """eval_model(calc_grad=True): d mu / d x and d sig / d x (eval/GpEvalModel.py:170-173, 319-354)."""
    x, f, g = O.synthetic_problem(n, d, seed)
    g_in = g if mask is None else g[mask]
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.set_data(x, f, np.zeros(n), g_in, np.zeros(g_in.shape), mask)
    th = O.bench_theta(d)
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th))
    rng = np.random.default_rng(200 + seed)
    xt = rng.uniform(-2, 2, (10, d))
    xt[:2] = x[:2] + 1e-2
    GP.set_hpara("set", 1, GP.make_hp_class(theta=th, varK=info.hp_varK, beta=info.hp_beta))
    mu, sig, dmu, dsig = GP.eval_model(xt, calc_grad=True)[:4]
    extra = {}
    if not GP.b_use_data_scl:
        s2, ds2 = GP.eval_model_var(xt, calc_grad=True)[:2]
        extra = dict(sig2=s2, dsig2dx=ds2)
    xs = GP.get_scl_x_w_dist()[0]
    fs, _, gs, _ = GP.get_scl_eval_data()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, fval=f, grad=g_in, theta=th, mode=mode, eta=GP._etaK,
                        hp_varK=info.hp_varK, hp_beta=info.hp_beta, x_test=xt, mu=mu, sig=sig, dmudx=dmu, dsigdx=dsig,
                        mask=np.zeros(0, bool) if mask is None else mask, x_scl=xs, fval_scl=fs, grad_scl=gs, **extra)
    if mask is None and mode == "precon":
        o = O.eval_model_grad(xs, fs, gs, th, info.hp_varK, info.hp_beta, xt, mode, GP._etaK)
        print(name, "oracle rel dmu", rel(o[2], dmu), "dsig", rel(o[3], dsig), flush=True)
    else:
        print(name, "stored", flush=True)


def case_surr_hess(name, n, d, mode, seed=0, npts=4):
    """eval_model(calc_grad=True, calc_hess=True), one point per call (eval/GpEvalModel.py:175-180, 356-382)."""
    x, f, g = O.synthetic_problem(n, d, seed)
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.set_data(x, f, np.zeros(n), g, np.zeros(g.shape))
    th = O.bench_theta(d)
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th))
    rng = np.random.default_rng(300 + seed)
    xt = rng.uniform(-2, 2, (npts, d))
    GP.set_hpara("set", 1, GP.make_hp_class(theta=th, varK=info.hp_varK, beta=info.hp_beta))
    res = [GP.eval_model(xt[i], calc_grad=True, calc_hess=True) for i in range(npts)]
    xs = GP.get_scl_x_w_dist()[0]
    fs, _, gs, _ = GP.get_scl_eval_data()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x, fval=f, grad=g, theta=th, mode=mode, eta=GP._etaK,
                        hp_varK=info.hp_varK, hp_beta=info.hp_beta, x_test=xt, x_scl=xs, fval_scl=fs, grad_scl=gs,
                        mu=np.array([r[0][0] for r in res]), sig=np.array([r[1][0] for r in res]),
                        dmudx=np.array([r[2][0] for r in res]), dsigdx=np.array([r[3][0] for r in res]),
                        d2mudx2=np.array([r[4][0] for r in res]), d2sigdx2=np.array([r[5][0] for r in res]))
    if mode == "precon":
        o = O.eval_model_hess(xs, fs, gs, th, info.hp_varK, info.hp_beta, xt[0], mode, GP._etaK)
        print(name, "oracle rel d2mu", rel(o[4][0], res[0][4][0]), "d2sig", rel(o[5][0], res[0][5][0]), flush=True)
    else:
        print(name, "stored", flush=True)


def case_ref_unit_lkd():
    """The scenarios of the reference's own unit test gpgradpy/unit_test/test_grad_lkd.py:26-147 (2 points in 1-D,
    last point without gradient, theta = 1.5e-3, varK penalty on; noise-free with zero std, and noisy with UNKNOWN
    noise so that varK, var_fval, var_fgrad are hyper-parameters), run through the reference for every conditioning
    mode.  (The test module itself does not import as shipped: trailing module-level script)"""
    x = 3 * np.array([[0.0], [1.0]])
    f = np.sum(x ** 2, axis=1)
    g_all = 2 * x
    mask = np.array([True, False])
    g = g_all[mask, :]
    th = 0.001 * np.linspace(1.5, 3, 1)
    out = dict(x=x, fval=f, grad=g, mask=mask, theta=th, varK=4.0, var_fval=3.0, var_fgrad=4.0)
    for mode in ("base", "rescale_origin", "rescale_eta_vary", "precon"):
        for noisy in (False, True):
            if noisy and "rescale" in mode:
                continue                      # test_grad_lkd.py:253-254
            GP = GaussianProcess(1, True, "SqExp", mode)
            GP.lkd_varK_pnlt_use = True
            if noisy:
                GP.set_data(x, f, None, g, None, mask)
                hp = GP.make_hp_class(None, th.copy(), np.nan, 4.0, 3.0, 4.0)
            else:
                GP.set_data(x, f, np.zeros(2), g, np.zeros(g.shape), mask)
                hp = GP.make_hp_class(None, th.copy(), np.nan, None, None, None)
            GP.cond_max_target = 1e5
            GP.cond_eta_is_const = True
            info = GP.calc_lkd_all(hp, calc_cond=(mode != "precon"), calc_grad=True, lkd_use_adj_mtd=True)[0]
            key = f"{mode}_{'noisy' if noisy else 'clean'}"
            out[key + "_lkd"], out[key + "_grad"] = info.ln_lkd, info.ln_lkd_grad
            out[key + "_beta"] = info.hp_beta
            if mode != "precon":
                out[key + "_cond"], out[key + "_cond_grad"] = info.cond, info.cond_grad
            print("unit_lkd", key, info.ln_lkd, info.ln_lkd_grad, info.cond, flush=True)
    np.savez_compressed(os.path.join(OUT, "ref_unit_test_grad_lkd.npz"), **out)


def case_c1_grid():
    """BASELINE configs[0] = the recipe of gpgradpy/plt/plt_cond.py:19-207 (the script itself needs matplotlib): 2-D
    Rosenbrock, n = 20 points in [0.9, 1.1]^2, a grid of (gamma_1, gamma_2) = sqrt(2 theta) over [1e-4, 1e8], nugget of
    the preconditioned SqExp kernel used for every mode (:98-99,176-177), varK = 0.1; LML and condition number per grid
    point through calc_lkd_all(calc_lkd=True, calc_cond=True, calc_grad=False) (:189)."""
    n_eval, dim, n_gamma, varK = 20, 2, 6, 0.1
    x, f, g = O.synthetic_problem(n_eval, dim, 2, 0.9, 1.1)
    GP0 = GaussianProcess(dim, use_grad=True, kernel_type="SqExp", wellcond_mtd="precon")
    nugget = GP0.calc_nugget(n_eval)[1]
    theta_vec = np.logspace(np.log10(0.5 * 1e-4 ** 2), np.log10(0.5 * 1e8 ** 2), n_gamma)
    out = dict(x=x, fval=f, grad=g, theta_vec=theta_vec, nugget=nugget, varK=varK)
    for mode in ("base", "precon", "rescale_origin"):
        GP = GaussianProcess(dim, True, "SqExp", mode)
        GP.cond_eta_is_const = True
        GP.set_data(x, f, np.zeros(n_eval), g, np.zeros(g.shape))
        GP._etaK = nugget
        GP._eta_Kgrad = nugget
        GP.cond_max = 1e10
        lkd = np.full((n_gamma, n_gamma), np.nan)
        cond = np.full((n_gamma, n_gamma), np.nan)
        for i in range(n_gamma):
            for j in range(n_gamma):
                hp = GP.make_hp_class(varK=varK, theta=np.array([theta_vec[i], theta_vec[j]]), kernel=GP.hp_kernel_default)
                info = GP.calc_lkd_all(hp, calc_lkd=True, calc_cond=True, calc_grad=False)[0]
                lkd[i, j] = np.nan if info.ln_lkd is None else info.ln_lkd
                cond[i, j] = info.cond
        out[mode + "_lkd"], out[mode + "_cond"] = lkd, cond
        print("c1_grid", mode, "n_ok", int(np.sum(np.isfinite(lkd))), "log10 cond range",
              np.round(np.log10(np.nanmin(cond)), 2), np.round(np.log10(np.nanmax(cond)), 2), flush=True)
    np.savez_compressed(os.path.join(OUT, "c1_plt_cond_grid.npz"), **out)


def case_fit(name, n, d, mode, seed=1, std_f=0.0, std_g=0.0):
    """set_hpara('optz') with the history protocol of SURVEY appendix B.12; stores the start point the reference's
    40-candidate scan picked, so that both implementations can be started from the same x0."""
    x, f, g = O.synthetic_problem(n, d, seed)
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.init_optz_surr(3)
    GP.set_data(x[:1], f[:1], std_f * np.ones(1), g[:1], std_g * np.ones((1, d)))
    GP.set_hpara("optz", 0)
    GP.set_data(x, f, std_f * np.ones(n), g, std_g * np.ones((n, d)))
    hp_x0, bound, _ = GP.select_hp_optz_x0(1, GP.hp_info_optz_lkd)
    GP.set_hpara("optz", 1)
    hp = GP.hp_vals
    info, ok = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=False)
    out = dict(x=x, fval=f, grad=g, mode=mode, hp_x0=hp_x0, lb=bound.lb, ub=bound.ub, theta=hp.theta, varK=hp.varK,
               std_f=std_f, std_g=std_g,
               beta=hp.beta, ln_lkd=info.ln_lkd, cond=info.cond, eta=GP._etaK,
               xvec_scale=GP.DataScl.xvec_scale if GP.b_use_data_scl else np.ones(d))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "theta", hp.theta, "lml", info.ln_lkd, "cond", f"{info.cond:.3e}", flush=True)


if __name__ == "__main__":
    case_cond("cond_d2_n12_base", 12, 2, "base", [0.6, 1.4])
    case_cond("cond_d3_n20_base", 20, 3, "base", [0.9, 0.5, 1.6], seed=1)
    case_cond("cond_d2_n16_rescale_origin", 16, 2, "rescale_origin", [0.5, 0.9], seed=2)
    case_cond("cond_d3_n14_precon", 14, 3, "precon", [0.3, 0.2, 0.5], seed=3)
    case_cond("cond_d2_n14_noisy_base", 14, 2, "base", [0.7, 1.1], seed=4, std_f=1e-2, std_g=5e-2, varK=40.0)
    case_cond("cond_d4_n40_base_illcond", 40, 4, "base", [0.02, 0.03, 0.05, 0.04], seed=5)
    case_eta_vary("etavary_d3_n18", 18, 3, [0.4, 0.7, 0.3], seed=6)
    case_surr_grad("surrgrad_d3_n20_precon", 20, 3, "precon")
    case_surr_grad("surrgrad_d2_n15_rescale_origin", 15, 2, "rescale_origin", seed=1)
    m = np.zeros(14, bool); m[:9] = True
    case_surr_grad("surrgrad_d3_n14_mask", 14, 3, "precon", seed=2, mask=m)
    case_surr_hess("surrhess_d3_n16_precon", 16, 3, "precon")
    case_surr_hess("surrhess_d2_n12_rescale_origin", 12, 2, "rescale_origin", seed=1)
    case_cond_fro("condfro_d2_n12_base", 12, 2, [0.6, 1.4])
    case_cond_fro("condfro_d2_n14_noisy_base", 14, 2, [0.7, 1.1], seed=4, std_f=1e-2, std_g=5e-2, varK=40.0)
    case_ref_unit_lkd()
    case_c1_grid()
    if "--no-fit" in sys.argv:
        sys.exit(0)
    case_fit("fit_d2_n20_base", 20, 2, "base")
    case_fit("fit_d2_n20_rescale_origin", 20, 2, "rescale_origin")
    case_fit("fit_d2_n20_noisy_precon", 20, 2, "precon", std_f=5.0, std_g=20.0)
