"""GPU: the GaussianProcess API (the drop-in boundary) against the reference outputs in tests/golden.

These read like the reference's own use: GP.set_data -> GP.calc_lkd_all / calc_all_K_w_chofac ->
GP.set_hpara -> GP.eval_model (gpgradpy/plt/plt_cond.py:154-207, eval/GpEvalModel.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _gp(g, mode, mask=None, std_f=0.0, std_g=0.0):
    from gpgradpy_b200.gp import GaussianProcess
    x = g["x"]
    n, d = x.shape
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.set_data(x, g["fval"], std_f * np.ones(n), g["grad"], std_g * np.ones(g["grad"].shape), mask)
    return GP


@pytest.mark.parametrize("name,mode", [("d4_n37_precon", "precon"), ("d3_n30_base", "base"),
                                       ("d3_n25_rescale_origin", "rescale_origin"), ("c1_d2_n20_precon", "precon")])
def test_calc_lkd_all_and_K_tuple(golden_dir, name, mode):
    g = _load(golden_dir, name)
    GP = _gp(g, mode)
    hp = GP.make_hp_class(theta=g["theta"])
    info, ok = GP.calc_lkd_all(hp, calc_grad=True)
    assert ok
    tol = 1e-6 if name.startswith("c1_") else 1e-8
    assert abs(info.ln_lkd - g["ln_lkd"]) < 1e-8 * abs(g["ln_lkd"])
    assert abs(info.hp_varK - g["hp_varK"]) < tol * g["hp_varK"]
    assert np.max(np.abs(info.ln_lkd_grad - g["ln_lkd_grad"])) < tol * np.max(np.abs(g["ln_lkd_grad"]))
    assert info.hp_beta.shape == (1,)
    # 7-tuple of calc_all_K_w_chofac (kernel/Kernel.py:307)
    Kern, Kcor, Kcov, fac, condK, etaK, idx = GP.calc_all_K_w_chofac(GP.get_scl_x_w_dist()[1], hp, varK=1)
    assert etaK == float(g["eta"]) and condK is None and idx is None
    assert np.allclose(np.asarray(Kern), g["Kern"], rtol=1e-11, atol=1e-13 * np.abs(g["Kern"]).max())
    assert np.allclose(np.asarray(Kcov), g["Kcov"], rtol=1e-11, atol=1e-13 * np.abs(g["Kcov"]).max())
    if mode == "precon":
        assert np.allclose(np.asarray(Kcor), g["Kcor"], rtol=1e-11, atol=1e-13)
        assert fac[1] is True
        Lg = np.tril(np.asarray(fac[0]))
    else:
        assert Kcor is None and fac[1] is False
        Lg = np.triu(np.asarray(fac[0])).T
    assert np.max(np.abs(Lg @ Lg.T - g["Kcov"])) < 1e-12 * np.max(np.abs(g["Kcov"]))
    assert np.max(np.abs(Lg - g["chol_lower"])) < 1e-6 * np.max(np.abs(g["chol_lower"]))


@pytest.mark.parametrize("name,mode", [("d4_n37_precon", "precon"), ("d3_n30_base", "base"),
                                       ("d3_n25_rescale_origin", "rescale_origin")])
def test_set_hpara_and_eval_model(golden_dir, name, mode):
    g = _load(golden_dir, name)
    GP = _gp(g, mode)
    hp = GP.make_hp_class(theta=g["theta"], varK=float(g["hp_varK"]), beta=g["hp_beta"])
    GP.set_hpara("set", 1, hp)
    mu, sig, a, b, c, e = GP.eval_model(g["x_test"])
    assert a is None and b is None and c is None and e is None
    assert np.max(np.abs(mu - g["mu"])) < 1e-8 * np.max(np.abs(g["mu"]))
    scale = np.max(g["sig"])
    assert np.max(np.abs(sig - g["sig"])) < 1e-6 * scale
    m1, s1 = GP.eval_model(g["x_test"][0], squeeze_nx=True)[:2]
    assert np.isscalar(m1) or np.ndim(m1) == 0
    assert abs(m1 - mu[0]) < 1e-12 * max(1.0, abs(mu[0]))
    GP.hp_vals = GP.make_hp_class(theta=2 * g["theta"], varK=1.0, beta=g["hp_beta"])
    with pytest.raises(Exception):
        GP.eval_model(g["x_test"])          # hp changed after setup_eval_model (eval/GpEvalModel.py:105-111)


def test_noisy_path(golden_dir):
    g = _load(golden_dir, "d3_n24_noisy_precon")
    GP = _gp(g, "precon", std_f=float(g["std_f"]), std_g=float(g["std_g"]))
    assert GP.b_has_noisy_data and GP.hp_info_optz_lkd.has_varK
    hp = GP.make_hp_class(theta=g["theta"], varK=float(g["varK"]))
    info, ok = GP.calc_lkd_all(hp, calc_grad=True)
    assert ok and abs(info.ln_lkd - g["ln_lkd"]) < 1e-8 * abs(g["ln_lkd"])
    ref = g["ln_lkd_grad"]
    assert np.max(np.abs(info.ln_lkd_grad - ref) / np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())) < 1e-8
    hp2 = GP.make_hp_class(theta=g["theta"], varK=float(g["varK"]), beta=g["hp_beta"])
    GP.set_hpara("set", 1, hp2)
    mu, sig = GP.eval_model(g["x_test"])[:2]
    assert np.max(np.abs(mu - g["mu"])) < 1e-8 * np.max(np.abs(g["mu"]))
    assert np.max(np.abs(sig - g["sig"])) < 1e-6 * np.max(g["sig"])


def test_masked_gradients(golden_dir):
    g = _load(golden_dir, "d3_n18_mask_prefix")
    GP = _gp(g, "precon", mask=g["mask"])
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=g["theta"]), calc_grad=True)
    assert ok and abs(info.ln_lkd - g["ln_lkd"]) < 1e-8 * abs(g["ln_lkd"])
    assert np.max(np.abs(info.ln_lkd_grad - g["ln_lkd_grad"])) < 1e-8 * np.max(np.abs(g["ln_lkd_grad"]))


def test_fit_precon_small():
    """set_hpara('optz'): history protocol of SURVEY appendix B.12, then the optimum must beat the start,
    have a small projected gradient, and agree with the CPU oracle at the optimum."""
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(24, 2, 1)
    GP = GaussianProcess(2, True, "SqExp", "precon")
    GP.init_optz_surr(3)
    GP.set_data(x[:1], f[:1], np.zeros(1), g[:1], np.zeros((1, 2)))
    GP.set_hpara("optz", 0)                      # one point: stores theta = 1e-2, no optimisation
    assert np.allclose(GP.hp_theta_all[0], 1e-2)
    GP.set_data(x, f, np.zeros(24), g, np.zeros((24, 2)))
    GP.set_hpara("optz", 1)
    th = GP.hp_vals.theta
    assert th.shape == (2,) and np.all(th > 0) and GP.hp_vals.varK > 0
    eta = GP._etaK
    ref = O.lkd_wo_noise(x, f, g, th, "precon", eta)
    info, ok = GP.calc_lkd_all(GP.hp_vals, calc_grad=True)
    assert ok and abs(info.ln_lkd - ref.ln_lkd) < 1e-8 * abs(ref.ln_lkd)
    assert abs(GP.hp_vals.varK - ref.hp_varK) < 1e-7 * ref.hp_varK
    lml_start = O.lkd_wo_noise(x, f, g, 1e-2 * np.ones(2), "precon", eta, calc_grad=False).ln_lkd
    assert info.ln_lkd >= lml_start
    glog = info.ln_lkd_grad * th * np.log(10)
    assert np.max(np.abs(glog)) < 1e-2 * max(1.0, abs(info.ln_lkd))
    mu, sig = GP.eval_model(x[:5] + 1e-4)[:2]
    assert np.max(np.abs(mu - f[:5])) < 1e-2 * np.max(np.abs(f)) + 1e-6


def test_gradient_free_gp():
    """use_grad=False (mode forced to 'base'): LML + gradient against the oracle with an all-False mask."""
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(30, 3, 2)
    GP = GaussianProcess(3, False, "SqExp", "precon")
    GP.set_data(x, f, np.zeros(30))
    th = O.bench_theta(3)
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th), calc_grad=True)
    mask = np.zeros(30, bool)
    ref = O.lkd_wo_noise(x, f, np.zeros((0, 3)), th, "base", GP._etaK, mask=mask)
    assert ok and abs(info.ln_lkd - ref.ln_lkd) < 1e-8 * abs(ref.ln_lkd)
    assert np.max(np.abs(info.ln_lkd_grad - ref.ln_lkd_grad)) < 1e-7 * np.max(np.abs(ref.ln_lkd_grad))


@pytest.mark.parametrize("name", ["surrgrad_d3_n20_precon", "surrgrad_d2_n15_rescale_origin", "surrgrad_d3_n14_mask"])
def test_eval_model_x_gradients(golden_dir, name):
    """eval_model(calc_grad=True) -> (mu, sig, dmudx, dsigdx, None, None) against the reference
    (eval/GpEvalModel.py:170-173, 319-354), incl. a rescale mode (derivatives mapped back to the initial
    coordinates) and partial gradients; plus a central finite difference of mu and sig themselves."""
    g = _load(golden_dir, name)
    mode = str(g["mode"])
    mask = g["mask"] if g["mask"].size else None
    GP = _gp(g, mode, mask)
    GP.set_hpara("set", 1, GP.make_hp_class(theta=g["theta"], varK=float(g["hp_varK"]), beta=g["hp_beta"]))
    xt = g["x_test"]
    mu, sig, dmu, dsig, h1, h2 = GP.eval_model(xt, calc_grad=True)
    assert h1 is None and h2 is None and dmu.shape == xt.shape and dsig.shape == xt.shape
    assert np.max(np.abs(mu - g["mu"])) < 1e-8 * np.max(np.abs(g["mu"]))
    assert np.max(np.abs(sig - g["sig"])) < 1e-6 * np.max(np.abs(g["sig"]))
    assert np.max(np.abs(dmu - g["dmudx"])) < 1e-8 * np.max(np.abs(g["dmudx"]))
    assert np.max(np.abs(dsig - g["dsigdx"])) < 1e-6 * np.max(np.abs(g["dsigdx"]))
    eps = 1e-5
    for j in range(xt.shape[1]):
        e = np.zeros(xt.shape[1]); e[j] = eps
        mp, sp = GP.eval_model(xt + e)[:2]
        mm, sm = GP.eval_model(xt - e)[:2]
        assert np.max(np.abs((mp - mm) / (2 * eps) - dmu[:, j])) < 1e-5 * np.max(np.abs(dmu))
        assert np.max(np.abs((sp - sm) / (2 * eps) - dsig[:, j])) < 1e-4 * np.max(np.abs(dsig))
    if "sig2" in g:     # eval_model_var (eval/GpEvalModel.py:200-317): variance and its x-gradient
        s2, ds2, h = GP.eval_model_var(xt, calc_grad=True)
        assert h is None
        assert np.max(np.abs(s2 - g["sig2"])) < 1e-6 * np.max(np.abs(g["sig2"]))
        assert np.max(np.abs(ds2 - g["dsig2dx"])) < 1e-6 * np.max(np.abs(g["dsig2dx"]))
    m1, s1, d1, ds1 = GP.eval_model(xt[3], calc_grad=True, squeeze_nx=True)[:4]
    assert np.isscalar(m1) or m1.shape == ()
    assert d1.shape == (xt.shape[1],) and np.allclose(d1, dmu[3]) and np.allclose(ds1, dsig[3])


@pytest.mark.parametrize("name", ["surrhess_d3_n16_precon", "surrhess_d2_n12_rescale_origin"])
def test_eval_model_hessians(golden_dir, name):
    """eval_model(calc_grad=True, calc_hess=True), one point per call like the reference
    (eval/GpEvalModel.py:175-180, 356-382): d2mudx2, d2sigdx2 [1, d, d]; plus a finite difference of the gradients."""
    g = _load(golden_dir, name)
    mode = str(g["mode"])
    GP = _gp(g, mode)
    GP.set_hpara("set", 1, GP.make_hp_class(theta=g["theta"], varK=float(g["hp_varK"]), beta=g["hp_beta"]))
    d = g["x"].shape[1]
    for p in range(g["x_test"].shape[0]):
        xt = g["x_test"][p]
        mu, sig, dmu, dsig, h_mu, h_sig = GP.eval_model(xt, calc_grad=True, calc_hess=True)
        assert h_mu.shape == (1, d, d) and h_sig.shape == (1, d, d)
        assert abs(mu[0] - g["mu"][p]) < 1e-8 * abs(g["mu"][p]) and abs(sig[0] - g["sig"][p]) < 1e-6 * abs(g["sig"][p])
        assert np.max(np.abs(dmu[0] - g["dmudx"][p])) < 1e-8 * np.max(np.abs(g["dmudx"][p]))
        assert np.max(np.abs(h_mu[0] - g["d2mudx2"][p])) < 1e-8 * np.max(np.abs(g["d2mudx2"][p]))
        assert np.max(np.abs(h_sig[0] - g["d2sigdx2"][p])) < 1e-5 * np.max(np.abs(g["d2sigdx2"][p]))
    xt = g["x_test"][0]
    _, _, _, _, h_mu, h_sig = GP.eval_model(xt, calc_grad=True, calc_hess=True)
    eps = 1e-5
    for j in range(d):
        e = np.zeros(d); e[j] = eps
        gp_, sp_ = GP.eval_model(xt + e, calc_grad=True)[2:4]
        gm_, sm_ = GP.eval_model(xt - e, calc_grad=True)[2:4]
        assert np.max(np.abs((gp_[0] - gm_[0]) / (2 * eps) - h_mu[0][:, j])) < 1e-5 * np.max(np.abs(h_mu))
        assert np.max(np.abs((sp_[0] - sm_[0]) / (2 * eps) - h_sig[0][:, j])) < 1e-4 * np.max(np.abs(h_sig))
    out = GP.eval_model(xt, calc_grad=True, calc_hess=True, squeeze_nx=True)
    assert out[4].shape == (d, d) and out[5].shape == (d, d)


def test_multistart_lockstep_equals_sequential_loop():
    """lkd_optz_start_mtd='lhs' (5 SLSQP starts, optz/OptzLkd.py:249-270): the lock-step batched driver selects
    exactly the optimum of the sequential loop (a batched candidate is evaluated bit-identically to a lone one)."""
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(40, 3, 4)
    out = []
    for lock in (True, False):
        GP = GaussianProcess(3, True, "SqExp", "precon")
        GP.lkd_optz_start_mtd = "lhs"
        GP.lockstep_multistart = lock
        GP.init_optz_surr(3)
        GP.set_data(x[:1], f[:1], np.zeros(1), g[:1], np.zeros((1, 3)))
        GP.set_hpara("optz", 0)
        GP.set_data(x, f, np.zeros(40), g, np.zeros((40, 3)))
        GP.set_hpara("optz", 1)
        out.append((GP.hp_vals.theta.copy(), GP.hp_vals.varK, GP.hp_optz_iter_mean[1], getattr(GP, "_lockstep_stats", None)))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1] and out[0][2] == out[1][2]
    st = out[0][3]
    assert st is not None and st["batch_sizes"][0] == 5 and st["n_batches"] < st["n_evals"]


def test_gradient_free_gp_posterior_and_x_gradients():
    """use_grad=False: posterior mean / std and their x-gradients (the gradient-enhanced kernels with n_g = 0)
    against the oracle with an all-False mask and against central finite differences."""
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(30, 3, 2)
    GP = GaussianProcess(3, False, "SqExp", "precon")
    GP.set_data(x, f, np.zeros(30))
    th = O.bench_theta(3) * 4
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th), calc_grad=True)
    assert ok
    GP.set_hpara("set", 1, GP.make_hp_class(theta=th, varK=info.hp_varK, beta=info.hp_beta))
    xs = np.random.default_rng(4).uniform(-2, 2, (9, 3))
    mu, sig, dmu, dsig = GP.eval_model(xs, calc_grad=True)[:4]
    mask = np.zeros(30, bool)
    mu_r, sig_r, dmu_r, dsig_r = O.eval_model_grad(x, f, np.zeros((0, 3)), th, info.hp_varK, info.hp_beta, xs, "base",
                                                   GP._etaK, mask=mask)
    assert np.max(np.abs(mu - mu_r)) < 1e-8 * np.max(np.abs(mu_r))
    assert np.max(np.abs(sig - sig_r)) < 1e-6 * np.max(np.abs(sig_r))
    assert np.max(np.abs(dmu - dmu_r)) < 1e-8 * np.max(np.abs(dmu_r))
    assert np.max(np.abs(dsig - dsig_r)) < 1e-6 * np.max(np.abs(dsig_r))
    eps = 1e-5
    for j in range(3):
        e = np.zeros(3); e[j] = eps
        mp, sp = GP.eval_model(xs + e)[:2]
        mm, sm = GP.eval_model(xs - e)[:2]
        assert np.max(np.abs((mp - mm) / (2 * eps) - dmu[:, j])) < 1e-5 * np.max(np.abs(dmu))
        assert np.max(np.abs((sp - sm) / (2 * eps) - dsig[:, j])) < 1e-4 * np.max(np.abs(dsig))


@pytest.mark.parametrize("mode,noisy", [("rescale_origin", False), ("base", False), ("precon", True), ("base", True)])
def test_candidate_scan_is_batched_in_every_mode(mode, noisy):
    """Batch point A (optz/GpHparaX0.py:39-45) outside the noise-free precon case: the 40 candidate rows are evaluated
    in ONE batched device call; in the base / rescale modes the condition number (optz/GpHparaX0.py:47-57) is only
    computed down the LML ranking until a feasible row is found.  Same selected start point as the reference's
    sequential loop (which this class still runs when rows do not share nugget / noise)."""
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    n, d = 60, 3
    x, f, g = O.synthetic_problem(n, d, 5)
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.init_optz_surr(3)
    sf, sg = (1e-2, 5e-2) if noisy else (0.0, 0.0)
    GP.set_data(x[:1], f[:1], sf * np.ones(1), g[:1], sg * np.ones((1, d)))
    GP.set_hpara("optz", 0)
    GP.set_data(x, f, sf * np.ones(n), g, sg * np.ones((n, d)))
    hi = GP.hp_info_optz_lkd
    assert GP._can_batch_scan(hi)
    x0_b = GP.select_hp_optz_x0(1, hi)[0]
    stats = dict(GP._scan_stats)
    assert stats["batched"] and stats["n_cond_evals"] <= 2          # 1 batched call + at most 2 condition numbers
    GP._can_batch_scan = lambda info: False
    x0_s = GP.select_hp_optz_x0(1, hi)[0]
    assert not GP._scan_stats["batched"]
    assert np.array_equal(x0_b, x0_s)
