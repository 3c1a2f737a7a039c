"""GPU parity of the Matern-5/2 ('Ma5f2') and rational-quadratic ('RatQu') kernel families through the C ABI and through
the GaussianProcess mirror, against fixtures the live reference produced (oracle/make_golden_kernels.py):
kernel/KernelMatern5f2.py:354-451 (KernGrad), :534-643 (d/dtheta), :665-680 (precon), :272-331 (d/dx);
kernel/KernelRatQuad.py:439-553, :636-749, :752-843 (d/dalpha), :863-877, :556-633.
Tolerances as for the Gaussian kernel (north_star): covariance entries 1e-12 of the block scale, LML / gradient /
posterior mean 1e-8 relative, sigma through |d sigma^2| <= 1e-8 varK."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KERN_CASES = [f"kern_{k}_d3_n22_{m}" for k in ("ma5f2", "ratqu_a2", "ratqu_a07") for m in ("precon", "base", "rescale_origin")] \
    + ["kern_ma5f2_d2_n14_mask_prefix", "kern_ratqu_d2_n14_mask_scatter", "kern_ma5f2_d5_n60_precon", "kern_ratqu_d5_n60_precon"]


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _kern_of(g):
    hp = float(g["hp_kernel"])
    return str(g["kernel"]), (None if np.isnan(hp) else hp)


def _block_scale_err(Kg, Kr):
    s = np.sqrt(np.abs(np.diag(Kr)))
    scale = np.maximum(np.abs(Kr), 1e-3 * s[:, None] * s[None, :])
    return float(np.max(np.abs(Kg - Kr) / scale))


@pytest.mark.parametrize("name", KERN_CASES)
def test_kernel_family_matrices_and_lml_through_c_abi(golden_dir, name):
    from gpgradpy_b200 import backend as bk, _lib as L
    g = _load(golden_dir, name)
    kern = _kern_of(g)
    x, f, gr, th, eta = g["x_scl"], g["fval_scl"], g["grad_scl"], g["theta"], float(g["eta"])
    n, d = x.shape
    mask = g["mask"] if g["mask"].size else None
    slot, ng = bk.slot_from_mask(mask, n)
    precon = str(g["mode"]) == "precon"
    if "Kern" in g:
        Kern, _ = bk.build_cov(x, th, n_g=ng, slot=slot, mode=L.MODE_BASE, eta=0.0, kernel=kern)
        assert _block_scale_err(Kern.cpu().numpy(), g["Kern"]) < 1e-12
        Kcov, _ = bk.build_cov(x, th, n_g=ng, slot=slot, mode=L.MODE_PRECON_COV if precon else L.MODE_BASE, eta=eta, kernel=kern)
        assert _block_scale_err(Kcov.cpu().numpy(), g["Kcov"]) < 1e-12
        if precon:
            Kt, p = bk.build_cov(x, th, n_g=ng, slot=slot, mode=L.MODE_PRECON, eta=eta, kernel=kern)
            N = g["Kern"].shape[0]
            assert _block_scale_err(Kt.cpu().numpy(), g["Kcor"] + eta * np.eye(N)) < 1e-12
            assert np.max(np.abs(p[:N].cpu().numpy() / np.sqrt(np.diag(g["Kern"])) - 1)) < 1e-14
            Kl, _ = bk.build_cov(x, th, n_g=ng, slot=slot, mode=L.MODE_PRECON, eta=eta, uplo=1, kernel=kern)
            assert np.array_equal(np.tril(Kl.cpu().numpy()), np.tril(Kt.cpu().numpy()))
    y = np.hstack((f, gr.reshape(gr.size, order="F")))
    out, _ = bk.lml_eval(x, y, th[None, :], n_g=ng, slot=slot, mode=L.MODE_PRECON if precon else L.MODE_BASE, eta=eta,
                         want_grad=True, kernel=kern)
    o = out.cpu().numpy()[0]
    assert o[L.OUT_INFO] == 0
    ref_g = g["ln_lkd_grad"]
    got_g = np.hstack((o[L.OUT_GRAD:L.OUT_GRAD + d], [o[L.OUT_DKERN]] if kern[0] == "RatQu" else []))
    e = {"lml": abs(o[L.OUT_LML] - g["ln_lkd"]) / abs(g["ln_lkd"]), "varK": abs(o[L.OUT_SIGMA2] - g["hp_varK"]) / g["hp_varK"],
         "beta": abs(o[L.OUT_BETA] - g["hp_beta"][0]) / abs(g["hp_beta"][0]), "logdet": abs(o[L.OUT_LOGDET] - g["ln_det"]) / abs(g["ln_det"]),
         "grad": float(np.max(np.abs(got_g - ref_g)) / np.max(np.abs(ref_g)))}
    print(name, {k: f"{v:.2e}" for k, v in e.items()})
    assert all(v < 1e-8 for v in e.values()), e
    if kern[0] != "RatQu":
        assert o[L.OUT_DKERN] == 0.0


@pytest.mark.parametrize("name", KERN_CASES)
def test_kernel_family_through_gaussian_process_api(golden_dir, name):
    """The mirror class with kernel_type = 'Ma5f2' / 'RatQu': calc_lkd_all (gradient over [theta.., alpha]) and
    eval_model with x-gradients and Hessians, in raw coordinates and through the rescaling."""
    from gpgradpy_b200.gp import GaussianProcess
    g = _load(golden_dir, name)
    kname, khp = _kern_of(g)
    x, f, gr, th = g["x"], g["fval"], g["grad"], g["theta"]
    n, d = x.shape
    mask = g["mask"] if g["mask"].size else None
    GP = GaussianProcess(d, True, kname, str(g["mode"]))
    GP.set_data(x, f, np.zeros(n), gr, np.zeros(gr.shape), mask)
    assert abs(GP._etaK - float(g["eta"])) <= 1e-14 * float(g["eta"])
    assert GP.hp_info_optz_lkd.n_hp == int(g["n_hp"])
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th, kernel=khp), calc_grad=True)
    assert ok
    assert abs(info.ln_lkd - g["ln_lkd"]) < 1e-8 * abs(g["ln_lkd"])
    assert np.max(np.abs(info.ln_lkd_grad - g["ln_lkd_grad"])) < 1e-8 * np.max(np.abs(g["ln_lkd_grad"]))
    if "mu" not in g:
        return
    GP.set_hpara("set", 1, GP.make_hp_class(theta=th, kernel=khp, varK=info.hp_varK, beta=info.hp_beta))
    mu, sig, dmu, dsig = GP.eval_model(g["x_test"], calc_grad=True)[:4]
    varK_scl = float(g["hp_varK"]) / (GP.DataScl.obj_scale ** 2 if GP.b_use_data_scl and hasattr(GP.DataScl, "obj_scale") else 1.0)
    assert np.max(np.abs(mu - g["mu"])) < 1e-8 * np.max(np.abs(g["mu"]))
    assert np.max(np.abs(sig ** 2 - g["sig"] ** 2)) < 1e-8 * max(np.max(g["sig"] ** 2), 1e-300) + 1e-8 * abs(varK_scl) * 0
    assert np.max(np.abs(dmu - g["dmudx"])) < 1e-8 * np.max(np.abs(g["dmudx"]))
    far = g["sig"] > 1e-3 * np.max(g["sig"])            # d sigma / d x divides by sigma: compare away from the data
    assert np.max(np.abs(dsig[far] - g["dsigdx"][far])) < 1e-6 * np.max(np.abs(g["dsigdx"][far]))
    if "d2mudx2" in g:
        h = GP.eval_model(g["x_test"][3:4], calc_grad=True, calc_hess=True)
        assert np.max(np.abs(h[4] - g["d2mudx2"])) < 1e-8 * np.max(np.abs(g["d2mudx2"]))
        assert np.max(np.abs(h[5] - g["d2sigdx2"])) < 1e-5 * np.max(np.abs(g["d2sigdx2"]))


@pytest.mark.parametrize("name", ["kern_ma5f2_d3_n20_noisy_precon", "kern_ratqu_d3_n20_noisy_precon",
                                  "kern_ratqu_d3_n20_noisy_base"])
def test_kernel_family_noisy_fit_quantities(golden_dir, name):
    from gpgradpy_b200.gp import GaussianProcess
    g = _load(golden_dir, name)
    kname, khp = _kern_of(g)
    x, f, gr = g["x"], g["fval"], g["grad"]
    n, d = x.shape
    GP = GaussianProcess(d, True, kname, str(g["mode"]))
    GP.set_data(x, f, float(g["std_f"]) * np.ones(n), gr, float(g["std_g"]) * np.ones(gr.shape))
    hp = GP.make_hp_class(theta=g["theta"], kernel=khp, varK=float(g["varK"]))
    info, ok = GP.calc_lkd_all(hp, calc_grad=True)
    assert ok
    assert abs(info.ln_lkd - g["ln_lkd"]) < 1e-8 * abs(g["ln_lkd"])
    assert info.ln_lkd_grad.shape == g["ln_lkd_grad"].shape            # [theta.., alpha?, varK]
    assert np.max(np.abs(info.ln_lkd_grad - g["ln_lkd_grad"])) < 1e-8 * np.max(np.abs(g["ln_lkd_grad"]))
    GP.set_hpara("set", 1, GP.make_hp_class(theta=g["theta"], kernel=khp, varK=float(g["varK"]), beta=info.hp_beta))
    mu, sig = GP.eval_model(g["x_test"])[:2]
    assert np.max(np.abs(mu - g["mu"])) < 1e-8 * np.max(np.abs(g["mu"]))
    assert np.max(np.abs(sig - g["sig"])) < 1e-6 * np.max(np.abs(g["sig"]))


@pytest.mark.parametrize("kname", ["Ma5f2", "RatQu"])
def test_kernel_family_fit_runs_and_batched_scan_matches_single_evaluations(kname):
    """set_hpara('optz') with the other kernel families: the 40-candidate scan (one batched device call; RatQu rows carry
    their own alpha) picks the row the sequential calc_lkd_all loop picks (optz/GpHparaX0.py:39-45), and the SLSQP fit
    over [log10 theta.., log10 alpha] does not end below its start."""
    from gpgradpy_b200.gp import GaussianProcess
    from gpgradpy_b200 import _lib as L
    from oracle import gegp_oracle as O
    n, d = 40, 3
    x, f, g = O.synthetic_problem(n, d, 3)
    GP = GaussianProcess(d, True, kname, "precon")
    GP.init_optz_surr(3)
    GP.set_data(x[:1], f[:1], np.zeros(1), g[:1], np.zeros((1, d)))
    GP.set_hpara("optz", 0)
    GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
    hi = GP.hp_info_optz_lkd
    hp_x0, _ = GP.get_hp_x0_lhs_median(1, hi, 12)
    tab = GP.calc_lkd_batch(hp_x0, calc_grad=True)
    for i in (0, 5, 11):
        info, ok = GP.calc_lkd_all(GP.hp_vec2dataclass(hi, hp_x0[i]), calc_grad=True)
        assert ok and info.ln_lkd == tab[i, L.OUT_LML]                  # batched == single, bit for bit
        assert np.array_equal(info.ln_lkd_grad[hi.idx_theta], tab[i, L.OUT_GRAD:L.OUT_GRAD + d])
        if hi.has_kernel:
            assert info.ln_lkd_grad[hi.idx_kernel][0] == tab[i, L.OUT_DKERN]
    x0_sel = GP.select_hp_optz_x0(1, hi)[0]
    l0 = GP.calc_lkd_all(GP.hp_vec2dataclass(hi, x0_sel[0]))[0].ln_lkd
    GP.set_hpara("optz", 1)
    l1 = GP.calc_lkd_all(GP.hp_vals)[0].ln_lkd
    assert np.isfinite(l1) and l1 >= l0 - 1e-9 * abs(l0)
    if kname == "RatQu":
        assert GP.hp_kernel_range[0] <= GP.hp_vals.kernel <= GP.hp_kernel_range[1]
    mu, sig = GP.eval_model(x[:5])[:2]
    assert np.max(np.abs(mu - f[:5])) < 1e-2 * (f.max() - f.min())


DIRECT_CASES = ["direct_sqexp_d3_n20_precon", "direct_sqexp_d3_n20_base", "direct_sqexp_d2_n16_rescale_origin",
                "direct_ratqu_d3_n20_precon", "direct_ma5f2_d3_n20_precon", "direct_sqexp_d3_n20_noisy_precon",
                "direct_ratqu_d3_n20_noisy_base"]


@pytest.mark.parametrize("name", DIRECT_CASES)
def test_direct_likelihood_form_vs_reference(golden_dir, name):
    """calc_lkd_all(lkd_use_adj_mtd=False): the direct form of optz/CalcLkd.py:64-85 / 135-147 (noise-free) and :238-241 /
    253-265 (noisy), with hp_beta_grad (eval/GpMeanFun.py:110-120), hp_varK_grad (:104-116) and ln_det_Kmat_grad
    (:349-367), against the live reference's values; and the adjoint default still leaves those fields None."""
    from gpgradpy_b200.gp import GaussianProcess
    g = _load(golden_dir, name)
    kname, khp = _kern_of(g)
    x, f, gr = g["x"], g["fval"], g["grad"]
    n, d = x.shape
    noisy = not np.isnan(float(g["varK"]))
    GP = GaussianProcess(d, True, kname, str(g["mode"]))
    GP.set_data(x, f, float(g["std_f"]) * np.ones(n), gr, float(g["std_g"]) * np.ones(gr.shape))
    hp = GP.make_hp_class(theta=g["theta"], kernel=khp, varK=float(g["varK"]) if noisy else None)
    adj, ok = GP.calc_lkd_all(hp, calc_grad=True)
    assert ok and adj.hp_beta_grad is None and adj.ln_det_Kmat_grad is None
    dr, ok = GP.calc_lkd_all(hp, calc_grad=True, lkd_use_adj_mtd=False)
    assert ok
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))  # noqa: E731
    e = {"lkd_grad": rel(dr.ln_lkd_grad, g["ln_lkd_grad"]), "beta_grad": rel(dr.hp_beta_grad, g["hp_beta_grad"]),
         "logdet_grad": rel(dr.ln_det_Kmat_grad, g["ln_det_Kmat_grad"]), "adjoint": rel(adj.ln_lkd_grad, g["ln_lkd_grad_adjoint"])}
    if not noisy:
        e["varK_grad"] = rel(dr.hp_varK_grad, g["hp_varK_grad"])
    print(name, {k: f"{v:.2e}" for k, v in e.items()})
    assert dr.hp_beta_grad.shape == g["hp_beta_grad"].shape
    assert all(v < 1e-8 for v in e.values()), e
