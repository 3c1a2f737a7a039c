"""GPU: the scenarios of the reference's own unit tests, through the same public calls.

gpgradpy/unit_test/test_grad_lkd.py:26-147 (2 points in 1-D, last point without gradient, theta = 1.5e-3, varK penalty
on; noise-free and UNKNOWN-noise variants where varK / var_fval / var_fgrad are hyper-parameters; every conditioning
mode).  That module does not import as shipped, so its scenarios were run through the reference by
oracle/make_golden_cond.py::case_ref_unit_lkd and the values stored in tests/golden/ref_unit_test_grad_lkd.npz; here
they are compared with the CUDA path, and the reference test's own forward-finite-difference check (eps 1e-8, rtol 1e-4,
atol 1e-5, :26-45,154-228) is repeated on the CUDA path."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("base", False), ("base", True), ("rescale_origin", False), ("rescale_eta_vary", False), ("precon", False),
         ("precon", True)]


def _setup(g, mode, noisy):
    from gpgradpy_b200.gp import GaussianProcess
    GP = GaussianProcess(1, True, "SqExp", mode)
    GP.lkd_varK_pnlt_use = True
    if noisy:
        GP.set_data(g["x"], g["fval"], None, g["grad"], None, g["mask"])
        mk = lambda th: GP.make_hp_class(None, th, np.nan, float(g["varK"]), float(g["var_fval"]), float(g["var_fgrad"]))
    else:
        GP.set_data(g["x"], g["fval"], np.zeros(2), g["grad"], np.zeros(g["grad"].shape), g["mask"])
        mk = lambda th: GP.make_hp_class(None, th, np.nan, None, None, None)
    GP.cond_max_target = 1e5
    GP.cond_eta_is_const = True
    return GP, mk


@pytest.mark.parametrize("mode,noisy", CASES)
def test_reference_unit_test_grad_lkd(golden_dir, mode, noisy):
    z = np.load(os.path.join(golden_dir, "ref_unit_test_grad_lkd.npz"))
    g = {k: z[k] for k in z.files}
    key = f"{mode}_{'noisy' if noisy else 'clean'}"
    GP, mk = _setup(g, mode, noisy)
    calc_cond = mode != "precon"
    info, ok = GP.calc_lkd_all(mk(g["theta"].copy()), calc_cond=calc_cond, calc_grad=True, lkd_use_adj_mtd=True)
    assert ok
    assert abs(info.ln_lkd - g[key + "_lkd"]) < 1e-8 * abs(g[key + "_lkd"])
    assert np.max(np.abs(info.ln_lkd_grad - g[key + "_grad"])) < 1e-7 * np.max(np.abs(g[key + "_grad"]))
    assert np.max(np.abs(info.hp_beta - g[key + "_beta"])) < 1e-8 * max(1.0, np.max(np.abs(g[key + "_beta"])))
    if calc_cond:
        assert abs(info.cond - g[key + "_cond"]) < 1e-8 * g[key + "_cond"]
        assert np.max(np.abs(info.cond_grad - g[key + "_cond_grad"])) < 1e-6 * np.max(np.abs(g[key + "_cond_grad"]))
    # the reference test's own check: forward finite difference in every optimised hyper-parameter
    eps, rtol, atol = 1e-8, 1e-4, 1e-5
    hi = GP.hp_info_optz_lkd
    pert = [("theta", int(np.min(hi.idx_theta)))]
    if noisy:
        pert += [("varK", hi.idx_varK), ("var_fval", hi.idx_var_fval), ("var_fgrad", hi.idx_var_fgrad)]
    for name, idx in pert:
        hp = mk(g["theta"].copy())
        if name == "theta":
            hp.theta[0] += eps
        else:
            setattr(hp, name, getattr(hp, name) + eps)
        info2 = GP.calc_lkd_all(hp, calc_cond=calc_cond, calc_grad=False)[0]
        np.testing.assert_allclose(info.ln_lkd_grad[idx], (info2.ln_lkd - info.ln_lkd) / eps, rtol=rtol,
                                   atol=max(atol, 1e-6 * abs(info.ln_lkd) / eps * 1e-8))
        if calc_cond:
            fd = (info2.cond - info.cond) / eps
            np.testing.assert_allclose(info.cond_grad[idx], fd, rtol=1e-3, atol=max(atol, 1e-7 * abs(info.cond) / eps))


def test_reference_unit_test_Kfull():
    """gpgradpy/unit_test/test_Kfull.py:25-118: the gradient-enhanced matrix against first / second finite differences
    of the base kernel (2 points in 2-D, theta = [1, 2], eps 1e-6, rtol = atol = 1e-4), on the CUDA builder."""
    from gpgradpy_b200 import backend as bk, _lib as L
    x = np.array([[0.1, 0.7], [0.9, 0.2]])
    th = np.array([1.0, 2.0])
    n, d = x.shape
    eps, tol = 1e-6, 1e-4
    K = bk.build_cov(x, th, mode=L.MODE_BASE, eta=0.0)[0].cpu().numpy()

    def kbase(xa, xb):                      # value block of a two-point build without gradients
        pts = np.vstack((xa, xb))
        slot = bk.slot_from_mask(np.zeros(2, bool), 2)[0]
        return bk.build_cov(pts, th, n_g=0, slot=slot, mode=L.MODE_BASE, eta=0.0)[0].cpu().numpy()[0, 1]

    for a in range(n):
        for b in range(n):
            for i in range(d):
                e = np.zeros(d); e[i] = eps
                fd_row = (kbase(x[a] + e, x[b]) - kbase(x[a] - e, x[b])) / (2 * eps)     # d k / d x_a,i
                fd_col = (kbase(x[a], x[b] + e) - kbase(x[a], x[b] - e)) / (2 * eps)     # d k / d x_b,i
                np.testing.assert_allclose(K[n + i * n + a, b], fd_row, rtol=tol, atol=tol)
                np.testing.assert_allclose(K[a, n + i * n + b], fd_col, rtol=tol, atol=tol)
                for j in range(d):
                    f = np.zeros(d); f[j] = 1e-4
                    fd2 = (kbase(x[a] + e, x[b] + f) - kbase(x[a] - e, x[b] + f)
                           - kbase(x[a] + e, x[b] - f) + kbase(x[a] - e, x[b] - f)) / (4 * eps * 1e-4)
                    np.testing.assert_allclose(K[n + i * n + a, n + j * n + b], fd2, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("noisy", [False, True])
@pytest.mark.parametrize("masked", [False, True])
def test_reference_unit_test_grad_Kmat(noisy, masked):
    """gpgradpy/unit_test/test_grad_Kmat.py:26-40,223-236: d Kcov / d hp against a forward finite difference of the
    matrix (3 points in 2-D scaled by 1e-3, theta = 10 [1.5, 3], eps 1e-6, rtol 1e-5, atol 1e-7; last point without
    gradient in the masked variant).  The CUDA path never materialises d Kcov / d hp, so the comparison is made on the
    quadratic forms v^T (d Kcov / d hp) v that gegp_quad_grad evaluates on the fly, for several random v."""
    from gpgradpy_b200 import backend as bk, _lib as L
    x = 1e-3 * np.array([[0.0, 0.0], [1.0, 0.3], [0.4, 1.0]])
    th = 10.0 * np.array([1.5, 3.0])
    n, d = x.shape
    mask = np.array([True, True, False]) if masked else np.ones(3, bool)
    slot, ng = bk.slot_from_mask(mask, n)
    N = n + ng * d
    eta, varK, eps = 1e-3, 4.0, 1e-6
    noise = None
    if noisy:
        noise = np.hstack((np.full(n, 3.0), np.full(ng * d, 4.0))) / varK

    def Kcov(theta, vk):
        nz = None if noise is None else noise * (varK / vk)     # noise / varK with the perturbed varK
        return bk.build_cov(x, theta, n_g=ng, slot=slot, noise=nz, mode=L.MODE_BASE, eta=eta,
                            varK=vk)[0].cpu().numpy()

    K0 = Kcov(th, varK if noisy else 1.0)
    rng = np.random.default_rng(3)
    for _ in range(3):
        v = rng.standard_normal(N)
        q = bk.quad_grad(x, th, v, n_g=ng, slot=slot, eta=eta, noisy=noisy, varK=varK).cpu().numpy()
        for m in range(d):
            tp = th.copy(); tp[m] += eps
            fd = v @ (Kcov(tp, varK if noisy else 1.0) - K0) @ v / eps
            np.testing.assert_allclose(q[L.OUT_GRAD + m], fd, rtol=1e-5, atol=1e-7 * max(1.0, abs(fd)))
        if noisy:
            fd = v @ (Kcov(th, varK + eps) - K0) @ v / eps
            np.testing.assert_allclose(q[L.OUT_DVARK], fd, rtol=1e-5, atol=1e-7 * max(1.0, abs(fd)))
            np.testing.assert_allclose(q[L.OUT_DVARF], np.sum(v[:n] ** 2), rtol=1e-12)
            np.testing.assert_allclose(q[L.OUT_DVARG], np.sum(v[n:] ** 2), rtol=1e-12)


@pytest.mark.parametrize("mode", ["base", "rescale_origin", "precon"])
def test_reference_unit_test_grad_surr(mode):
    """gpgradpy/unit_test/test_grad_surr.py:16-31,101-244: x-derivatives and Hessians of the surrogate against central
    finite differences (2 points in 1-D at 0 and 2, test point at 0.8, theta = 0.1, varK = 1e3, beta = mean(f),
    eps 1e-4, rtol 1e-4, atol 1e-8), in the three modes the shipped test names stand for
    (None -> base, 'req_vmin' -> rescale_origin, precon)."""
    from gpgradpy_b200.gp import GaussianProcess
    x = np.array([[0.0], [2.0]])
    f = np.sum(x ** 2, axis=1)
    g = 2 * x
    GP = GaussianProcess(1, True, "SqExp", mode)
    GP.init_optz_surr(2)
    GP.set_data(x, f, np.zeros(2), g, np.zeros((2, 1)))
    GP.set_hpara("set", 0, hp_vals=GP.make_hp_class(beta=np.atleast_1d(np.mean(f)), kernel=np.nan,
                                                    theta=0.1 * np.linspace(1, 2, 1), varK=1e3))
    GP.setup_eval_model()
    xt = np.array([[0.8]])
    eps, rtol, atol = 1e-4, 1e-4, 1e-8
    mu, sig, dmu, dsig, h_mu, h_sig = GP.eval_model(xt, calc_grad=True, calc_hess=True)
    mp, sp, dmp, dsp = GP.eval_model(xt + eps, calc_grad=True)[:4]
    mm, sm, dmm, dsm = GP.eval_model(xt - eps, calc_grad=True)[:4]
    np.testing.assert_allclose(dmu[0, 0], (mp[0] - mm[0]) / (2 * eps), rtol=rtol, atol=atol)
    np.testing.assert_allclose(dsig[0, 0], (sp[0] - sm[0]) / (2 * eps), rtol=rtol, atol=atol)
    np.testing.assert_allclose(h_mu[0, 0, 0], (dmp[0, 0] - dmm[0, 0]) / (2 * eps), rtol=rtol, atol=atol)
    np.testing.assert_allclose(h_sig[0, 0, 0], (dsp[0, 0] - dsm[0, 0]) / (2 * eps), rtol=rtol, atol=atol)
    # the surrogate interpolates values and gradients at the data (up to the nugget)
    m0, s0, d0 = GP.eval_model(x, calc_grad=True)[:3]
    assert np.max(np.abs(m0 - f)) < 1e-5 and np.max(np.abs(d0 - g)) < 1e-4
