"""GPU: edge cases of the hot path against the CPU oracle -- a single data point, odd point counts (generic builder
instead of the fast one), many dimensions with few points, no gradient at any point, leaf-size boundaries, extreme
length scales, coincident points (collision: singular without the nugget), and the failure contract."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(n, d, seed=0, mode="precon", theta=None, mask=None, tol=1e-8, sig_tol=1e-6):
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(n, d, seed)
    th = O.bench_theta(d) if theta is None else np.asarray(theta, float)
    gin = g if mask is None else g[mask]
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.set_data(x, f, np.zeros(n), gin, np.zeros(gin.shape), mask)
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th), calc_grad=True)
    ref = O.lkd_wo_noise(x, f, gin, th, mode, GP._etaK, mask=mask)
    assert ok == ref.chofac_good
    if not ok:
        return GP, info, ref
    assert abs(info.ln_lkd - ref.ln_lkd) < tol * abs(ref.ln_lkd)
    assert abs(info.hp_varK - ref.hp_varK) < 10 * tol * ref.hp_varK
    assert np.max(np.abs(info.ln_lkd_grad - ref.ln_lkd_grad)) < 10 * tol * np.max(np.abs(ref.ln_lkd_grad))
    GP.set_hpara("set", 1, GP.make_hp_class(theta=th, varK=info.hp_varK, beta=info.hp_beta))
    xs = np.random.default_rng(9).uniform(-2, 2, (7, d))
    mu, sig = GP.eval_model(xs)[:2]
    mu_r, sig_r, _, _ = O.eval_model(x, f, gin, th, ref.hp_varK, ref.hp_beta, xs, mode, GP._etaK, mask=mask)
    assert np.max(np.abs(mu - mu_r)) < 100 * tol * max(1.0, np.max(np.abs(mu_r)))
    assert np.max(np.abs(sig - sig_r)) < sig_tol * max(1e-300, np.max(np.abs(sig_r)))
    return GP, info, ref


def test_single_point():
    """n = 1 (N = 1 + d): eta falls back to eta_base (base/GpWellCond.py:124-125); the fit is skipped for it
    (optz/GpHparaOptz.py:152-157) but LML and the posterior are defined."""
    _check(1, 3)
    _check(1, 1)


@pytest.mark.parametrize("n,d", [(33, 3), (127, 1), (5, 2), (43, 2)])
def test_odd_point_counts(n, d):
    """odd n: the 16-byte-aligned fast builder does not apply; N around the 128 leaf size (127*2 = 254, 43*3 = 129)."""
    # dense 1-D data: 1 - k*^T K^-1 k* cancels to ~1e-6, sigma is only defined to ~1e-5 relative (cond(K) ~ 1/eta)
    _check(n, d, sig_tol=1e-4 if d == 1 else 1e-6)


def test_many_dimensions_few_points():
    _check(12, 50, tol=1e-7)


@pytest.mark.parametrize("N_target", [128, 256, 384])
def test_leaf_size_boundaries(N_target):
    """N exactly a multiple of the 128 leaf (n (d + 1) = 128, 256, 384 with d = 1)."""
    _check(N_target // 2, 1, sig_tol=1e-4)


def test_no_point_has_a_gradient():
    """use_grad=True with an all-False mask: set_data takes np.max of the empty gradient-noise array and raises
    ValueError in the reference (GaussianProcess.py:317); same error behaviour here (the supported way to drop all
    gradients is use_grad=False, tests/test_gpu_api.py::test_gradient_free_gp)."""
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(20, 3, 0)
    GP = GaussianProcess(3, True, "SqExp", "precon")
    with pytest.raises(ValueError):
        GP.set_data(x, f, np.zeros(20), g[:0], np.zeros((0, 3)), np.zeros(20, bool))


@pytest.mark.parametrize("scale", [1e-7, 1e3])
def test_extreme_length_scales(scale):
    """theta -> 0 (K -> rank one, only the nugget keeps it definite) and theta large (K -> diagonal)."""
    _check(24, 2, theta=scale * np.array([1.0, 2.0]), tol=1e-6, sig_tol=1e-4)


def test_coincident_points():
    """Two identical data points (collision): K + eta P^2 is still positive definite thanks to the nugget, but only
    just; the CUDA path must agree with LAPACK on whether the factorisation succeeds and, if it does, on the values to
    the accuracy cond(K) ~ 1/eta allows."""
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(16, 2, 1)
    x[7] = x[3]; f[7] = f[3]; g[7] = g[3]
    th = O.bench_theta(2)
    GP = GaussianProcess(2, True, "SqExp", "precon")
    GP.set_data(x, f, np.zeros(16), g, np.zeros((16, 2)))
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th), calc_grad=True)
    ref = O.lkd_wo_noise(x, f, g, th, "precon", GP._etaK)
    assert ok == ref.chofac_good
    if ok:
        assert abs(info.ln_lkd - ref.ln_lkd) < 1e-5 * abs(ref.ln_lkd)
    # without any nugget the matrix is exactly singular: failure contract (info > 0 -> b_chofac_good False, cond returned)
    GP2 = GaussianProcess(2, True, "SqExp", "base")
    GP2.cond_eta_set_mtd, GP2.cond_eta_dflt = "dflt_eta", 0.0
    GP2.set_data(x, f, np.zeros(16), g, np.zeros((16, 2)))
    info2, ok2 = GP2.calc_lkd_all(GP2.make_hp_class(theta=th), calc_cond=True, calc_grad=True)
    if not ok2:
        assert info2.ln_lkd is None and np.isfinite(info2.cond) and info2.cond > 1e10
    tup = GP2.calc_all_K_w_chofac(None, GP2.make_hp_class(theta=th), varK=1)
    assert (tup[3] is None) == (not ok2)
