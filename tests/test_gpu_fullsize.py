"""GPU: BASELINE full sizes (configs[1] N=5500 and the north-star configs[2] N=21000), where the reference cannot run
(its dK/dtheta tensor alone is 70.6 GB at N=21000): size-independent properties instead of an oracle run --
K^-1 K v = v with the explicit inverse the evaluation leaves behind, the gradient against a central finite
difference of the LML itself, bit-reproducibility, and interpolation of values AND gradients at the training points
by the posterior (a gradient-enhanced GP reproduces both)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d", [(500, 10), (1000, 20)])
def test_full_size_properties(n, d):
    import torch
    from gpgradpy_b200 import backend as bk, _lib as L
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    N = n * (d + 1)
    x, f, g = O.synthetic_problem(n, d, 0)
    th = O.bench_theta(d)
    GP = GaussianProcess(d, True, "SqExp", "precon")
    GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
    hp = GP.make_hp_class(theta=th)
    info, ok = GP.calc_lkd_all(hp, calc_grad=True)
    assert ok and np.isfinite(info.ln_lkd) and info.hp_varK > 0

    # (1) the explicit inverse in the workspace really is the inverse of the matrix that was factored
    v = bk.lml_views(n, n, d)
    Kinv = v["Kinv"].clone()
    K = bk.build_cov(GP._X_dev, th, mode=L.MODE_PRECON, eta=GP._etaK)[0]
    ld = bk.ld_of(N)
    rng = np.random.default_rng(1)
    z = bk.to_dev(np.pad(rng.standard_normal(N), (0, ld - N)))
    t1 = torch.zeros(ld, dtype=torch.float64, device="cuda")
    t2 = torch.zeros(ld, dtype=torch.float64, device="cuda")
    bk.symv(Kinv, z, t1, N)
    bk.symv(K, t1, t2, N)
    err = float((t2[:N] - z[:N]).norm() / z[:N].norm())
    assert err < 1e-7, err
    assert float((Kinv[:, :N] - Kinv[:, :N].T).abs().max()) == 0.0          # mirrored store: exactly symmetric
    del Kinv, K

    # (2) gradient vs central finite difference of the LML along a random direction in log10(theta)
    u = rng.standard_normal(d)
    u /= np.linalg.norm(u)
    h = 1e-4
    lp = GP.calc_lkd_all(GP.make_hp_class(theta=th * 10 ** (h * u)))[0].ln_lkd
    lm = GP.calc_lkd_all(GP.make_hp_class(theta=th * 10 ** (-h * u)))[0].ln_lkd
    fd = (lp - lm) / (2 * h)
    an = float(np.dot(info.ln_lkd_grad * th * np.log(10), u))
    assert abs(fd - an) < 1e-5 * max(1.0, abs(an)), (fd, an)

    # (3) bit-reproducible
    info2, _ = GP.calc_lkd_all(hp, calc_grad=True)
    assert info2.ln_lkd == info.ln_lkd and np.array_equal(info2.ln_lkd_grad, info.ln_lkd_grad)

    # (4) posterior interpolates values and gradients at the training points; sigma vanishes there
    GP.set_hpara("set", 1, GP.make_hp_class(theta=th, varK=info.hp_varK, beta=info.hp_beta))
    idx = np.arange(0, n, max(1, n // 12))
    mu, sig, dmu, dsig = GP.eval_model(x[idx], calc_grad=True)[:4]
    rng_f = f.max() - f.min()
    # (not exact: the nugget eta P^2 regularises the interpolation)
    assert np.max(np.abs(mu - f[idx])) < 1e-3 * rng_f
    assert np.max(np.abs(dmu - g[idx])) < 1e-2 * np.max(np.abs(g))
    assert np.max(sig) < 5e-2 * np.sqrt(info.hp_varK)
    xs = rng.uniform(-2, 2, (64, d))
    mu2, sig2 = GP.eval_model(xs)[:2]
    assert np.all(sig2 > 0) and np.all(np.isfinite(mu2))
    bk.free_workspace()
    torch.cuda.empty_cache()
