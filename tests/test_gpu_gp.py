"""GPU parity of the GP hot path through the C ABI: covariance build, LML + gradient, posterior.

Checked against (a) the reference's own outputs stored in tests/golden (made by oracle/make_golden.py) and
(b) the CPU oracle (oracle/gegp_oracle.py) on seeded inputs.  Tolerances follow BASELINE.json north_star:
covariance entries 1e-12 relative (to the block scale where the closed form cancels), LML / gradient /
posterior 1e-8 relative.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MODES = {"base": 0, "precon": 1, "rescale_origin": 0}


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _block_scale_err(Kg, Kr):
    """max |dK| / max(|Kr|, row/col scale): entries where (2 th - 4 th^2 r^2) cancels are compared to the block scale."""
    s = np.sqrt(np.abs(np.diag(Kr)))
    scale = np.maximum(np.abs(Kr), 1e-3 * s[:, None] * s[None, :])
    return float(np.max(np.abs(Kg - Kr) / scale))


MAT_CASES = ["c1_d2_n20_precon", "c1_d2_n20_base", "d2_n20_wide_precon", "d4_n37_precon", "d3_n30_base",
             "d3_n25_rescale_origin", "d1_n9_precon", "d3_n18_mask_prefix", "d2_n12_mask_scatter"]


@pytest.mark.parametrize("name", MAT_CASES)
def test_build_cov_vs_reference(golden_dir, name):
    from gpgradpy_b200 import backend as bk, _lib as L
    g = _load(golden_dir, name)
    x = g["x_scl"]
    n, d = x.shape
    mask = g["mask"] if g["mask"].size else None
    slot, ng = bk.slot_from_mask(mask, n)
    mode = str(g["mode"])
    eta = float(g["eta"])
    Kern, _ = bk.build_cov(x, g["theta"], n_g=ng, slot=slot, mode=L.MODE_BASE, eta=0.0)
    e = _block_scale_err(Kern.cpu().numpy(), g["Kern"])
    print(name, "Kern err", e)
    assert e < 1e-12
    if mode == "precon":
        Kcor_eta, p = bk.build_cov(x, g["theta"], n_g=ng, slot=slot, mode=L.MODE_PRECON, eta=eta)
        N = g["Kern"].shape[0]
        ref = g["Kcor"] + eta * np.eye(N)
        e = _block_scale_err(Kcor_eta.cpu().numpy(), ref)
        print(name, "Kcor+eta err", e)
        assert e < 1e-12
        pr = np.sqrt(np.diag(g["Kern"]))
        assert _rel(p[:N].cpu().numpy(), pr) < 1e-14
        Kcov, _ = bk.build_cov(x, g["theta"], n_g=ng, slot=slot, mode=L.MODE_PRECON_COV, eta=eta)
        e = _block_scale_err(Kcov.cpu().numpy(), g["Kcov"])
        print(name, "Kcov err", e)
        assert e < 1e-12
        # lower-only variant writes the same lower triangle
        Kl, _ = bk.build_cov(x, g["theta"], n_g=ng, slot=slot, mode=L.MODE_PRECON, eta=eta, uplo=1)
        assert np.array_equal(np.tril(Kl.cpu().numpy()), np.tril(Kcor_eta.cpu().numpy()))
    else:
        Kcov, _ = bk.build_cov(x, g["theta"], n_g=ng, slot=slot, mode=L.MODE_BASE, eta=eta)
        e = _block_scale_err(Kcov.cpu().numpy(), g["Kcov"])
        print(name, "Kcov err", e)
        assert e < 1e-12


@pytest.mark.parametrize("name", ["d3_n24_noisy_precon", "d3_n24_noisy_base"])
def test_build_cov_noisy(golden_dir, name):
    from gpgradpy_b200 import backend as bk, _lib as L
    g = _load(golden_dir, name)
    x = g["x"]
    mode = str(g["mode"])
    varK, eta = float(g["varK"]), float(g["eta"])
    m = L.MODE_PRECON_COV if mode == "precon" else L.MODE_BASE
    Kcov, _ = bk.build_cov(x, g["theta"], noise=g["noise_vec"] / varK, mode=m, eta=eta, varK=varK)
    e = _block_scale_err(Kcov.cpu().numpy(), g["Kcov"])
    print(name, "noisy Kcov err", e)
    assert e < 1e-12


LML_CASES = ["c1_d2_n20_precon", "c1_d2_n20_base", "c1_d2_n20_rescale_origin", "d2_n20_wide_precon", "d4_n37_precon",
             "d3_n30_base", "d3_n25_rescale_origin", "d5_n64_precon_seed1", "d1_n9_precon", "d3_n18_mask_prefix",
             "c4_d5_n200_precon", "c2_d10_n500_precon"]


@pytest.mark.parametrize("name", LML_CASES)
def test_lml_grad_vs_reference(golden_dir, name):
    from gpgradpy_b200 import backend as bk, _lib as L
    g = _load(golden_dir, name)
    x, f, gr = g["x_scl"], g["fval_scl"], g["grad_scl"]
    n, d = x.shape
    mask = g["mask"] if g["mask"].size else None
    slot, ng = bk.slot_from_mask(mask, n)
    y = np.hstack((f, gr.reshape(gr.size, order="F")))
    mode = L.MODE_PRECON if str(g["mode"]) == "precon" else L.MODE_BASE
    out, alpha = bk.lml_eval(x, y, g["theta"][None, :], n_g=ng, slot=slot, mode=mode, eta=float(g["eta"]),
                             want_grad=True, want_alpha=True)
    o = out.cpu().numpy()[0]
    assert o[L.OUT_INFO] == 0
    e_lml = _rel(o[L.OUT_LML], g["ln_lkd"])
    e_vk = _rel(o[L.OUT_SIGMA2], g["hp_varK"])
    e_beta = _rel(o[L.OUT_BETA], g["hp_beta"][0])
    e_ld = _rel(o[L.OUT_LOGDET], g["ln_det"])
    gg = o[L.OUT_GRAD:L.OUT_GRAD + d]
    e_grad = float(np.max(np.abs(gg - g["ln_lkd_grad"])) / np.max(np.abs(g["ln_lkd_grad"])))
    print(f"{name}: lml {e_lml:.2e} varK {e_vk:.2e} beta {e_beta:.2e} logdet {e_ld:.2e} grad {e_grad:.2e}")
    # north_star tolerance: 1e-8 relative.  Config 1 clusters 20 points in [0.9,1.1]^2 so cond(K) sits at the
    # 1e10 target and two CPU implementations already differ by 3e-8 (tests/golden/golden_report.json,
    # tests/test_oracle_golden.py); those cases are held to cond*eps ~ 1e-6.
    tol = 1e-6 if name.startswith("c1_") else 1e-8
    assert e_lml < 1e-8 and e_ld < 1e-8 and e_vk < tol and e_grad < tol
    assert e_beta < 1e-7   # beta is a ratio of two ill-conditioned sums; the reference itself moves ~1e-9


@pytest.mark.parametrize("name", ["d3_n24_noisy_precon", "d3_n24_noisy_base"])
def test_lml_noisy_vs_reference(golden_dir, name):
    from gpgradpy_b200 import backend as bk, _lib as L
    g = _load(golden_dir, name)
    x, f, gr = g["x"], g["fval"], g["grad"]
    n, d = x.shape
    y = np.hstack((f, gr.reshape(gr.size, order="F")))
    mode = L.MODE_PRECON if str(g["mode"]) == "precon" else L.MODE_BASE
    out, _ = bk.lml_eval(x, y, g["theta"][None, :], mode=mode, eta=float(g["eta"]), noise=g["noise_vec"],
                         varK_batch=np.array([float(g["varK"])]), want_grad=True)
    o = out.cpu().numpy()[0]
    ref_grad = g["ln_lkd_grad"]          # [theta.., varK]
    got = np.hstack((o[L.OUT_GRAD:L.OUT_GRAD + d], o[L.OUT_DVARK]))
    e_lml = _rel(o[L.OUT_LML], g["ln_lkd"])
    e_grad = float(np.max(np.abs(got - ref_grad) / np.maximum(np.abs(ref_grad), 1e-3 * np.max(np.abs(ref_grad)))))
    print(f"{name}: lml {e_lml:.2e} grad {e_grad:.2e} beta {_rel(o[L.OUT_BETA], g['hp_beta'][0]):.2e}")
    assert e_lml < 1e-8 and e_grad < 1e-8


PRED_CASES = ["c1_d2_n20_precon", "c1_d2_n20_base", "d2_n20_wide_precon", "d4_n37_precon", "d3_n30_base",
              "d5_n64_precon_seed1", "d1_n9_precon", "c4_d5_n200_precon", "c2_d10_n500_precon"]


@pytest.mark.parametrize("name", PRED_CASES)
def test_predict_vs_reference(golden_dir, name):
    from gpgradpy_b200 import backend as bk, _lib as L
    g = _load(golden_dir, name)
    x, f, gr = g["x_scl"], g["fval_scl"], g["grad_scl"]
    y = np.hstack((f, gr.reshape(gr.size, order="F")))
    mode = L.MODE_PRECON if str(g["mode"]) == "precon" else L.MODE_BASE
    varK, beta = float(g["hp_varK"]), float(g["hp_beta"][0])
    st = bk.predict_setup(x, y, g["theta"], beta, mode=mode, eta=float(g["eta"]))
    mu, sig, sig2, nneg = bk.predict(st, g["x_test"], varK)
    assert int(st.info.item()) == 0
    mu, sig, sig2 = mu.cpu().numpy(), sig.cpu().numpy(), sig2.cpu().numpy()
    e_mu = float(np.max(np.abs(mu - g["mu"])) / np.max(np.abs(g["mu"])))
    # sigma: |d sig^2| <= 1e-8 varK, and relative where sig^2/varK is not cancellation noise
    ds2 = np.abs(sig ** 2 - g["sig"] ** 2) / varK
    big = (g["sig"] ** 2 / varK) > 1e-6
    e_sig_rel = float(np.max(np.abs(sig[big] - g["sig"][big]) / g["sig"][big])) if big.any() else 0.0
    print(f"{name}: mu {e_mu:.2e} dsig2/varK {ds2.max():.2e} sig rel {e_sig_rel:.2e} nneg {int(nneg.item())}")
    assert e_mu < 1e-8 and ds2.max() < 1e-8 and e_sig_rel < 1e-6


def test_candidate_batch_vs_reference(golden_dir):
    from gpgradpy_b200 import backend as bk, _lib as L
    g = _load(golden_dir, "c4_d5_n200_cand8")
    x, f, gr = g["x"], g["fval"], g["grad"]
    n, d = x.shape
    y = np.hstack((f, gr.reshape(gr.size, order="F")))
    th = 10.0 ** g["log10_theta"]
    out, _ = bk.lml_eval(x, y, th, mode=L.MODE_PRECON, eta=float(g["eta"]), want_grad=True)
    o = out.cpu().numpy()
    e_lml = _rel(o[:, L.OUT_LML], g["ln_lkd"])
    gscale = np.max(np.abs(g["ln_lkd_grad"]), axis=1, keepdims=True)
    e_grad = float(np.max(np.abs(o[:, L.OUT_GRAD:L.OUT_GRAD + d] - g["ln_lkd_grad"]) / gscale))
    print(f"batch: lml {e_lml:.2e} grad {e_grad:.2e}")
    assert np.all(o[:, L.OUT_INFO] == 0)
    assert e_lml < 1e-8 and e_grad < 1e-7
    assert int(np.argmax(o[:, L.OUT_LML])) == int(np.nanargmax(g["ln_lkd"]))
    # batched == one-at-a-time, bit for bit (deterministic reductions)
    o1 = np.vstack([bk.lml_eval(x, y, th[i:i + 1], mode=L.MODE_PRECON, eta=float(g["eta"]), want_grad=True)[0].cpu().numpy()
                    for i in range(th.shape[0])])
    assert np.array_equal(o1, o)


def test_lml_vs_oracle_medium():
    """Seeded d=6, n=150 (N=1050) against the CPU oracle (both forms)."""
    from gpgradpy_b200 import backend as bk, _lib as L
    from oracle import gegp_oracle as O
    x, f, gr = O.synthetic_problem(150, 6, 3)
    th = O.bench_theta(6)
    eta = O.nugget(150, 6, "precon")[1]
    ref = O.lkd_wo_noise_lean(x, f, gr, th, "precon", eta)
    y = O.make_data_vec(f, gr)
    out, alpha = bk.lml_eval(x, y, th[None, :], mode=L.MODE_PRECON, eta=eta, want_grad=True, want_alpha=True)
    o = out.cpu().numpy()[0]
    e_lml = _rel(o[L.OUT_LML], ref.ln_lkd)
    e_grad = float(np.max(np.abs(o[L.OUT_GRAD:] - ref.ln_lkd_grad)) / np.max(np.abs(ref.ln_lkd_grad)))
    e_alpha = float(np.max(np.abs(alpha.cpu().numpy()[0] - ref.alpha)) / np.max(np.abs(ref.alpha)))
    print(f"medium: lml {e_lml:.2e} grad {e_grad:.2e} alpha {e_alpha:.2e}")
    assert e_lml < 1e-8 and e_grad < 1e-8 and e_alpha < 1e-6


def test_cuda_graph_replay_matches_stream_path():
    """The captured-graph path used by the optimiser loop gives bit-identical results, also after the data buffers
    are refreshed in place and for several candidates in a row."""
    import torch
    from gpgradpy_b200 import backend as bk, _lib as L
    from oracle import gegp_oracle as O
    x, f, gr = O.synthetic_problem(90, 4, 5)
    y = O.make_data_vec(f, gr)
    eta = O.nugget(90, 4, "precon")[1]
    X, Y = bk.to_dev(x), bk.to_dev(y)
    th0 = O.bench_theta(4)
    for k in range(3):
        th = th0 * (1.0 + 0.1 * k)
        a = bk.lml_eval(X, Y, th[None, :], mode=L.MODE_PRECON, eta=eta, want_grad=True)[0].cpu().numpy()
        b = bk.lml_eval_graphed(X, Y, th[None, :], mode=L.MODE_PRECON, eta=eta, want_grad=True).cpu().numpy()
        assert np.array_equal(a, b)
    assert bk.replay_stats["replays"] >= 3
    # new data written into the same buffers: the graph must see it
    x2, f2, gr2 = O.synthetic_problem(90, 4, 6)
    X.copy_(torch.as_tensor(x2)); Y.copy_(torch.as_tensor(O.make_data_vec(f2, gr2)))
    a = bk.lml_eval(X, Y, th0[None, :], mode=L.MODE_PRECON, eta=eta, want_grad=True)[0].cpu().numpy()
    b = bk.lml_eval_graphed(X, Y, th0[None, :], mode=L.MODE_PRECON, eta=eta, want_grad=True).cpu().numpy()
    assert np.array_equal(a, b)
    ref = O.lkd_wo_noise(x2, f2, gr2, th0, "precon", eta)
    assert abs(b[0, L.OUT_LML] - ref.ln_lkd) < 1e-8 * abs(ref.ln_lkd)
