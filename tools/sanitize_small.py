"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): one LML + gradient evaluation (N = 400: four leaves,
look-ahead streams, the cluster chain step), the direct-form terms, a batched scan and a posterior with x-gradients and
Hessians, for the three kernel families.  Usage (GPU box):
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200.gp import GaussianProcess
from oracle import gegp_oracle as O

n, d = 100, 3
x, f, g = O.synthetic_problem(n, d, 0)
th = O.bench_theta(d) * 3
xs = np.random.default_rng(1).uniform(-2, 2, (9, d))
for kname, khp in (("SqExp", None), ("Ma5f2", None), ("RatQu", 1.5)):
    GP = GaussianProcess(d, True, kname, "precon")
    GP.use_cuda_graphs = False
    GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th, kernel=khp), calc_grad=True, lkd_use_adj_mtd=False)
    ref = O.lkd_wo_noise(x, f, g, th, "precon", GP._etaK, kernel=(kname, khp))
    e = abs(info.ln_lkd - ref.ln_lkd) / abs(ref.ln_lkd)
    eg = np.max(np.abs(info.ln_lkd_grad - ref.ln_lkd_grad)) / np.max(np.abs(ref.ln_lkd_grad))
    rows = np.log10(np.hstack((np.tile(th, (6, 1)) * np.linspace(0.5, 2, 6)[:, None],) + ((np.full((6, 1), khp),) if khp else ())))
    tab = GP.calc_lkd_batch(rows, calc_grad=True)
    GP.set_hpara("set", 1, GP.make_hp_class(theta=th, kernel=khp, varK=info.hp_varK, beta=info.hp_beta))
    mu, sig, dmu, dsig = GP.eval_model(xs, calc_grad=True)[:4]
    h = GP.eval_model(xs[:1], calc_grad=True, calc_hess=True)
    torch.cuda.synchronize()
    print(f"{kname}: ok={ok} lml rel err {e:.1e} grad {eg:.1e} scan ok {int((tab[:, 4] == 0).sum())}/6 mu[0]={mu[0]:.6f}", flush=True)
    assert ok and e < 1e-8 and eg < 1e-8
# base mode with the condition-number constraint quantities (Lanczos kernels)
GP = GaussianProcess(d, True, "SqExp", "base")
GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th), calc_grad=True, calc_cond=True)
torch.cuda.synchronize()
print(f"base: ok={ok} cond={info.cond:.3e}", flush=True)
print("SANITIZE_RUN_DONE")
