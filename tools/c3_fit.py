"""BASELINE configs[2]: d=20, n=1000 (N=21000) preconditioned GE-GP hyper-parameter fit through the public API
(history protocol of SURVEY appendix B.12: one-point call first), then posterior mean / std at 10k test points.

    python tools/c3_fit.py [n] [d] [nx]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200.gp import GaussianProcess
from gpgradpy_b200 import backend as bk
from oracle import gegp_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nx = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
x, f, g = O.synthetic_problem(n, d, 0)
GP = GaussianProcess(d, True, "SqExp", "precon")
GP.init_optz_surr(3)
GP.set_data(x[:1], f[:1], np.zeros(1), g[:1], np.zeros((1, d)))
GP.set_hpara("optz", 0)
GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
n_eval = [0]
orig = GP._eval_rows
def counted(theta_rows, **kw):
    n_eval[0] += int(np.prod(tuple(theta_rows.shape))) // d
    return orig(theta_rows, **kw)
GP._eval_rows = counted
torch.cuda.synchronize(); t0 = time.perf_counter()
GP.set_hpara("optz", 1)
torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
hp = GP.hp_vals
info, ok = GP.calc_lkd_all(hp, calc_grad=True)
glog = info.ln_lkd_grad * hp.theta * np.log(10)
print(f"c3 fit n={n} d={d} N={n*(d+1)}: {t_fit:.1f} s, {n_eval[0]} LML evaluations "
      f"(pick x0 {GP.time_pick_hp0_all[1]:.1f} s, SLSQP {GP.time_hp_optz_all[1]:.1f} s, iters {GP.hp_optz_iter_mean[1]:.0f}, "
      f"success {GP.hp_optz_success[1]:.0f}); LML {info.ln_lkd:.6f}, varK {hp.varK:.4e}, "
      f"max |dLML/dlog10 theta| {np.max(np.abs(glog)):.3e}, theta range [{hp.theta.min():.3e}, {hp.theta.max():.3e}]", flush=True)
xs = np.random.default_rng(1).uniform(-2, 2, (nx, d))
torch.cuda.synchronize(); t0 = time.perf_counter()
mu, sig = GP.eval_model(xs)[:2]
torch.cuda.synchronize(); t_pred = time.perf_counter() - t0
ft = O.rosenbrock(xs)
print(f"c3 predict nx={nx} through eval_model (host in, host out): {t_pred*1e3:.1f} ms; "
      f"rmse/range {np.sqrt(np.mean((mu-ft)**2))/(ft.max()-ft.min()):.3e}; "
      f"fraction |err| < 3 sig: {np.mean(np.abs(mu-ft) < 3*sig + 1e-12):.3f}", flush=True)
