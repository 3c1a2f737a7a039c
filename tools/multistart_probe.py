"""Batch point B: a 5-start ('lhs') fit, sequential loop vs lock-step batched multi-start (same optimum)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200.gp import GaussianProcess
from oracle import gegp_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
d = int(sys.argv[2]) if len(sys.argv) > 2 else 10
x, f, g = O.synthetic_problem(n, d, 0)
res = {}
for lock in (False, True, False, True):
    GP = GaussianProcess(d, True, "SqExp", "precon")
    GP.lkd_optz_start_mtd = "lhs"
    GP.lockstep_multistart = lock
    GP.init_optz_surr(3)
    GP.set_data(x[:1], f[:1], np.zeros(1), g[:1], np.zeros((1, d)))
    GP.set_hpara("optz", 0)
    GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    GP.set_hpara("optz", 1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = getattr(GP, "_lockstep_stats", None)
    res[lock] = GP.hp_vals.theta.copy()
    print(f"n={n} d={d} N={n*(d+1)} 5-start fit, lockstep={lock}: {dt:.2f} s, mean iters {GP.hp_optz_iter_mean[1]:.1f}"
          + (f", {st['n_evals']} evaluations in {st['n_batches']} device batches" if (lock and st) else ""), flush=True)
print("same optimum:", np.array_equal(res[True], res[False]))
