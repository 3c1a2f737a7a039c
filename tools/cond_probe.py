"""Cost of the device condition number (Lanczos) next to the LML+gradient evaluation it rides on."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200.gp import GaussianProcess
from gpgradpy_b200 import backend as bk
from oracle import gegp_oracle as O
for n, d in [(200, 5), (500, 10)]:
    x, f, g = O.synthetic_problem(n, d, 0)
    GP = GaussianProcess(d, True, "SqExp", "base")
    GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
    th = O.bench_theta(d)
    for rep in range(3):
        hp = GP.make_hp_class(theta=th * (1 + 0.01 * rep))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        i0, _ = GP.calc_lkd_all(hp, calc_grad=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        i1, _ = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=True)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        lc = GP._last_cond
        print(f"n={n} d={d} N={n*(d+1)}: lml+grad {1e3*(t1-t0):.1f} ms; with cond+grad {1e3*(t2-t1):.1f} ms; cond {i1.cond:.4e} "
              f"resid {lc['resid_max']:.1e}/{lc['resid_min']:.1e} cycles {lc.get('cycles')}", flush=True)
