import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cProfile, pstats
import bench
from gpgradpy_b200 import backend as bk, _lib as L
from gpgradpy_b200.gp import GaussianProcess
n, d = 500, 10
x, f, g, theta = bench.make_problem(n, d)
GP = GaussianProcess(d, True, "SqExp", "precon")
GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
def step(s):
    GP._dev_ready = False
    return GP.calc_lkd_all(GP.make_hp_class(theta=bench.step_theta(theta, s, 0)), calc_grad=True)
for s in range(3): step(s)
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(10): step(s)
torch.cuda.synchronize()
print("e2e ms/step", (time.perf_counter() - t0) / 10 * 1e3)
def tm(name, fn, k=20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); print(f"{name:30s} {(time.perf_counter()-t)/k*1e3:8.3f} ms")
tm("mem_get_info", torch.cuda.mem_get_info)
tm("to_dev theta", lambda: bk.to_dev(theta[None, :]))
tm("torch.empty out", lambda: torch.empty((1, 19), dtype=torch.float64, device="cuda"))
X, Y, TH = bk.to_dev(x), bk.to_dev(GP.make_data_vec(f, g)), bk.to_dev(theta[None, :])
tm("lml_eval device", lambda: bk.lml_eval(X, Y, TH, mode=L.MODE_PRECON, eta=GP._etaK, want_grad=True))
tm("lml_eval + cpu()", lambda: bk.lml_eval(X, Y, TH, mode=L.MODE_PRECON, eta=GP._etaK, want_grad=True)[0].cpu())
pr = cProfile.Profile(); pr.enable()
for s in range(10): step(s)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
