"""A few LML(+gradient) evaluations of one configuration (target of ncu launch lists / captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O
n, d = int(sys.argv[1]), int(sys.argv[2])
grad = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1
x, f, g = O.synthetic_problem(n, d, 0); th = O.bench_theta(d); eta = O.nugget(n, d, "precon")[1]
y = O.make_data_vec(f, g)
X = bk.to_dev(x); Y = bk.to_dev(y)
TH = bk.to_dev(np.tile(th[None, :], (B, 1)) * (1 + 0.001 * np.arange(B)[:, None]))
for _ in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    out, _ = bk.lml_eval(X, Y, TH, mode=L.MODE_PRECON, eta=eta, want_grad=grad)
    e1.record(); torch.cuda.synchronize()
    print(f"n={n} d={d} B={B} grad={grad}: {e0.elapsed_time(e1):.3f} ms  lml[0]={out[0,0].item():.12e}", flush=True)
