import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O
n, d = 200, 5
x, f, g = O.synthetic_problem(n, d, 0); y = O.make_data_vec(f, g); eta = O.nugget(n, d, "precon")[1]
B = 64
cand = 10.0 ** np.random.default_rng(0).uniform(-5, 1, (B, d))
X, Y, C = bk.to_dev(x), bk.to_dev(y), bk.to_dev(cand)
for grad in (False, True):
    t = bk.lml_eval(X, Y, C, mode=L.MODE_PRECON, eta=eta, want_grad=grad)[0].cpu().numpy()
    t2 = bk.lml_eval(X, Y, C, mode=L.MODE_PRECON, eta=eta, want_grad=grad)[0].cpu().numpy()
    one = np.vstack([bk.lml_eval(X, Y, C[i:i + 1], mode=L.MODE_PRECON, eta=eta, want_grad=grad)[0].cpu().numpy() for i in range(B)])
    one2 = np.vstack([bk.lml_eval(X, Y, C[i:i + 1], mode=L.MODE_PRECON, eta=eta, want_grad=grad)[0].cpu().numpy() for i in range(B)])
    nc = 9 if not grad else 9 + d
    print("grad", grad, "batch repeat equal", np.array_equal(t[:, :nc], t2[:, :nc]), "single repeat equal", np.array_equal(one[:, :nc], one2[:, :nc]),
          "batch==single", np.array_equal(t[:, :nc], one[:, :nc]))
    dif = np.abs(t[:, :nc] - one[:, :nc])
    print("   max abs diff per column", dif.max(axis=0))
    print("   rows differing", np.where(dif.max(axis=1) > 0)[0][:20])
