"""Debug: is a candidate evaluated in a batch of B bit-identical to the same candidate evaluated alone, for small B?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O
n, d = int(sys.argv[1]), int(sys.argv[2])
x, f, g = O.synthetic_problem(n, d, 0)
y = O.make_data_vec(f, g)
eta = O.nugget(n, d, "precon")[1]
X, Y = bk.to_dev(x), bk.to_dev(y)
rng = np.random.default_rng(3)
TH = 10.0 ** rng.uniform(-3, -1, (6, d))
single = np.vstack([bk.lml_eval(X, Y, TH[i:i + 1], mode=L.MODE_PRECON, eta=eta, want_grad=True)[0].cpu().numpy() for i in range(6)])
graphed = np.vstack([bk.lml_eval_graphed(X, Y, bk.to_dev(TH[i:i + 1]), mode=L.MODE_PRECON, eta=eta, want_grad=True).cpu().numpy() for i in range(6)])
print("graph == eager single:", np.array_equal(single, graphed))
for B in (2, 3, 4, 5, 6):
    out = bk.lml_eval(X, Y, TH[:B], mode=L.MODE_PRECON, eta=eta, want_grad=True)[0].cpu().numpy()
    same = np.array_equal(out, single[:B])
    print(f"B={B}: identical to singles: {same}", "" if same else f"max rel diff {np.max(np.abs(out - single[:B]) / np.maximum(1e-300, np.abs(single[:B]))):.2e} cols {np.unique(np.argwhere(out != single[:B])[:, 1])}")
