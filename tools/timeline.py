"""Debug: dump the per-launch timeline of one potrf (GEGP_TIMELINE=<file> must be set in the environment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O
n, d = int(sys.argv[1]), int(sys.argv[2])
N = n * (d + 1)
x, f, g = O.synthetic_problem(n, d, 0); th = O.bench_theta(d); eta = O.nugget(n, d, "precon")[1]
X = bk.to_dev(x); TH = bk.to_dev(th)
ld = bk.ld_of(N)
buf = torch.empty((N + 2, ld), dtype=torch.float64, device="cuda")
dinv = bk.dinv_buffer(N)
for it in range(3):
    bk.build_cov(X, TH, mode=L.MODE_PRECON, eta=eta, out=buf[:N], uplo=1)
    torch.cuda.synchronize()
    if it == 2:
        L.profile_begin(False)
    bk.potrf(buf, N, 0, dinv)
    torch.cuda.synchronize()
L.profile_end()
