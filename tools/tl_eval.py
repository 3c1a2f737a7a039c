"""Debug: per-launch timeline of one LML+gradient evaluation (GEGP_TIMELINE=<file> must be set in the environment).
Prints, beside the dump, per-label totals and the intervals in which no GEMM was running."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O
n, d = int(sys.argv[1]), int(sys.argv[2])
grad = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
x, f, g = O.synthetic_problem(n, d, 0); th = O.bench_theta(d); eta = O.nugget(n, d, "precon")[1]
y = O.make_data_vec(f, g)
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1
X = bk.to_dev(x); Y = bk.to_dev(y); TH = bk.to_dev(np.tile(th[None, :], (B, 1)) * (1 + 0.001 * np.arange(B)[:, None]))
for it in range(3):
    torch.cuda.synchronize()
    if it == 2:
        L.profile_begin(False)
    out, _ = bk.lml_eval(X, Y, TH, mode=L.MODE_PRECON, eta=eta, want_grad=grad)
    torch.cuda.synchronize()
L.profile_end()
rows = []
for ln in open(os.environ["GEGP_TIMELINE"]):
    p = ln.split()
    rows.append((p[0], p[1], float(p[2]), float(p[3])))
ib = max(i for i, r in enumerate(rows) if r[0].startswith("build"))   # the marks accumulate over the warm-up calls too
rows = rows[ib:]
t0 = rows[0][2]
rows = [(l, s, a - t0, b - t0) for l, s, a, b in rows]
t_end = max(r[3] for r in rows)
tot = {}
for lab, st, a, b in rows:
    k = lab.split(":")[0]
    mf = float(lab.split(":")[4]) if lab.count(":") >= 4 else 0.0
    tot.setdefault(k, [0, 0.0, 0.0]); tot[k][0] += 1; tot[k][1] += b - a; tot[k][2] += mf
print(f"span {t_end:.1f} us, {len(rows)} launches")
for k, (c, t, mf) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:10s} {c:5d} launches  {t:9.1f} us summed  {mf*1e-3:9.2f} GFLOP  {mf/t if t > 0 else 0:6.2f} TFLOP/s")
chain = [r[3] for r in rows if r[0].startswith(("potf2", "cprep"))]
if chain:
    print(f"chain ends at {max(chain):.1f} us")
if len(sys.argv) > 4:   # per-launch table, longest first
    for lab, st, a, b in sorted(rows, key=lambda r: -(r[3] - r[2]))[:int(sys.argv[4])]:
        mf = float(lab.split(":")[4]) if lab.count(":") >= 4 else 0.0
        print(f"  {lab:40s} {a:9.1f} {b - a:8.1f} us  {mf/(b-a) if b > a else 0:6.2f} TFLOP/s")
