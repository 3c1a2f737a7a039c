import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk
np.set_printoptions(precision=3, linewidth=250)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(0)
G = rng.standard_normal((N, N + 8)); K = G @ G.T / (N + 8) + 0.5 * np.eye(N)
ld = bk.ld_of(N)
A = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
A[:, :N] = torch.as_tensor(np.tril(K)).cuda()
info, dinv = bk.potrf(A, N, 0)
torch.cuda.synchronize()
print("info", int(info.item()))
Lg = np.tril(A[:, :N].cpu().numpy()); Lr = np.linalg.cholesky(K)
E = np.abs(Lg - Lr)
print("max err", E.max())
bad = np.argwhere(E > 1e-10)
print("n bad", len(bad), "first bad", bad[:10].tolist())
print("rows with errors:", sorted(set(bad[:, 0].tolist()))[:20], "cols:", sorted(set(bad[:, 1].tolist()))[:40])
print("Lg[32:36,:6]\n", Lg[32:36, :6], "\nLr[32:36,:6]\n", Lr[32:36, :6])
