"""Phase clocks of potf2_inv_kernel (needs tools/libgegp_dbg.so built with -DGEGP_LEAF_CLOCKS)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgegp_dbg.so")
from gpgradpy_b200 import backend as bk
lib = _lib.load()
N = 128
rng = np.random.default_rng(0)
G = rng.standard_normal((N, N + 8)); K = G @ G.T / (N + 8) + 0.5 * np.eye(N)
ld = bk.ld_of(N)
Kt = torch.as_tensor(np.tril(K)).cuda()
A = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
dinv = bk.dinv_buffer(N)
for it in range(3):
    A[:, :N] = Kt
    torch.cuda.synchronize()
    bk.potrf(A, N, 0, dinv)
    torch.cuda.synchronize()
    out = (ctypes.c_longlong * 16)()
    lib.gegp_debug_leaf_clocks(out)
    c = list(out)
    names = ["load", "panel0", "upd0", "panel1", "upd1", "panel2", "upd2", "panel3", "upd3(none)", "tail"]
    idx = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
    print("iter", it, "total cycles", c[10] - c[0])
    for i, nm in enumerate(names):
        print(f"   {nm:12s} {c[idx[i + 1]] - c[idx[i]]:8d}")
