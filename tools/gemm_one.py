"""One large DMMA GEMM launch (target of `ncu --set full`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpgradpy_b200 import backend as bk
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A = torch.randn((M, M), dtype=torch.float64, device="cuda")
B = torch.randn((M, M), dtype=torch.float64, device="cuda")
C = torch.zeros((M, M), dtype=torch.float64, device="cuda")
for _ in range(2):
    bk.dgemm(A, B, C, transb=True)
torch.cuda.synchronize()
print("ok")
