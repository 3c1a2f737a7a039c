"""Measure library FP64 peaks on the box (denominators only; not a product path)."""
import time, json, torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
res = {}
def ev(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b); ms = ev(lambda: torch.matmul(a, b))
    res[f"dgemm_{n}_tflops"] = 2 * n**3 / ms * 1e-9
    print(f"cuBLAS DGEMM n={n}: {ms:.2f} ms {res[f'dgemm_{n}_tflops']:.2f} TFLOP/s", flush=True)
    del a, b
# sustained DGEMM
n = 8192; a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
torch.cuda.synchronize(); t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    for _ in range(10): torch.matmul(a, b)
    k += 10; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
res["dgemm_8192_sustained_tflops"] = 2 * n**3 * k / ms * 1e-9
print(f"cuBLAS DGEMM sustained: {res['dgemm_8192_sustained_tflops']:.2f} TFLOP/s", flush=True)
del a, b
for n in (5500, 21000):
    x = torch.randn(n, n + 8, dtype=torch.float64, device=dev); k = x @ x.T + n * torch.eye(n, dtype=torch.float64, device=dev); del x
    torch.linalg.cholesky(k); ms = ev(lambda: torch.linalg.cholesky(k), reps=2)
    res[f"cusolver_potrf_{n}_tflops"] = n**3 / 3 / ms * 1e-9
    print(f"cuSOLVER potrf n={n}: {ms:.2f} ms {res[f'cusolver_potrf_{n}_tflops']:.2f} TFLOP/s", flush=True)
    del k
json.dump(res, open("gpurun_out/fp64_peaks.json", "w"), indent=1)
