#!/usr/bin/env python
"""bench.py -- LML + hyper-parameter-gradient evaluations per second of the gradient-enhanced GP hot path.

    python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's CUDA path (one JSON line)
    python bench.py --impl reference --steps 2 --warmup 1     # the reference algorithm on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # N independent candidate streams (weak scaling)

A "step" is one evaluation of LML and its full theta-gradient (build -> Cholesky -> solves -> inverse ->
gradient contraction) on the BASELINE.json configuration `configs[1]` (d=10, n=500, N=5500, preconditioned).
`value` times the device path with inputs resident in HBM; `e2e` goes through the public
GaussianProcess.calc_lkd_all API with host buffers (H2D of X, y, theta and D2H of the result every step).
The north-star size (d=20, n=1000, N=21000) is measured in the same run and reported under "c3".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "LML+gradient evals/sec (fp64, preconditioned GE-GP)"
UNIT = "evals/s"
WORKLOADS = {"c2": (500, 10), "c3": (1000, 20), "c1": (20, 2), "c4": (200, 5)}
FP64_DMMA_PEAK_TFLOPS = 37.13   # measured on this pool's B200: profiles/r01/dmma_peak.log (DMMA.8x8x4 issue peak)
FP64_DGEMM_TFLOPS = 36.17       # cuBLAS DGEMM 16384^3 on the same box: profiles/r01/cublas_peak.log


# ----------------------------------------------------------------------------------------------- synthetic inputs
def rosenbrock(x, a=10.0):
    if x.shape[1] == 1:
        return np.sin(3.0 * x[:, 0]) + x[:, 0] ** 2
    return np.sum(a * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1.0 - x[:, :-1]) ** 2, axis=1)


def rosenbrock_grad(x, a=10.0):
    if x.shape[1] == 1:
        return 3.0 * np.cos(3.0 * x) + 2.0 * x
    g = np.zeros_like(x)
    g[:, :-1] += -2.0 * (1.0 - x[:, :-1]) - 4.0 * a * x[:, :-1] * (x[:, 1:] - x[:, :-1] ** 2)
    g[:, 1:] += 2.0 * a * (x[:, 1:] - x[:, :-1] ** 2)
    return g


def make_problem(n, d, seed=0):
    """X ~ U[-2,2]^d (default_rng(seed)), Rosenbrock a=10, theta_i = 0.05 linspace(.5,1.5,d) 10/d."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2.0, 2.0, (n, d))
    theta = 0.05 * np.linspace(0.5, 1.5, d) * (10.0 / d)
    return x, rosenbrock(x), rosenbrock_grad(x), theta


def step_theta(theta, step, rank):
    """A different candidate every step and rank (no result can be cached)."""
    return theta * (1.0 + 0.003 * step + 0.0007 * rank)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm
def _use_all_host_threads():
    """Give the host BLAS every core (returns the thread count actually in effect)."""
    want = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=want)
        got = [p.get("num_threads", 1) for p in threadpoolctl.threadpool_info() if p.get("user_api") == "blas"]
        return int(max(got)) if got else want
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", want))


def run_reference(args):
    """The reference's algorithm on the host cores (oracle port of calc_lkd_all, noise-free precon)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import gegp_oracle as O     # the one place bench.py executes oracle/ as the thing measured
    cores = _use_all_host_threads()          # torchrun exports OMP_NUM_THREADS=1: undo that for the CPU arm
    n, d = WORKLOADS[args.workload]
    x, f, g, theta = make_problem(n, d)
    eta = O.nugget(n, d, "precon")[1]
    steps = max(1, min(args.steps, 3))       # bounded sample: one step is ~15-30 s of CPU work at c2
    xs, fs, gs, ths = make_problem(40, 3)
    for _ in range(max(1, min(args.warmup, 1))):
        O.lkd_wo_noise(xs, fs, gs, ths, "precon", O.nugget(40, 3, "precon")[1])
    t0 = time.perf_counter()
    for s in range(steps):
        O.lkd_wo_noise(x, f, g, step_theta(theta, s, 0), "precon", eta, calc_grad=True)
    dt = time.perf_counter() - t0
    val = steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: d={d}, n={n}, N={n * (d + 1)}, precon, LML+grad", "seed": 0},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{steps} full LML+gradient evaluation(s) of the same workload "
                                       "(oracle/gegp_oracle.lkd_wo_noise: NumPy/SciPy restatement that materialises "
                                       "dK/dtheta and K^-1 like the reference)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- CUDA arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from gpgradpy_b200 import backend as bk, _lib as L
    from gpgradpy_b200.gp import GaussianProcess
    from gpgradpy_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, d = WORKLOADS[args.workload]
    N = n * (d + 1)
    x, f, g, theta = make_problem(n, d)
    GP = GaussianProcess(d, True, "SqExp", "precon")
    GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
    eta = GP._etaK
    y = GP.make_data_vec(f, g)
    X_dev, y_dev = bk.to_dev(x), bk.to_dev(y)
    K, W = args.steps, max(args.warmup, 3)
    thetas = torch.stack([bk.to_dev(step_theta(theta, s, rank)) for s in range(K + W)])

    def device_step(s):
        # the product path of the optimiser's inner loop: one captured CUDA graph per problem, replayed per candidate
        out = bk.lml_eval_graphed(X_dev, y_dev, thetas[s:s + 1], mode=L.MODE_PRECON, eta=eta, want_grad=True)
        if world > 1:
            out = parallel.gather_rows(out, world)   # scalar results of all ranks, one NCCL all_gather
        return out

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(W):
        device_step(s)
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.profile_begin(False)
    replayed0 = bk.replay_stats["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    last = None
    for s in range(W, W + K):
        last = device_step(s)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    launches = L.profile_end()["launches"] + bk.replay_stats["kernel_launches"] - replayed0
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    last = last.clone()
    info_ok = bool((last[:, L.OUT_INFO] == 0).all().item())
    value = world * K / (ms_total * 1e-3)

    # ---- end to end through the public API with host buffers (pinned H2D of X, y, theta; D2H of the result)
    for s in range(2):
        GP._dev_ready = False
        GP.calc_lkd_all(GP.make_hp_class(theta=step_theta(theta, s, rank)), calc_grad=True)
    sync()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        GP._dev_ready = False                    # forces the host -> device copy of X and y again
        info, good = GP.calc_lkd_all(GP.make_hp_class(theta=step_theta(theta, s, rank)), calc_grad=True)
    sync()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = {"value": world * K / float(t_e2e.item()), "unit": UNIT,
           "h2d_bytes_per_step": int(8 * (n * d + N + d)), "d2h_bytes_per_step": int(8 * L.out_len(d)),
           "api": "GaussianProcess.calc_lkd_all(hp_vals, calc_grad=True)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (gemm_f64_kernel, fp64 DMMA): time every launch with CUDA events
    def ev_ms(fn, reps=3):
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    def phase_numbers(n_, d_, reps=2):
        N_ = n_ * (d_ + 1)
        x_, f_, g_, th_ = make_problem(n_, d_)
        Xd, yd, thd = bk.to_dev(x_), bk.to_dev(np.hstack((f_, g_.reshape(g_.size, order="F")))), bk.to_dev(th_[None, :])
        eta_ = GaussianProcess(d_, True).calc_nugget(n_)[1]
        run = lambda grad: bk.lml_eval(Xd, yd, thd, mode=L.MODE_PRECON, eta=eta_, want_grad=grad)  # noqa: E731
        run(True)
        ms_grad, ms_val = ev_ms(lambda: run(True), reps), ev_ms(lambda: run(False), reps)
        # per-launch GEMM timing: with the look-ahead on, GEMMs of different streams overlap and their event-bracketed
        # durations would count the same wall time more than once, so this pass runs the single-stream schedule
        old_la = L.load().gegp_set_option(L.OPT_LOOKAHEAD, 0)
        try:
            run(True)
            L.profile_begin(True)
            for _ in range(reps):
                run(True)
            pr = L.profile_end()
        finally:
            L.load().gegp_set_option(L.OPT_LOOKAHEAD, old_la)
        gemm_ms = pr["gemm_ms"] / reps
        ld = bk.ld_of(N_)
        buf = torch.empty((N_ + 2, ld), dtype=torch.float64, device="cuda")
        dinv_ = bk.dinv_buffer(N_)
        # the builder runs for tens of microseconds at c2: time a back-to-back burst so that the host-side launch
        # preparation of one call hides behind the previous kernel (every launch rewrites all 8 N^2 bytes; the
        # matrix is larger than L2)
        burst = 20 if N_ < 10000 else 4

        def build_burst(uplo):
            for _ in range(burst):
                bk.build_cov(Xd, thd[0], mode=L.MODE_PRECON, eta=eta_, out=buf[:N_], uplo=uplo)
        ms_full = ev_ms(lambda: build_burst(0), reps) / burst
        ms_low = ev_ms(lambda: build_burst(1), reps) / burst

        def fac():
            bk.build_cov(Xd, thd[0], mode=L.MODE_PRECON, eta=eta_, out=buf[:N_], uplo=1)
            bk.potrf(buf, N_, 0, dinv_)
        ms_one_low = ev_ms(lambda: bk.build_cov(Xd, thd[0], mode=L.MODE_PRECON, eta=eta_, out=buf[:N_], uplo=1), reps)
        ms_chol = ev_ms(fac, reps) - ms_one_low
        del buf
        return {"n": n_, "d": d_, "N": N_, "lml_grad_ms": ms_grad, "lml_only_ms": ms_val,
                "lml_grad_evals_per_s": 1e3 / ms_grad, "overall_tflops_N3": N_ ** 3 / ms_grad * 1e-9,
                "cholesky_ms": ms_chol, "cholesky_tflops": N_ ** 3 / 3 / ms_chol * 1e-9,
                "cholesky_frac_of_dmma_peak": N_ ** 3 / 3 / ms_chol * 1e-9 / FP64_DMMA_PEAK_TFLOPS,
                "gemm_ms_per_eval": gemm_ms, "gemm_launches_per_eval": pr["gemm_launches"] // reps,
                "gemm_tflops_executed": pr["gemm_flops"] / reps / gemm_ms * 1e-9,
                "gemm_share_of_step": gemm_ms / ms_grad,
                "build_full_ms": ms_full, "build_full_gbs": 8.0 * N_ * N_ / ms_full * 1e-6,
                "build_lower_ms": ms_low, "build_lower_gbs": 4.0 * N_ * (N_ + 1) / ms_low * 1e-6}

    if args.no_phases:      # launch-list / ncu runs: only the timed loop and the e2e loop
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "impl": "b200", "gpu_launches": int(launches), "e2e": e2e,
                "note": "--no-phases run (profiling aid), not a bench line"}
        print(json.dumps(line), flush=True)
        return
    ph = phase_numbers(n, d, reps=3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:   # dram bytes of the dominant launch from the committed ncu --set full capture (per launch)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01", "dominant_kernel_ncu.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": "fp64 DMMA.8x8x4 GEMM engine = all O(N^3) work of one evaluation: gemm_f64_kernel<64,64,32,32,{4|2}> "
                          "(cp.async ring; 2 stages when the launch has >= 4 CTAs per SM) carries most of this workload's "
                          "products (< 400 tiles of 128x128 each), gemm_f64_kernel<32,32,16,16,4> the K=128 updates on the "
                          "factorisation's critical path, gemm_tma_nt_kernel<128,64,4,2> (TMA + mbarrier ring) the final "
                          "U U^T product and everything at the c3 size",
                "bound": "tensor", "achieved": N ** 3 / ph["gemm_ms_per_eval"] * 1e-9, "peak": FP64_DMMA_PEAK_TFLOPS,
                "unit": "TFLOP/s", "frac": N ** 3 / ph["gemm_ms_per_eval"] * 1e-9 / FP64_DMMA_PEAK_TFLOPS,
                "traffic": traffic,
                "algorithmic_flops_per_eval": float(N) ** 3,
                "note": "achieved = N^3 (SURVEY 8d: N^3/3 factor + 2N^3/3 inverse) / summed CUDA-event duration of all "
                        "GEMM launches of one evaluation (timed on the single-stream schedule: with the look-ahead on, "
                        "launches of different streams overlap); traffic = dram bytes of ONE captured launch of the dominant kernel "
                        "(profiles/r01/ncu_full_summary_v2.txt); peak = measured fp64 DMMA issue peak of this pool's B200 "
                        "(profiles/r01/dmma_peak.log; cuBLAS DGEMM reaches %.2f); MEASURED_PEAKS.json has no fp64 entry"
                        % FP64_DGEMM_TFLOPS,
                "hbm_build": {"bound": "hbm", "achieved": ph["build_full_gbs"], "peak": hbm_peak, "unit": "GB/s",
                              "frac": ph["build_full_gbs"] / hbm_peak, "bytes": 8.0 * N * N,
                              "lower_only_gbs": ph["build_lower_gbs"]}}
    extra = {}
    if world == 1:
        # batch point B (multi-start rows in lock step): several candidates per device call fill the GPU where one
        # N ~ 5000 evaluation is bound by the latency of its leaf chain.  Reported beside the headline, never instead.
        try:
            Bc = 4
            thb = torch.stack([bk.to_dev(step_theta(theta, 1000 + s, rank)) for s in range(Bc)])
            runb = lambda: bk.lml_eval(X_dev, y_dev, thb, mode=L.MODE_PRECON, eta=eta, want_grad=True)  # noqa: E731
            runb()
            msb = ev_ms(runb, 3)
            extra["batched"] = {"candidates_per_call": Bc, "ms_per_call": msb, "evals_per_s": Bc / (msb * 1e-3),
                                "note": "same workload, 4 candidate thetas per gegp_lml_eval call (lock-step multi-start)"}
        except Exception as exc:
            extra["batched"] = {"error": repr(exc)}
    if args.workload != "c3" and not args.no_c3 and world == 1:
        try:
            extra["c3"] = phase_numbers(*WORKLOADS["c3"], reps=2)
        except Exception as exc:  # report, never hide
            extra["c3"] = {"error": repr(exc)}
    bk.free_workspace()
    torch.cuda.empty_cache()

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import gegp_oracle as O   # checker / baseline only
        cores = _use_all_host_threads()
        t0 = time.perf_counter()
        ref = O.lkd_wo_noise(x, f, g, step_theta(theta, W + K - 1, 0), "precon", eta, calc_grad=True)
        dt = time.perf_counter() - t0
        got = last.cpu().numpy()[0]
        cpu_baseline = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "1 full LML+gradient evaluation of the same workload (NumPy/SciPy port of the "
                                  "reference algorithm, all host BLAS threads)",
                        "parity_vs_gpu_last_step": {
                            "lml_rel": abs(got[L.OUT_LML] - ref.ln_lkd) / abs(ref.ln_lkd),
                            "grad_rel": float(np.max(np.abs(got[L.OUT_GRAD:] - ref.ln_lkd_grad)) / np.max(np.abs(ref.ln_lkd_grad)))}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: d={d}, n={n}, N={N}, precon, LML+grad (BASELINE configs[1])"
                       if args.workload == "c2" else f"{args.workload}: d={d}, n={n}, N={N}, precon, LML+grad",
                       "per_rank": "one evaluation per step; every rank evaluates its own candidate theta",
                       "l2": f"working set {3 * 8 * N * N / 1e6:.0f} MB per evaluation > 126 MB L2 and rewritten "
                             "from scratch every step (no reuse across steps); no explicit flush",
                       "seed": 0},
            "impl": "b200", "info_ok": info_ok, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
            "roofline": roofline, "phases": ph, "cpu_baseline": cpu_baseline}
    line.update(extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-c3", action="store_true", help="skip the extra north-star (N=21000) measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-phases", action="store_true", help="profiling aid: skip the roofline / phase measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
