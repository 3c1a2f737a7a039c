"""List-scheduling model of a GEMM launch whose tiles do different amounts of work (triangular operands clip the k range
per tile).  CTAs are dispatched in linear blockIdx order onto `slots` resident CTA slots; the makespan is compared with
the ideal (total work / slots).  This is the model behind the longest-first tile order of the column-clipped products
(gemm.cu / gemm_tma.cu, KHI_N0): CPU only, no GPU needed.

    python tools/sched_sim.py
"""
import heapq


def makespan(jobs, slots):
    h = [0.0] * slots
    heapq.heapify(h)
    for w in jobs:
        heapq.heappush(h, heapq.heappop(h) + w)
    return max(h)


def report(name, jobs, slots):
    print(f"{name:58s} CTAs {len(jobs):5d}  makespan / ideal = {makespan(jobs, slots) / (sum(jobs) / slots):.3f}")


if __name__ == "__main__":
    OV = 40                       # fixed cost of a tile (prologue + epilogue) in units of k
    S64, STMA = 148 * 4, 148 * 2  # resident CTAs: 2-stage 64 x 64 kernel, TMA kernel
    R = C = 32                    # second product of the inverse at N = 5500, block size 2048: 32 x 32 tiles of 64
    w = lambda c: (c + 1) * 64 + OV          # k range [0, n0 + 64)
    report("U_ab = -T W^T, row-major, short columns first", [w(c) for r in range(R) for c in range(C)], S64)
    report("U_ab = -T W^T, row-major, long columns first", [w(c) for r in range(R) for c in reversed(range(C))], S64)
    report("U_ab = -T W^T, longest tiles first over the grid", [w(c) for c in reversed(range(C)) for r in range(R)], S64)
    R, C = 64, 22                 # the ragged last pair (4096 x 1404)
    report("ragged pair, row-major, long columns first", [w(c) for r in range(R) for c in reversed(range(C))], S64)
    report("ragged pair, longest tiles first over the grid", [w(c) for c in reversed(range(C)) for r in range(R)], S64)
    R = C = 32                    # first product T = U_aa L_ba^T: k from m0, row-major is already longest-first
    report("T = U_aa L_ba^T, row-major", [(2048 - r * 64) + OV for r in range(R) for c in range(C)], S64)
    for N in (5500, 21000):       # K^-1 = U U^T on the TMA kernel: 128 x 64 tiles, lower part, k from m0
        jobs = []
        for by in range((N + 127) // 128):
            for bx in range((N + 63) // 64):
                jobs.append((max(N - by * 128, 16) + 60) if bx < 2 * (by + 1) else 2)
        report(f"K^-1 = U U^T (N = {N}), row-major", jobs, STMA)
