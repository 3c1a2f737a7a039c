"""CPU model of the look-ahead factorisation's update schedule (csrc/potrf.cu, chol_node_la).

The recursion deals every trailing update out as pieces -- the chain step with the column update beside it, the K = LEAF
window pieces, the bulk pieces, the part of a node's own window its left child's window did not reach.  The comment in
potrf.cu argues by induction that every (source leaf, target block column) pair is applied exactly once; this test walks
the same recursion on the host (same split rule, same window arithmetic, same piece ranges) for many matrix sizes and
window widths and counts.  It checks the DESIGN of the schedule, not the CUDA code."""
import numpy as np
import pytest

LEAF = 128


def split_point(k):
    h = (k + 1) // 2
    k1 = ((h + LEAF - 1) // LEAF) * LEAF
    if k1 >= k:
        k1 = ((k - 1) // LEAF) * LEAF
    return k1


def walk(row0, k, ext, window, ops, m_cols):
    """ops: (source columns [s0, s1), target columns [t0, t1)), global indices; m_cols: columns of the whole matrix."""
    if k <= LEAF:
        return
    k1 = split_point(k)
    kc = k - k1
    w = min(kc, LEAF)
    ext_l = min(window, kc + ext)
    walk(row0, k1, ext_l, window, ops, m_cols)
    base = row0 + k1

    def piece(a, b, kk):
        if b <= a or base + a >= m_cols:       # (rows below column a exist: mc > a)
            return
        ops.append((base - kk, base, base + a, min(base + b, m_cols)))

    piece(0, w, LEAF)                           # chain step (diagonal block) + column update beside it (rows below)
    for a in range(w, ext_l, LEAF):             # window of the left child: its last leaf only
        piece(a, min(a + LEAF, ext_l), LEAF)
    if kc > ext_l:                              # bulk (cut into just-in-time pieces: a partition of this range)
        piece(ext_l, kc, k1)
    if ext > 0 and kc + ext > ext_l:            # the part of this node's own window the left child's did not reach
        piece(max(kc, ext_l), kc + ext, k1)
    walk(base, kc, ext, window, ops, m_cols)


@pytest.mark.parametrize("window", [128, 256, 512])
@pytest.mark.parametrize("N", [129, 256, 300, 384, 640, 1000, 1200, 1408, 2500, 2684, 5500, 8192, 21000])
def test_every_source_leaf_reaches_every_target_block_column_exactly_once(N, window):
    ops = []
    walk(0, N, 0, window, ops, N)
    nb = (N + LEAF - 1) // LEAF
    count = np.zeros((nb, nb), dtype=np.int64)   # [source leaf, target block column]
    for s0, s1, t0, t1 in ops:
        assert s0 % LEAF == 0 and s1 % LEAF == 0 and t0 % LEAF == 0 and s0 >= 0 and t0 >= s1
        for s in range(s0 // LEAF, s1 // LEAF):
            for t in range(t0 // LEAF, (t1 + LEAF - 1) // LEAF):
                count[s, t] += 1
    want = np.triu(np.ones((nb, nb), dtype=np.int64), 1)
    bad = np.argwhere(count != want)
    assert bad.size == 0, f"N={N} window={window}: (source leaf, target column) pairs not applied exactly once: {bad[:8].tolist()}"


def test_split_point_keeps_leaf_alignment():
    for k in range(129, 6000, 7):
        k1 = split_point(k)
        assert k1 % LEAF == 0 and 0 < k1 < k and k1 >= k - k1 - LEAF
