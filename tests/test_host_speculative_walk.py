"""The idea behind the one-pass Lindley scan (csrc/k3_onepass.cu, k3_repair in csrc/k3_lindley.cu), restated in numpy so
that it is checked without a GPU: a tile can be walked from the prefix it has of ITSELF, because the wait at the start of
segment s is max(a_s, w_tile + b_s) and after the first idle period inside the tile that is a_s whatever the carry
w_tile is; only the segments where the two differ have to be walked again.

  W[n+1] = max(0, W[n] + x[n])            most_queue/sim/base.py:300-303 with one channel
  segment map  w -> max(a, w + b)          (a, b) folded sequentially over the segment's jobs
"""

import numpy as np
import pytest

L, SEGS = 16, 32          # jobs per segment, segments per tile (the kernels use 16 x 256)
TILE = L * SEGS


def walk(x, w0):
    w = np.empty(len(x))
    cur = w0
    for i, xi in enumerate(x):
        w[i] = cur
        cur = max(0.0, cur + xi)
    return w, cur


def fold(x):
    a, b = -np.inf, 0.0
    for xi in x:
        a, b = max(0.0, a + xi), b + xi
    return a, b


def then(e, l):   # later o earlier
    return max(l[0], e[0] + l[1]), e[1] + l[1]


@pytest.mark.parametrize("rho,max_share", [(0.5, 0.08), (0.7, 0.12), (0.9, 0.45), (1.2, 1.0)])
def test_speculative_walk_plus_repair_equals_the_sequential_walk(rho, max_share):
    rng = np.random.default_rng(int(rho * 100))
    n_tiles = 60
    N = n_tiles * TILE
    A = rng.exponential(1.0, N + 1)
    S = rng.exponential(rho, N)
    x = S - A[1:]
    W_ref, _ = walk(x, 0.0)

    W = np.empty(N)
    w_tile, repaired = 0.0, 0
    for t in range(n_tiles):
        xt = x[t * TILE:(t + 1) * TILE]
        # what the tile knows of itself: every segment's exclusive prefix inside the tile
        ex, acc = [], (-np.inf, 0.0)
        for s in range(SEGS):
            ex.append(acc)
            acc = then(acc, fold(xt[s * L:(s + 1) * L]))
        # speculative walk: every segment but the first from a_s
        for s in range(1, SEGS):
            W[t * TILE + s * L:t * TILE + (s + 1) * L], _ = walk(xt[s * L:(s + 1) * L], ex[s][0])
        # ... and when the carry into the tile is known, the repair of the segments that depended on it
        for s in range(SEGS):
            w0 = max(ex[s][0], w_tile + ex[s][1])
            if s == 0 or w0 != ex[s][0]:
                W[t * TILE + s * L:t * TILE + (s + 1) * L], _ = walk(xt[s * L:(s + 1) * L], w0)
                repaired += 1
        w_tile = max(acc[0], w_tile + acc[1])
    np.testing.assert_allclose(W, W_ref, rtol=0, atol=1e-9)
    share = repaired / (n_tiles * SEGS)
    assert share <= max_share, share
    if rho <= 0.7:
        # the repaired segments form a prefix of their tile: a queue that has emptied once has forgotten the carry
        assert share < 3.0 / SEGS + 0.05


def test_a_segment_that_is_not_repaired_starts_from_exactly_a_s():
    """The test the kernels make, max(a_s, w_tile + b_s) != a_s, is exact: where it says "equal", the speculative walk
    started from the very bits the true walk starts from (no tolerance involved)."""
    rng = np.random.default_rng(3)
    x = rng.exponential(0.7, 4 * TILE) - rng.exponential(1.0, 4 * TILE)
    for w_tile in (0.0, 0.3, 2.5, 40.0):
        acc = (-np.inf, 0.0)
        cur = w_tile
        for s in range(4 * SEGS):
            seg = x[s * L:(s + 1) * L]
            if s:
                w0 = max(acc[0], w_tile + acc[1])
                assert abs(w0 - cur) < 1e-9             # the prefix map reproduces the sequential walk
                if w0 == acc[0] and acc[0] == cur:      # unrepaired and no rounding difference: identical start
                    assert walk(seg, acc[0])[1] == walk(seg, cur)[1]
            _, cur = walk(seg, cur)
            acc = then(acc, fold(seg))
