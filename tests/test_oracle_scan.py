"""The Lindley oracle against the event-loop oracle (which is pinned bit-exactly to the reference's
QsSim): for c = 1 and an infinite buffer the recursion W[n+1] = max(0, W[n] + S[n] - A[n+1])
reproduces QsSim's waits and sojourns; the tree evaluation agrees with the sequential one to rounding."""

import numpy as np

from oracle import oracle


def test_lindley_matches_event_loop_oracle():
    rng = np.random.default_rng(5)
    N = 20000
    A = rng.gamma(3.18878, 1 / 3.18878, N + 50)
    S = 0.395878 * rng.random(N + 50) ** (-1 / 2.30171)
    ev = oracle.fifo_replay(1, None, A, S, N, samples=True)
    W, o = oracle.lindley(A, S, N, tree=False)
    # the event loop works on absolute clocks (base.py:236,272,302), the recursion on differences
    np.testing.assert_allclose(W, ev.w_samples[:N], rtol=0, atol=2e-11 * A[:N].sum())
    np.testing.assert_allclose(W + S[:N], ev.v_samples, rtol=0, atol=2e-11 * A[:N].sum())
    assert o.taked == ev.taked and o.served == ev.served
    np.testing.assert_allclose(o.w_sum, ev.w_sum, rtol=1e-9)
    np.testing.assert_allclose(o.v_sum, ev.v_sum, rtol=1e-9)
    np.testing.assert_allclose(o.sum_a + (W[-1] + S[N - 1]), ev.ttek, rtol=1e-13)


def test_tree_agrees_with_sequential():
    rng = np.random.default_rng(6)
    for N in (1, 15, 16, 17, 4095, 4096, 4097, 70000):
        A = rng.exponential(1.0, N + 1)
        S = rng.exponential(0.8, N)
        Ws, os_ = oracle.lindley(A, S, N, tree=False)
        Wt, ot = oracle.lindley(A, S, N, tree=True)
        np.testing.assert_allclose(Wt, Ws, rtol=0, atol=1e-9)
        assert (ot.taked, ot.served) == (os_.taked, os_.served)
        np.testing.assert_allclose([ot.a, ot.b], [os_.a, os_.b], rtol=1e-12, atol=1e-9)
        # carry semantics: walking from w_in equals max(a, w_in + b) at the end
        W2, o2 = oracle.lindley(A, S, N, w_in=3.5, tree=False)
        w_end = max(0.0, W2[-1] + S[N - 1] - A[N])
        assert abs(w_end - max(os_.a, 3.5 + os_.b)) < 1e-9
