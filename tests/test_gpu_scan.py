"""K3 (max-plus Lindley scan) through the C ABI: per-job waits bit-exact against the oracle's
evaluation of the same association tree, carries across ranges, generator form against K2."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class P:
    def __init__(self, **k):
        self.__dict__.update(k)


GAMMA = P(mu=3.18878, alpha=3.18878)      # mean 1.0, cv 0.56  (cfg3 source)
PARETO = P(alpha=2.30171, K=0.395878)     # mean 0.7, cv 1.2   (cfg3 server)


@pytest.mark.parametrize("N", [1, 15, 16, 17, 255 * 16 + 3, 4096, 4097, 100_003, 1_300_000])
def test_replay_waits_bitwise_vs_oracle_tree(N):
    from most_queue_b200.sim.lindley import replay_chain
    from oracle import oracle

    rng = np.random.default_rng(N)
    A = rng.gamma(3.18878, 1 / 3.18878, N + 1)
    S = 0.395878 * rng.random(N) ** (-1 / 2.30171)
    out, W = replay_chain(A, S, N)
    Wt, ot = oracle.lindley(A, S, N, tree=True)
    assert np.array_equal(W, Wt)                      # bit for bit, same association tree
    Ws, os_ = oracle.lindley(A, S, N, tree=False)
    np.testing.assert_allclose(W, Ws, rtol=0, atol=1e-9)   # and equal to the sequential walk to rounding
    assert (out["taked"], out["served"]) == (ot.taked, ot.served)
    assert out["pairs"][0] == (ot.a, ot.b)
    np.testing.assert_allclose(out["w_sum"], ot.w_sum, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(out["v_sum"], ot.v_sum, rtol=1e-12)
    np.testing.assert_allclose([out["sum_a"], out["sum_s"]], [ot.sum_a, ot.sum_s], rtol=1e-12)
    assert out["tail_raw"] == ot.w_last_raw


def test_ranges_with_carries_equal_oracle_per_range():
    """The multi-GPU path on one GPU: 5 contiguous ranges, summaries folded into carries; every range's
    waits equal the oracle's tree evaluation of that range from the same carry, bit for bit."""
    from most_queue_b200.sim.lindley import replay_chain
    from oracle import oracle

    rng = np.random.default_rng(42)
    N = 300_000
    A = rng.exponential(1.0, N + 1)
    S = rng.exponential(0.93, N)   # heavy load: carries are large and matter
    out, W = replay_chain(A, S, N, ranges=5)
    bounds = [N * g // 5 for g in range(6)]
    for g in range(5):
        lo, hi = bounds[g], bounds[g + 1]
        Wt, ot = oracle.lindley(A[lo:], S[lo:], hi - lo, w_in=out["carries"][g], tree=True)
        assert np.array_equal(W[lo:hi], Wt)
        assert out["pairs"][g] == (ot.a, ot.b)
    Ws, _ = oracle.lindley(A, S, N, tree=False)
    np.testing.assert_allclose(W, Ws, rtol=0, atol=1e-8)
    assert max(out["carries"]) > 1.0


def test_generator_chain_matches_k2_single_channel():
    """A short chain simulated by the scan and by the event-loop kernel (c=1, replication 0) consumes
    the same Philox words: the moments agree to fp32 accuracy."""
    from most_queue_b200.sim.base import QsSim
    from most_queue_b200.sim.lindley import run_chain

    N, seed = 200_000, 31
    res = run_chain(("Gamma", GAMMA), ("Pa", PARETO), N, seed=seed)
    qs = QsSim(1, seed=seed)
    qs.set_sources(GAMMA, "Gamma")
    qs.set_servers(PARETO, "Pa")
    ev = qs.run(N)
    assert res.served == N and abs(res.taked - qs.taked) <= 1
    np.testing.assert_allclose(res.w[0], ev.w[0], rtol=2e-3)
    np.testing.assert_allclose(res.v[:2], ev.v[:2], rtol=2e-3)
    np.testing.assert_allclose(res.p[0], ev.p[0], rtol=2e-3)
    np.testing.assert_allclose(res.ttek, qs.ttek, rtol=1e-4)
    assert abs(res.utilization - 0.7) < 1e-4


def test_long_generator_chain_mg1_theory():
    """M/M/1 sanity on 2e8 jobs: E[W] = rho / (mu - lambda), utilisation = rho."""
    from most_queue_b200.sim.lindley import run_chain

    res = run_chain(("M", 1.0), ("M", 1.0 / 0.7), 200_000_000, seed=7)
    assert abs(res.w[0] - 0.7 * 0.7 / 0.3) < 0.01
    assert abs(res.measured_utilization - 0.7) < 2e-4
    assert abs(res.v[0] - (0.7 * 0.7 / 0.3 + 0.7)) < 0.01


def test_time_in_state_levels_match_event_loop_oracle():
    """p on the scan path (job-local level walk) against the event-loop oracle on the same arrays, with the
    chain cut into ranges as a multi-GPU run would."""
    from most_queue_b200.sim.lindley import replay_chain
    from oracle import oracle

    rng = np.random.default_rng(77)
    for N, ranges, mean_s in ((50_000, 1, 0.7), (130_000, 3, 0.9)):
        A = rng.gamma(3.18878, 1 / 3.18878, N + 400)   # the event loop also consumes the arrivals after job N-1
        S = rng.exponential(mean_s, N + 400)
        out, _ = replay_chain(A[:N + 1], S[:N], N, ranges=ranges, want_w=False, want_p=True)
        ev = oracle.fifo_replay(1, None, A, S, N, p_bins=64)
        p_ref = ev.p_time / ev.ttek
        # only jobs 0..N-1 are counted on the scan path: the few arrivals after the last one are missing
        np.testing.assert_allclose(out["p"][:24], p_ref[:24], rtol=0, atol=40.0 / ev.ttek + 2e-6)
        assert abs(sum(out["p"]) + out["p_tail"] - 1.0) < 1e-9
        np.testing.assert_allclose(out["ttek"], ev.ttek, rtol=1e-12)


def test_generator_chain_p_matches_k2():
    from most_queue_b200.sim.base import QsSim
    from most_queue_b200.sim.lindley import run_chain

    N, seed = 300_000, 9
    res = run_chain(("Gamma", GAMMA), ("Pa", PARETO), N, seed=seed, want_p=True)
    qs = QsSim(1, seed=seed)
    qs.set_sources(GAMMA, "Gamma")
    qs.set_servers(PARETO, "Pa")
    ev = qs.run(N)
    np.testing.assert_allclose(res.p[:12], ev.p[:12], rtol=0, atol=3e-3)
    assert abs(res.p[0] - 0.3) < 0.01


def test_generator_chain_sums_equal_the_sums_of_the_variates():
    """The chain's sums of gaps and services are the fp64 sums of the very fp32 variates the generator spec defines
    (here drawn a second time through the debug entry point), with and without the draw stash, in one range or two:
    catches a walk that reads a tile's operands while another warp already stages the next tile."""
    import os

    from most_queue_b200 import _lib as L

    ctx = L.default_context()
    src, srv = L.make_dist("Gamma", GAMMA), L.make_dist("Pa", PARETO)
    N, seed = 30_000_000, 3
    S = ctx.variates(srv, seed, 0, 1, N).astype(np.float64)
    A = ctx.variates(src, seed, 0, 0, N).astype(np.float64)
    def run(n=N, w_in=0.0, job0=0, **env):
        for k, v in env.items():
            os.environ[k] = v
        try:
            return ctx.scan_apply(n, w_in, job0=job0, src=src, srv=srv, seed=seed, last_range=False)[0]
        finally:
            for k in env:
                del os.environ[k]

    one = run(MQ_SCAN_SPEC="1")                                        # speculative walk in the summary pass + repair
    stash = run()                                                      # summary with draw stash, walk pass (the default)
    assert np.array_equal(stash, run(MQ_SCAN_NO_SPEC="1"))
    nostash = run(MQ_SCAN_NO_STASH="1")                                # no room for a stash: speculative mode by itself
    assert np.array_equal(nostash, one)
    two = run(MQ_SCAN_NO_SPEC="1", MQ_SCAN_NO_STASH="1")               # everything drawn twice
    three = run(MQ_SCAN_NO_SPEC="1", MQ_SCAN_NO_PREFIX="1")            # the walk folds and scans again
    h = N // 2 + 7
    p1 = run(h)
    p2 = run(N - h, p1[L.S_WOUT], job0=h)
    q1 = run(h, MQ_SCAN_SPEC="1")
    q2 = run(N - h, q1[L.S_WOUT], job0=h, MQ_SCAN_SPEC="1")
    assert q1[L.S_WOUT] == p1[L.S_WOUT] and q2[L.S_WOUT] == p2[L.S_WOUT]
    for st in (one, stash, two):
        np.testing.assert_allclose(st[L.S_SUMS], S.sum(), rtol=1e-12)
        np.testing.assert_allclose(st[L.S_SUMA], A.sum(), rtol=1e-12)
    np.testing.assert_allclose(p1[L.S_SUMS] + p2[L.S_SUMS], S.sum(), rtol=1e-12)
    np.testing.assert_allclose(p1[L.S_SUMA], A[:h].sum(), rtol=1e-12)
    # the walk that reuses the summary pass's per-segment prefixes does the same arithmetic as the one that folds and
    # scans again: identical statistics
    assert np.array_equal(three, stash)
    # the range summary, the carry out and the tail are the same bits on every path (one association tree)
    for row in (L.S_A, L.S_B, L.S_WOUT, L.S_TAILRAW, L.S_TAILV, L.S_TAKED, L.S_SERVED):
        assert one[row] == stash[row] == two[row], row
    # the waits are functions of the same draws: all forms and both cuts agree to rounding (all four moments)
    for row in list(range(L.S_W, L.S_W + 4)) + list(range(L.S_V, L.S_V + 4)):
        np.testing.assert_allclose(one[row], stash[row], rtol=1e-11)
        np.testing.assert_allclose(two[row], stash[row], rtol=1e-11)
        np.testing.assert_allclose(p1[row] + p2[row], stash[row], rtol=1e-10)
        np.testing.assert_allclose(q1[row] + q2[row], stash[row], rtol=1e-10)


@pytest.mark.parametrize("rho", [0.9, 1.1])
def test_speculative_generator_walk_near_and_above_saturation(rho):
    """Generator form, speculative mode against the walk pass when the carry rarely (or never) dies: the repair then
    re-walks most (all) segments; M/M/1 so that the run is cheap."""
    import os

    from most_queue_b200 import _lib as L

    ctx = L.default_context()
    src, srv = L.make_dist("M", 1.0), L.make_dist("M", 1.0 / rho)
    N = 5_000_000 + 321
    os.environ["MQ_SCAN_SPEC"] = "1"
    try:
        spec, _ = ctx.scan_apply(N, 0.5, src=src, srv=srv, seed=9, last_range=True)
        assert ctx.last_timing()["launches"] == 6  # summary, 3 x tile scan, repair, reduction
    finally:
        del os.environ["MQ_SCAN_SPEC"]
    walk, W = ctx.scan_apply(N, 0.5, src=src, srv=srv, seed=9, last_range=True, want_w=True)
    for row in (L.S_A, L.S_B, L.S_WOUT, L.S_TAILRAW, L.S_TAILV, L.S_TAKED, L.S_SERVED):
        assert spec[row] == walk[row], row
    for row in list(range(L.S_W, L.S_W + 4)) + list(range(L.S_V, L.S_V + 4)) + [L.S_SUMA, L.S_SUMS]:
        np.testing.assert_allclose(spec[row], walk[row], rtol=1e-10)
    assert W[0] == 0.5 and (rho < 1 or W[-1] > 1e3)


def test_levels_of_ranges_equal_single_range_on_a_fresh_context():
    """with host pointers the walk of a non-last range reads gaps past A[jobs] for the time-in-state levels;
    the staged copy must hold them.  A FRESH context (nothing left in the staging buffer by an earlier call) and three
    ranges must give the levels of the single range to rounding."""
    from most_queue_b200 import _lib
    from most_queue_b200.sim import lindley

    rng = np.random.default_rng(11)
    N = 180_000
    A = rng.exponential(1.0, N + 1)
    S = rng.exponential(0.85, N)

    def run(ranges):
        ctx = _lib.Context()
        old = _lib._default_ctx.copy()
        _lib._default_ctx.clear()
        _lib._default_ctx[-1] = ctx
        try:
            out, _ = lindley.replay_chain(A, S, N, ranges=ranges, want_w=False, want_p=True)
        finally:
            _lib._default_ctx.clear()
            _lib._default_ctx.update(old)
            ctx.close()
        return out

    one, three = run(1), run(3)
    np.testing.assert_allclose(three["levels"], one["levels"], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(three["p"][:24], one["p"][:24], rtol=0, atol=1e-12)


@pytest.mark.parametrize("spec", [False, True])
def test_summary_keep_then_apply_walks_from_the_kept_state(spec, monkeypatch):
    """MQ_SCAN_KEEP (the multi-GPU sequence summary -> exchange -> apply on one context): the apply that follows the kept
    summary gives the statistics of the stand-alone apply bit for bit, and a call in between invalidates the kept state
    without changing the answer."""
    from most_queue_b200 import _lib

    if spec:  # the speculative mode: the kept summary has already walked, the apply only repairs
        monkeypatch.setenv("MQ_SCAN_SPEC", "1")
    ctx = _lib.default_context()
    src, srv = _lib.make_dist("Gamma", GAMMA), _lib.make_dist("Pa", PARETO)
    n, j0, ahead = 3_000_017, 1_000_000, 5_000
    alone, _ = ctx.scan_apply(n, 0.37, job0=j0, src=src, srv=srv, seed=9, last_range=False, jobs_ahead=ahead)
    t_alone = ctx.last_timing()["sim_ms"]
    ab = ctx.scan_summary(n, job0=j0, src=src, srv=srv, seed=9, keep=True, jobs_ahead=ahead)
    kept, _ = ctx.scan_apply(n, 0.37, job0=j0, src=src, srv=srv, seed=9, last_range=False, jobs_ahead=ahead)
    t_kept = ctx.last_timing()["sim_ms"]
    assert np.array_equal(kept, alone)
    assert (kept[_lib.S_A], kept[_lib.S_B]) == ab
    assert t_kept < 0.8 * t_alone           # it only walked
    # something else uses the context in between: the kept state must not be trusted
    ctx.scan_summary(n, job0=j0, src=src, srv=srv, seed=9, keep=True, jobs_ahead=ahead)
    ctx.sim_fifo(2, None, _lib.make_dist("M", 1.0), _lib.make_dist("M", 0.8), 200, 64, 0, 1, p_bins=8)
    again, _ = ctx.scan_apply(n, 0.37, job0=j0, src=src, srv=srv, seed=9, last_range=False, jobs_ahead=ahead)
    assert np.array_equal(again, alone)
    if spec:  # per-job waits need the walk pass: a kept speculative summary is set aside and the range redone
        ctx.scan_summary(n, job0=j0, src=src, srv=srv, seed=9, keep=True, jobs_ahead=ahead)
        st_w, W = ctx.scan_apply(n, 0.37, job0=j0, src=src, srv=srv, seed=9, last_range=False, jobs_ahead=ahead, want_w=True)
        assert W[0] == 0.37 and st_w[_lib.S_WOUT] == alone[_lib.S_WOUT]
        np.testing.assert_allclose(st_w[_lib.S_W], alone[_lib.S_W], rtol=1e-11)
    # a different range description does not match the key
    ctx.scan_summary(n, job0=j0, src=src, srv=srv, seed=9, keep=True, jobs_ahead=ahead)
    other, _ = ctx.scan_apply(n, 0.37, job0=j0 + 16, src=src, srv=srv, seed=9, last_range=False, jobs_ahead=ahead)
    assert not np.array_equal(other[:8], alone[:8])
