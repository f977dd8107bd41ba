"""K3 replay form in one pass (csrc/k3_onepass.cu: TMA-staged tiles, mbarrier pipeline, decoupled look-back) against
the three-launch path (MQ_SCAN_TWO_PASS=1) and the oracle: the association tree is the specification, so per-job waits,
range summaries and the tail are equal BIT FOR BIT; the power sums agree to rounding (different partition into CTAs)."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _both(A, S, N, **kw):
    from most_queue_b200.sim.lindley import replay_chain

    os.environ.pop("MQ_SCAN_TWO_PASS", None)
    one = replay_chain(A, S, N, **kw)
    os.environ["MQ_SCAN_TWO_PASS"] = "1"
    try:
        two = replay_chain(A, S, N, **kw)
    finally:
        os.environ.pop("MQ_SCAN_TWO_PASS", None)
    return one, two


# 2 tiles; ragged last tile; exactly full last tile; > 1024 tiles (chunks of 2, 3 and 10 tiles); a partial last group
@pytest.mark.parametrize("N", [4097, 8192, 70_001, (1 << 20) + 5, 4096 * 1500 + 77, 4096 * 2500, 40_000_000])
def test_one_pass_equals_three_launch_path_bitwise(N):
    rng = np.random.default_rng(N % 1000)
    A = rng.exponential(1.0, N + 1)
    S = rng.exponential(0.9, N)
    (o1, W1), (o2, W2) = _both(A, S, N)
    assert np.array_equal(W1, W2)
    assert o1["pairs_apply"] == o2["pairs_apply"] == o2["pairs"]
    assert o1["tail_raw"] == o2["tail_raw"] and o1["ttek"] == pytest.approx(o2["ttek"], rel=1e-14)
    assert (o1["taked"], o1["served"]) == (o2["taked"], o2["served"])
    np.testing.assert_allclose(o1["w_sum"], o2["w_sum"], rtol=1e-11)
    np.testing.assert_allclose(o1["v_sum"], o2["v_sum"], rtol=1e-11)
    np.testing.assert_allclose([o1["sum_a"], o1["sum_s"]], [A[:N].sum(), S.sum()], rtol=1e-12)


@pytest.mark.parametrize("N", [4097, 300_000, 4096 * 1100 + 9])
def test_one_pass_waits_bitwise_vs_oracle_tree(N):
    from most_queue_b200.sim.lindley import replay_chain
    from oracle import oracle

    rng = np.random.default_rng(N)
    A = rng.gamma(3.18878, 1 / 3.18878, N + 1)
    S = 0.395878 * rng.random(N) ** (-1 / 2.30171)
    out, W = replay_chain(A, S, N)
    Wt, ot = oracle.lindley(A, S, N, tree=True)
    assert np.array_equal(W, Wt)
    assert out["pairs_apply"][0] == (ot.a, ot.b)
    assert out["tail_raw"] == ot.w_last_raw
    np.testing.assert_allclose(out["w_sum"], ot.w_sum, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(out["v_sum"], ot.v_sum, rtol=1e-11)


@pytest.mark.parametrize("rho", [0.97, 1.25])
def test_one_pass_when_the_carry_never_dies(rho):
    """The fused kernel walks every segment speculatively from its in-tile prefix and repairs the segments whose start
    wait depends on the carry into the tile.  Near and above saturation that is most or ALL of them (an unstable queue
    never forgets): the repair then rewrites every wait and every sum, and must still give the three-launch values."""
    rng = np.random.default_rng(11)
    N = 1_500_000
    A = rng.exponential(1.0, N + 1)
    S = rng.exponential(rho, N)
    (o1, W1), (o2, W2) = _both(A, S, N)
    assert np.array_equal(W1, W2)
    assert o1["pairs_apply"] == o2["pairs_apply"] and o1["tail_raw"] == o2["tail_raw"]
    np.testing.assert_allclose(o1["w_sum"], o2["w_sum"], rtol=1e-10)
    np.testing.assert_allclose(o1["v_sum"], o2["v_sum"], rtol=1e-10)
    if rho > 1:
        assert W1[-1] > 1e4  # the wait really grew without bound


def test_one_pass_ranges_with_carries():
    """Ranges with carried waits (the multi-GPU form): every range goes through the fused kernel with w_in != 0."""
    rng = np.random.default_rng(7)
    N = 3_000_000
    A = rng.exponential(1.0, N + 1)
    S = rng.exponential(0.95, N)
    (o1, W1), (o2, W2) = _both(A, S, N, ranges=4)
    assert np.array_equal(W1, W2)
    assert o1["pairs_apply"] == o2["pairs_apply"] == o1["pairs"]
    assert max(o1["carries"]) > 1.0
    (o3, W3), _ = _both(A, S, N, ranges=1)
    np.testing.assert_allclose(W1, W3, rtol=0, atol=1e-8)


def test_one_pass_device_pointers_and_misaligned_fallback():
    """Device-resident arrays (the bench's form); an 8-byte-aligned (not 16) pointer cannot be fetched by TMA and must
    fall back to the three-launch path with the same result."""
    import torch

    from most_queue_b200 import _lib

    ctx = _lib.default_context()
    n = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.empty(n + 3, device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
    S = torch.empty(n + 2, device="cuda", dtype=torch.float64).exponential_(1.0 / 0.8, generator=g)
    torch.cuda.synchronize()
    st1, _ = ctx.scan_apply(n, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
    assert ctx.last_timing()["launches"] == 2  # the fused kernel + the reduction: no silent fall-back to three launches
    os.environ["MQ_SCAN_TWO_PASS"] = "1"
    try:
        st2, _ = ctx.scan_apply(n, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
    finally:
        os.environ.pop("MQ_SCAN_TWO_PASS", None)
    assert st1[_lib.S_A] == st2[_lib.S_A] and st1[_lib.S_B] == st2[_lib.S_B]
    assert st1[_lib.S_TAILRAW] == st2[_lib.S_TAILRAW]
    np.testing.assert_allclose(st1[:10], st2[:10], rtol=1e-11)
    # odd element offset: 8-byte aligned only
    st3, _ = ctx.scan_apply(n, 0.0, A=A.data_ptr() + 8, S=S.data_ptr() + 8, device_ptrs=True)
    Ah, Sh = A[1:n + 2].cpu().numpy(), S[1:n + 1].cpu().numpy()
    np.testing.assert_allclose(st3[_lib.S_SUMS], Sh.sum(), rtol=1e-12)
    np.testing.assert_allclose(st3[_lib.S_SUMA], Ah[:n].sum(), rtol=1e-12)


def test_one_pass_is_reproducible_launch_after_launch():
    """Every launch of the fused kernel on the same arrays gives the same bits (static tile order, fixed-order sums).
    This is the test that would have caught the parity alias of round 2: with ONE "stage is full" mbarrier per stage and
    two compute groups alternating on it, a group that arrived while the other group's tile was still in flight asked
    try_wait.parity for a phase two ahead of the current one, was let through and folded stale data -- one launch in a
    hundred with wrong sums (1e-6 relative: inside every statistical tolerance), one in two thousand hung.  Cheap enough
    to keep: 1500 launches in about a second."""
    import torch

    from most_queue_b200 import _lib

    ctx = _lib.default_context()
    g = torch.Generator(device="cuda").manual_seed(5)
    for n, launches in (((1 << 24) + 12345, 1000), ((1 << 26) + 777, 500)):
        A = torch.empty(n + 1, device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
        S = torch.empty(n, device="cuda", dtype=torch.float64).exponential_(1.0 / 0.8, generator=g)
        torch.cuda.synchronize()
        ref = None
        for i in range(launches):
            st, _ = ctx.scan_apply(n, 0.25, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
            assert ctx.last_timing()["launches"] == 2
            key = tuple(st[:18])
            if ref is None:
                ref = key
            assert key == ref, (n, i)
        del A, S
