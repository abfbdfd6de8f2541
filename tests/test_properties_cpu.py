"""Property tests (hypothesis) of the size-independent facts the GPU suites lean on, on the CPU oracle and the host
logic: any row sharding + merge equals the unsharded search, ties always resolve to the lower id, scaling rows or
queries by positive factors changes nothing (cosine), k > N pads, shard_bounds partitions exactly."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import ragfin_oracle as O
from ragfin_b200.sharded import shard_bounds

FAST = settings(max_examples=25, deadline=None)


@FAST
@given(n=st.integers(0, 5000), world=st.integers(1, 9))
def test_shard_bounds_partition_the_rows(n, world):
    nxt = 0
    for r in range(world):
        row0, cnt = shard_bounds(n, world, r)
        assert row0 == nxt and cnt >= 0
        nxt = row0 + cnt
    assert nxt == n
    sizes = [shard_bounds(n, world, r)[1] for r in range(world)]
    assert sizes == sorted(sizes, reverse=True)            # contiguous blocks of ceil(n / world), the remainder last


@FAST
@given(seed=st.integers(0, 10_000), n=st.integers(1, 400), k=st.integers(1, 40), cuts=st.lists(st.integers(0, 400), max_size=5),
       dtype=st.sampled_from(O.DTYPES), dup=st.sampled_from([0, 3, 7]))
def test_any_sharding_merges_to_the_unsharded_result(seed, n, k, cuts, dtype, dup):
    stored = O.normalize_rows(O.synth_rows(seed, 0, n, 24, dup_every=dup), dtype)
    q = O.synth_rows(seed + 1, 0, 3, 24)
    full = O.cosine_topk(q, stored, k)
    edges = sorted({0, n, *[c for c in cuts if c < n]})
    parts = [O.cosine_topk(q, stored[a:b], k, id_base=a) for a, b in zip(edges[:-1], edges[1:])]
    mi, ms = O.merge_topk([p[0] for p in parts], [p[1] for p in parts], k)
    assert np.array_equal(mi, full[0]) and np.array_equal(ms.view(np.uint32), full[1].view(np.uint32))
    m = min(n, k)
    assert (full[0][:, m:] == -1).all() and np.isneginf(full[1][:, m:]).all()          # k > N pads
    for qi in range(3):                                                                  # descending, ties to the lower id
        s, i = full[1][qi, :m], full[0][qi, :m]
        assert all(s[j] > s[j + 1] or (s[j] == s[j + 1] and i[j] < i[j + 1]) for j in range(m - 1))


@FAST
@given(seed=st.integers(0, 10_000), e_rows=st.integers(-6, 6), e_q=st.integers(-6, 6))
def test_cosine_ignores_power_of_two_scaling(seed, e_rows, e_q):
    """Scaling by powers of two is exact in fp32, so the normalised rows - and every score bit - must not change."""
    x = O.synth_rows(seed, 0, 300, 40)
    q = O.synth_rows(seed + 1, 0, 2, 40)
    a = O.cosine_topk(q, O.normalize_rows(x, "bf16"), 10)
    b = O.cosine_topk(q * np.float32(2.0 ** e_q), O.normalize_rows(x * np.float32(2.0 ** e_rows), "bf16"), 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


@FAST
@given(seed=st.integers(0, 10_000), n=st.integers(2, 200), copies=st.integers(2, 12))
def test_exact_duplicates_come_back_in_row_order(seed, n, copies):
    x = O.synth_rows(seed, 0, n, 32)
    q = O.synth_rows(seed + 1, 0, 1, 32)
    rows = np.sort(np.random.default_rng(seed).choice(n, size=min(copies, n), replace=False))
    x[rows] = q[0] * 3.0                                                                  # identical best rows
    ids, sc = O.cosine_topk(q, O.normalize_rows(x, "f16"), len(rows))
    assert ids[0].tolist() == rows.tolist() and len(set(sc[0].view(np.uint32).tolist())) == 1
