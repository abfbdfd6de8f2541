"""CPU suite for the host-side planning of the tcgen05 path (pure arithmetic behind ragfin_debug_plan, no device):
the invariants exactness rests on.  A slice plan that skipped or repeated a row, or a bound-pass sample that visited a
tile twice (its rank-th largest block maximum would no longer be reached by `rank` DISTINCT rows) would make results
wrong without any kernel being at fault."""
import ctypes
import itertools

import pytest

KGM, KGN = 128, 256


@pytest.fixture(scope="module")
def plan():
    from ragfin_b200 import _lib
    L = _lib.load()

    def run(nq, n_rows, num_sms=148, k=10, cluster=0):
        out = (ctypes.c_int64 * 10)()
        _lib.check(L.ragfin_debug_plan(nq, n_rows, num_sms, k, cluster, out))
        keys = ("C", "QT", "S", "rows_per_slice", "grid", "append", "bound", "nblk", "g", "bstride")
        return dict(zip(keys, [int(v) for v in out]))
    return run


def cand_per_query(k):   # K' of DESIGN.md 2.3
    want = k + max(16, k // 4)
    for kp in (32, 64, 128, 256):
        if want <= kp:
            return kp
    return 256


SHAPES = list(itertools.product([1, 2, 16, 17, 128, 129, 300, 512, 1024, 4096],
                                [1, 255, 256, 257, 10_000, 16 * 256, 70_000, 1_250_000, 10_000_000, 12_500_000, 2_000_000_000],
                                [1, 10, 100, 256]))


@pytest.mark.parametrize("num_sms,cluster", [(148, 0), (148, 1), (148, 2), (148, 4), (132, 0), (8, 4)])
def test_slices_cover_every_row_once_and_the_grid_fits(plan, num_sms, cluster):
    for nq, n, k in SHAPES:
        p = plan(nq, n, num_sms, k, cluster)
        assert p["QT"] == -(-nq // KGM)
        assert p["C"] in (1, 2, 4) and (cluster == 0 or p["C"] == cluster)
        assert p["rows_per_slice"] % KGN == 0 and p["rows_per_slice"] > 0
        assert p["S"] * p["rows_per_slice"] >= n, (nq, n, k, p)              # the slices reach the last row ...
        assert (p["S"] - 1) * p["rows_per_slice"] < n, (nq, n, k, p)         # ... and none of them is empty
        assert p["grid"] % p["C"] == 0 and p["C"] <= p["grid"] <= num_sms, (nq, n, k, p)
        groups = -(-p["QT"] // p["C"])
        assert p["grid"] // p["C"] <= groups * p["S"]                        # no cluster without a work item


def test_automatic_cluster_size(plan):
    assert [plan(nq, 10_000_000)["C"] for nq in (1, 128, 129, 256, 257, 384, 896, 897, 1024, 4096)] == [1, 1, 2, 2, 2, 2, 2, 2, 2, 2]   # pairs from two query tiles up (profiles/r02/policy_sweep.log)


def test_bound_pass_samples_distinct_tiles_and_never_the_last(plan):
    for nq, n, k in SHAPES:
        p = plan(nq, n, 148, k, 0)
        n_tiles = -(-n // KGN)
        kp = cand_per_query(k)
        assert p["append"] == int(k <= 256 and n_tiles >= 4 * k and n < 2 ** 31 - KGN)
        assert p["bound"] == int(p["append"] or n_tiles >= 4 * kp)
        if not p["bound"]:
            continue
        rank = k if p["append"] else kp
        sample_tiles = p["nblk"] * p["g"]
        assert p["nblk"] >= 2 * rank, (nq, n, k, p)                          # the rank-th largest block maximum exists, with slack
        assert p["nblk"] <= 1024                                             # bound_select_kernel's shared-memory staging
        assert p["g"] >= 1 and p["bstride"] >= 1
        assert sample_tiles <= n_tiles // 2, (nq, n, k, p)                   # at most half of the corpus is sampled
        assert (sample_tiles - 1) * p["bstride"] <= n_tiles - 2, (nq, n, k, p)   # distinct tiles (stride >= 1), last tile never


def test_sample_fraction_policy(plan):
    """One query tile: ~0.4 % (the pass is latency); tensor-bound batches: 0.8 % x sqrt(k)."""
    n = 10_000_000
    n_tiles = -(-n // KGN)
    p = plan(1, n, 148, 10)
    assert 0.003 < p["nblk"] * p["g"] / n_tiles < 0.012
    p = plan(4096, n, 148, 10)
    assert 0.02 < p["nblk"] * p["g"] / n_tiles < 0.035
    p = plan(4096, n, 148, 100)
    assert 0.06 < p["nblk"] * p["g"] / n_tiles < 0.10


def test_plan_rejects_bad_arguments():
    from ragfin_b200 import _lib
    L = _lib.load()
    out = (ctypes.c_int64 * 10)()
    assert L.ragfin_debug_plan(0, 10, 148, 10, 0, out) == _lib.EINVAL
    assert L.ragfin_debug_plan(1, 10, 148, 10, 3, out) == _lib.EINVAL
    assert L.ragfin_debug_plan(1, 10, 148, 10, 0, None) == _lib.EINVAL


def test_fused_sweep_visits_every_tile_of_a_slice_exactly_once():
    """csrc/sweep_fused.cuh visits the tiles of a slice in a strided permutation starting mid-slice (so that a corpus sorted by
    similarity cannot make every row beat the running bound).  The library's own function, run on the host for every slice
    length up to 3 000 tiles and a few large ones: a permutation of 0 .. n - 1, first tile n // 2, consecutive visits far apart."""
    import numpy as np
    from ragfin_b200 import _lib
    L = _lib.load()
    for n in list(range(1, 3001)) + [39063, 48829, 65536, 100003, 390625]:
        out = np.empty(n, np.int32)
        _lib.check(L.ragfin_debug_fused_tile_order(n, out.ctypes.data))
        assert np.array_equal(np.sort(out), np.arange(n)), n
        assert out[0] == n // 2
        if n >= 16:
            step = np.abs(np.diff(out.astype(np.int64)))
            step = np.minimum(step, n - step)
            assert step.min() >= n // 4, (n, step.min())          # no two consecutive visits close to each other
    assert L.ragfin_debug_fused_tile_order(0, None) == _lib.EINVAL


def test_host_buffers_numpy_and_torch():
    """engine._host_queries / _host_out: what Index.search and PeerExchange.search_sharded_host hand to the C ABI.  numpy and
    torch CPU buffers give the address of the caller's own memory; anything the library could not write into is refused."""
    import numpy as np
    import torch
    from ragfin_b200 import engine
    qn = np.arange(2 * 8, dtype=np.float32).reshape(2, 8)
    obj, ptr, nq = engine._host_queries(qn, 8)
    assert obj is qn and ptr == qn.ctypes.data and nq == 2
    obj, ptr, nq = engine._host_queries(qn[0].tolist(), 8)                  # one query as a list
    assert nq == 1 and obj.shape == (1, 8) and obj.dtype == np.float32
    qt = torch.from_numpy(qn)
    obj, ptr, nq = engine._host_queries(qt, 8)
    assert obj is qt and ptr == qt.data_ptr() == qn.ctypes.data and nq == 2
    obj, ptr, nq = engine._host_queries(qt.double()[:, :], 8)               # wrong dtype: converted, not reinterpreted
    assert obj.dtype == torch.float32 and torch.equal(obj, qt) and ptr == obj.data_ptr()
    obj, ptr, nq = engine._host_queries(qt.t().contiguous().t(), 8)         # not contiguous: compacted
    assert obj.is_contiguous() and torch.equal(obj, qt)
    obj, ptr, nq = engine._host_queries(qt[1], 8)
    assert nq == 1 and tuple(obj.shape) == (1, 8)
    for bad in (np.zeros((2, 7), np.float32), torch.zeros(2, 9), np.zeros((2, 2, 8), np.float32)):
        with pytest.raises(ValueError):
            engine._host_queries(bad, 8)

    out, p = engine._host_out(None, 2, 5, np.int64, "out_ids")
    assert out.shape == (2, 5) and out.dtype == np.int64 and p == out.ctypes.data
    mine = torch.empty((2, 5), dtype=torch.int64)
    out, p = engine._host_out(mine, 2, 5, np.int64, "out_ids")
    assert out is mine and p == mine.data_ptr()
    mine = np.empty((2, 5), np.float32)
    out, p = engine._host_out(mine, 2, 5, np.float32, "out_scores")
    assert out is mine and p == mine.ctypes.data
    for bad in (torch.empty((2, 5), dtype=torch.int32), torch.empty((5, 2), dtype=torch.int64), torch.empty((2, 10), dtype=torch.int64)[:, ::2],
                np.empty((2, 5), np.int32), np.empty((2, 10), np.int64)[:, ::2], np.empty((3, 5), np.int64)):
        with pytest.raises(ValueError):
            engine._host_out(bad, 2, 5, np.int64, "out_ids")
