"""CPU model of the append mode's selection rule: with eps >= |approx - exact| for every row,
  (1) T_lb := k-th largest per-block maximum of a sample  <=  T := k-th largest approximate score of the corpus,
  (2) every row of the exact top-k has approx >= T - 2 eps >= T_lb - 2 eps (so the sweep appends it),
  (3) the finalize cut T - 2 eps keeps it.
Approximate scores are modelled as fp32 dot products of the 16-bit-rounded query with the stored rows; exact scores
come from the oracle.  This checks the reasoning (and the eps formula), not the CUDA kernels."""
import numpy as np
import pytest

from oracle import ragfin_oracle as O


def _eps(qhat, q16, ld, dtype):
    """ragfin_api.cu eps_gemm_const + gemm.cuh prep_queries_kernel / qconv_kernel"""
    const = (ld + 64) * 2.0 ** -21 * 1.0625 + 2.0 ** -21
    if dtype == "f32":
        const += 2.0 * 2.0 ** -10 * 1.0625
    d = np.sqrt(((qhat.astype(np.float64) - q16.astype(np.float64)) ** 2).sum(axis=1))
    return (const + d * 1.0078125 + 1e-9).astype(np.float32)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("k", [1, 10, 100])
def test_threshold_keeps_every_row_of_the_exact_topk(dtype, k):
    n, dim, nq, tile = 40960, 96, 6, 256
    x = O.synth_rows(901, 0, n, dim, dup_every=67)
    q = O.synth_rows(902, 0, nq, dim)
    x[5000:5040] = q[1] * 0.7                       # duplicates of a best row: ties at the top
    stored = O.normalize_rows(x, dtype)              # fp32 array holding storage-rounded values
    qhat = O.normalize_rows(q, "f32")
    q16 = O.round_to_storage(qhat, dtype)
    approx = (q16.astype(np.float32) @ stored.T.astype(np.float32)).astype(np.float32)     # [nq, n], fp32 accumulate
    eps = _eps(qhat, q16, dim, dtype)
    for qi in range(nq):
        exact = O.exact_scores(stored, qhat[qi])
        assert np.all(np.abs(approx[qi].astype(np.float64) - exact.astype(np.float64)) <= eps[qi]), "eps is not a bound"
        order = np.lexsort((np.arange(n), -exact.astype(np.float64)))
        top = order[:k]
        # bound pass: every 4th tile is sampled (every tile for k = 100), one block per sample tile
        blocks = approx[qi].reshape(n // tile, tile)[::(1 if k == 100 else 4)].max(axis=1)
        assert len(blocks) >= k
        t_lb = np.sort(blocks)[::-1][k - 1]
        t = np.sort(approx[qi])[::-1][k - 1]
        assert t_lb <= t                                                        # (1)
        thr = np.float32(t_lb) - 2 * eps[qi] - np.float32(2.0 ** -22)
        assert np.all(approx[qi][top] >= thr)                                   # (2): appended
        cut = np.float32(t) - 2 * eps[qi] - np.float32(2.0 ** -22)
        assert np.all(approx[qi][top] >= cut)                                   # (3): rescored
        appended = np.flatnonzero(approx[qi] >= thr)
        kept = appended[approx[qi][appended] >= cut]
        # the exact top-k recomputed from the kept rows only is the oracle's
        sub = exact[kept]
        sub_order = kept[np.lexsort((kept, -sub.astype(np.float64)))][:k]
        assert np.array_equal(sub_order, top)


@pytest.mark.parametrize("n,k,nq", [(300_000, 10, 3), (70_000, 1, 2), (70_000, 16, 4), (41_000, 10, 1)])
def test_self_seeded_sampling_rule(n, k, nq):
    """CPU model of gemm variant 5 (csrc/gemm_rows_seeded.cuh, not yet run on a GPU): every CTA samples the first
    `sample_tiles` tiles of its first slice; blocks are 128-row half tiles; thr = k-th largest block maximum - 2 eps.
    The slice plan comes from the library's own planner (ragfin_debug_plan) and the sample-tile rule restates the
    dispatch in ragfin_api.cu.  Checks: blocks hold distinct rows, at least 2k blocks, the threshold keeps the exact
    top-k and the number of appended rows stays under the buffer capacity (16 384)."""
    import ctypes
    from ragfin_b200 import _lib
    dim, dtype, tile, half = 64, "bf16", 256, 128
    out = (ctypes.c_int64 * 10)()
    _lib.check(_lib.load().ragfin_debug_plan(nq, n, 148, k, 0, out))
    C, QT, S, rps, grid, append = [int(v) for v in out[:6]]
    assert C == 1 and QT == 1 and append == 1 and grid <= S
    n_tiles = -(-n // tile)
    sample_tiles = min(4, max(1, -(-(-(-n_tiles // 256)) // grid)))          # ragfin_api.cu: seed_tiles
    assert grid * 2 * sample_tiles >= 2 * k
    x = O.synth_rows(911, 0, n, dim, dup_every=67)
    q = O.synth_rows(912, 0, nq, dim)
    x[5000:5040] = q[0] * 0.7
    stored = O.normalize_rows(x, dtype)
    qhat = O.normalize_rows(q, "f32")
    q16 = O.round_to_storage(qhat, dtype)
    approx = (q16.astype(np.float32) @ stored.T.astype(np.float32)).astype(np.float32)
    eps = _eps(qhat, q16, dim, dtype)
    seen = np.zeros(n, bool)
    block_rows = []
    for cta in range(grid):                                                  # first slice of CTA `cta` is slice `cta`
        r0, r1 = cta * rps, min((cta + 1) * rps, n)
        first_tiles = max(0, -(-(r1 - r0) // tile))
        for t in range(min(sample_tiles, first_tiles)):
            for h in range(2):
                lo, hi = r0 + t * tile + h * half, min(r0 + t * tile + (h + 1) * half, r1)
                if lo < hi:
                    assert not seen[lo:hi].any()                             # blocks of DISTINCT rows
                    seen[lo:hi] = True
                    block_rows.append((lo, hi))
    assert len(block_rows) >= 2 * k
    for qi in range(nq):
        exact = O.exact_scores(stored, qhat[qi])
        top = np.lexsort((np.arange(n), -exact.astype(np.float64)))[:k]
        blocks = np.array([approx[qi][lo:hi].max() for lo, hi in block_rows], np.float32)
        t_lb = np.sort(blocks)[::-1][k - 1]
        assert t_lb <= np.sort(approx[qi])[::-1][k - 1]
        thr = np.float32(t_lb) - 2 * eps[qi] - np.float32(2.0 ** -22)
        assert np.all(approx[qi][top] >= thr)
        assert int((approx[qi] >= thr).sum()) < 16384
