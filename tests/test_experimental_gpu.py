"""GPU parity of the opt-in variants that have NOT yet run on a GPU (written after round 1's GPU budget was spent).

Skipped unless RAGFIN_EXPERIMENTAL=1: a kernel nobody has executed must not be able to hang or fail the round-end
`pytest -m gpu` run.  Once scripts/pair_check.py / scripts/pdl_check.py have passed on a B200, drop the gate.
    RAGFIN_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_experimental_gpu.py -m gpu -x -q
"""
import os

import numpy as np
import pytest

from oracle import ragfin_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("RAGFIN_EXPERIMENTAL") != "1", reason="unverified variants: set RAGFIN_EXPERIMENTAL=1"),
              pytest.mark.timeout(120)]


def _assert_same(got, want, what=""):
    gi, gs = got
    wi, ws = want
    assert np.array_equal(gi, wi), f"{what}: ids differ\n{gi}\n{wi}"
    assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32)), f"{what}: score bits differ\n{gs}\n{ws}"


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("n,dim,nq,k", [(30000, 768, 256, 10), (20011, 384, 130, 5), (50000, 128, 1000, 10), (120000, 768, 513, 100),
                                        (66000, 100, 257, 10), (40000, 1024, 384, 1)])
def test_two_sm_mma_sweep_matches_oracle(coracle, dtype, n, dim, nq, k):
    """Variant 4 (csrc/gemm_pair.cuh): >= 2 query tiles in append mode sweep with tcgen05 cta_group::2 pairs; odd tile
    counts leave the second CTA of the last pair on zero padding.  Same bits as the oracle and as variant 1."""
    import ragfin_b200
    x = O.synth_rows(296, 0, n, dim, dup_every=61, zero_every=1999)
    q = O.synth_rows(297, 0, nq, dim)
    x[4000:4030] = q[2] * 2.0        # 30 exact duplicates of one query: ties resolved by row id
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), k)
    idx = ragfin_b200.Index(dim, dtype, capacity=n, device=0)
    idx.add(x)
    idx.set_gemm_min_batch(1)
    idx.set_gemm_variant(4)
    got = idx.search(q, k)
    st = idx.stats()
    assert st["path"] == 1 and st["queries_rescanned"] == 0
    _assert_same(got, want, f"2-SM pairs {dtype} n={n} dim={dim} nq={nq} k={k}")
    idx.set_gemm_variant(1)
    _assert_same(idx.search(q, k), want, "single-CTA MMA")


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_two_sm_mma_raw_scores_equal_single_cta(dtype):
    import torch
    import ragfin_b200
    x = O.synth_rows(5, 0, 5000, 768)
    q = torch.from_numpy(O.synth_rows(6, 0, 300, 768)).cuda()
    idx = ragfin_b200.Index(768, dtype, capacity=5000, device=0)
    idx.add(x)
    idx.set_gemm_variant(1)
    idx.set_gemm_cluster(2)
    base = idx.debug_gemm_scores(q).cpu()
    idx.set_gemm_variant(4)
    got = idx.debug_gemm_scores(q).cpu()
    assert torch.equal(got, base)     # same k-order per output element


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("n,dim,nq,k", [(70000, 128, 1, 10), (70000, 128, 2, 1), (70000, 128, 16, 10), (150000, 64, 5, 16),
                                        (40000, 1024, 7, 10), (66000, 100, 3, 10), (35000, 768, 16, 5), (300000, 768, 1, 10)])
def test_self_seeded_sweep_matches_oracle(coracle, dtype, n, dim, nq, k):
    """Variant 5 (csrc/gemm_rows_seeded.cuh): the <= 16-query sweep samples, synchronises grid-wide and computes its own
    thresholds - two launches fewer than variant 3, same bits as the oracle; repeated calls reuse the monotonic counter."""
    import ragfin_b200
    x = O.synth_rows(196, 0, n, dim, dup_every=61, zero_every=1999)
    q = O.synth_rows(197, 0, nq, dim)
    if nq >= 3:
        x[4000:4030] = q[2] * 2.0
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), k)
    idx = ragfin_b200.Index(dim, dtype, capacity=n, device=0)
    idx.add(x)
    idx.set_gemm_min_batch(1)
    idx.set_gemm_variant(3)
    _assert_same(idx.search(q, k), want, "variant 3")
    base = idx.stats()["launches"]
    idx.set_gemm_variant(5)
    for _rep in range(3):
        got = idx.search(q, k)
        st = idx.stats()
        assert st["path"] == 1 and st["queries_rescanned"] == 0 and st["launches"] == base - 2
        _assert_same(got, want, f"self-seeded {dtype} n={n} dim={dim} nq={nq} k={k}")


def test_view_shares_the_matrix_and_is_read_only(coracle):
    """ragfin_create_view: a second handle over the same device matrix with its own workspace; searches through parent
    and view on two streams at once return what each returns alone."""
    import torch
    import ragfin_b200
    x = O.synth_rows(7, 0, 70000, 128, dup_every=61)
    q = O.synth_rows(8, 0, 5, 128)
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, "bf16"), 10)
    idx = ragfin_b200.Index(128, "bf16", capacity=80000, device=0)
    idx.add(x)
    v = idx.view()
    assert len(v) == 70000
    _assert_same(v.search(q, 10), want, "view")
    with pytest.raises(ragfin_b200.RagfinError):
        v.add(x[:1])
    qd = torch.from_numpy(q).cuda()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for i in range(20):
        h, s = (idx, s1) if i % 2 == 0 else (v, s2)
        with torch.cuda.stream(s):
            outs.append(h.search_device(qd, 10, stream=s))
    torch.cuda.synchronize()
    for ids, sc in outs:
        _assert_same((ids.cpu().numpy(), sc.cpu().numpy()), want, "interleaved")
    v.close()
    idx.close()
