"""CPU suite: the oracle against itself (numpy vs C), against an exactly-rounded sum, and
against the committed golden fixtures.  No GPU, no /root/reference."""
import json
import math
import os
from fractions import Fraction

import numpy as np
import pytest

from oracle import ragfin_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_synth_is_fp32_exact_and_matches_c(coracle):
    a = O.synth_rows(1234, 7, 257, 100, dup_every=5, zero_every=13)
    b = coracle.synth_rows(1234, 7, 257, 100, 5, 13)
    assert np.array_equal(a, b)
    assert np.array_equal(a, (a.astype(np.float64) * 65536).round() / 65536)  # multiples of 2^-16
    rows = np.arange(7, 7 + 257)
    assert not a[rows % 13 == 12].any()
    dup = (rows % 5 == 4) & (rows % 13 != 12) & ((rows - 1) % 13 != 12)
    assert np.array_equal(a[dup], a[np.flatnonzero(dup) - 1])


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("dim", [384, 768, 100, 33])
def test_normalize_numpy_equals_c(coracle, dtype, dim):
    x = O.synth_rows(3, 0, 300, dim, zero_every=17)
    a, b = O.normalize_rows(x, dtype), coracle.normalize_rows(x, dtype)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    n = np.linalg.norm(a.astype(np.float64), axis=1)
    nz = np.arange(300) % 17 != 16
    tol = {"f32": 1e-6, "bf16": 4e-3, "f16": 5e-4}[dtype]
    assert np.all(np.abs(n[nz] - 1.0) < tol) and not a[~nz].any()


def test_bf16_rounding_is_rne():
    x = np.array([1.0, 1.00390625, 1.01171875, 1.0078125 + 2 ** -9, -1.00390625, 3.0e-39], dtype=np.float32)
    # 1 + 2^-8 is a tie between 1.0 and 1 + 2^-7 -> even mantissa (1.0); 1 + 3*2^-8 ties up to 1 + 2^-6
    got = O.round_to_storage(x, "bf16")
    assert got[0] == 1.0 and got[1] == 1.0 and got[2] == np.float32(1.015625) and got[4] == -1.0
    assert got[3] == np.float32(1.0078125)  # tie between 1+2^-7 and 1+2^-6 -> even mantissa


@pytest.mark.parametrize("dtype", O.DTYPES)
def test_canonical_dot_within_half_ulp_of_exact(dtype):
    """The canonical fp64 sum, rounded to fp32, equals the exactly rounded dot product except
    possibly when the exact value sits within 2^-50 of a rounding boundary (never in practice)."""
    rng = np.random.default_rng(0)
    st = O.normalize_rows(O.synth_rows(5, 0, 40, 768), dtype)
    q = O.normalize_rows(O.synth_rows(6, 0, 1, 768), "f32")[0]
    got = O.exact_scores(st, q)
    for r in range(40):
        exact = sum(Fraction(float(a)) * Fraction(float(b)) for a, b in zip(st[r], q))
        want = np.float32(float(exact))  # Fraction -> float is correctly rounded; then fp64 -> fp32
        assert abs(float(got[r]) - float(exact)) <= 2 ** -24 * max(abs(float(exact)), 2 ** -126) * 1.0001
        assert got[r] == want or abs(float(exact) - (float(got[r]) + float(want)) / 2) < 1e-12
    assert rng is not None


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("k", [1, 5, 10, 100])
def test_topk_numpy_equals_c(coracle, dtype, k):
    x = O.synth_rows(11, 0, 700, 384, dup_every=7, zero_every=31)
    st = O.normalize_rows(x, dtype)
    q = O.synth_rows(12, 0, 6, 384)
    i1, s1 = O.cosine_topk(q, st, k)
    i2, s2 = coracle.cosine_topk(q, st, k)
    assert np.array_equal(i1, i2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
    # descending, ties by lower id
    for qi in range(6):
        for j in range(k - 1):
            assert s1[qi, j] > s1[qi, j + 1] or (s1[qi, j] == s1[qi, j + 1] and i1[qi, j] < i1[qi, j + 1])


def test_topk_k_larger_than_n_pads(coracle):
    st = O.normalize_rows(O.synth_rows(1, 0, 16, 384), "f32")
    q = O.synth_rows(2, 0, 2, 384)
    for mod in (O, coracle):
        ids, sc = mod.cosine_topk(q, st, 20)
        assert (ids[:, 16:] == -1).all() and np.isneginf(sc[:, 16:]).all()
        assert sorted(ids[0, :16].tolist()) == list(range(16))


def test_duplicates_tie_break_to_lower_id(coracle):
    x = O.synth_rows(9, 0, 64, 768)
    x[40] = x[3]
    x[50] = x[3] * 2.0            # same direction, different norm: cosine ties exactly after normalisation
    st = O.normalize_rows(x, "f32")
    ids, sc = coracle.cosine_topk(x[3:4], st, 3)
    assert ids[0, 0] == 3 and ids[0, 1] == 40 and sc[0, 0] == sc[0, 1]


def test_zero_query_and_zero_rows_score_zero(coracle):
    x = O.synth_rows(9, 0, 32, 64, zero_every=4)
    st = O.normalize_rows(x, "f32")
    ids, sc = coracle.cosine_topk(np.zeros((1, 64), np.float32), st, 5)
    assert ids[0].tolist() == [0, 1, 2, 3, 4] and not sc.any() and not np.signbit(sc).any()


def test_sharded_merge_equals_unsharded(coracle):
    st = O.normalize_rows(O.synth_rows(21, 0, 1000, 384, dup_every=9), "bf16")
    q = O.synth_rows(22, 0, 4, 384)
    full = coracle.cosine_topk(q, st, 10)
    parts = [coracle.cosine_topk(q, st[a:b], 10, id_base=a) for a, b in ((0, 250), (250, 500), (500, 750), (750, 1000))]
    mi, ms = O.merge_topk([p[0] for p in parts], [p[1] for p in parts], 10)
    assert np.array_equal(mi, full[0]) and np.array_equal(ms, full[1])


def test_golden_fixtures(coracle):
    """Fixtures written by scripts/make_golden.py (numpy oracle); both oracles must reproduce them."""
    with open(os.path.join(GOLDEN, "topk_cases.json")) as f:
        cases = json.load(f)
    assert len(cases["cases"]) >= 6
    for c in cases["cases"]:
        x = O.synth_rows(c["seed"], 0, c["n"], c["dim"], c["dup_every"], c["zero_every"])
        q = O.synth_rows(c["seed"] + 1, 0, c["nq"], c["dim"])
        for mod in (O, coracle):
            st = mod.normalize_rows(x, c["dtype"])
            ids, sc = mod.cosine_topk(q, st, c["k"])
            assert ids.tolist() == c["ids"], c["name"]
            assert sc.view(np.uint32).tolist() == c["score_bits"], c["name"]


def test_golden_chunk_collection(coracle):
    """The reference's 16-chunk collection shape (ids/periods/types from chunks.json) with
    deterministic stand-in embeddings (the MiniLM encoder is not available offline)."""
    with open(os.path.join(GOLDEN, "fin_chunks_collection.json")) as f:
        g = json.load(f)
    assert len(g["chunks"]) == 16 and g["dim"] == 384
    x = O.synth_rows(g["seed"], 0, 16, 384)
    st = coracle.normalize_rows(x, "f32")
    q = O.synth_rows(g["seed"] + 1, 0, len(g["queries"]), 384)
    ids, sc = coracle.cosine_topk(q, st, 3)
    for qi, want in enumerate(g["queries"]):
        assert [g["chunks"][i]["id"] for i in ids[qi]] == want["top3_ids"]
        assert sc[qi].view(np.uint32).tolist() == want["top3_score_bits"]
    assert math.isfinite(float(sc.max()))


def test_product_generator_matches_oracle_generator():
    from ragfin_b200.synthetic import synth_rows
    a = synth_rows(1234, 123456789, 700, 768, dup_every=7, zero_every=11)
    b = O.synth_rows(1234, 123456789, 700, 768, dup_every=7, zero_every=11)
    assert np.array_equal(a, b)


def test_fast_cpu_port_within_1e5_of_canonical(coracle):
    """The timed CPU baseline (faiss-style sgemv / blocked sgemm + topk) agrees with the canonical
    oracle: same index sets on tie-free data and scores within 1e-5 relative."""
    import torch
    from oracle import fast_cpu
    st = coracle.normalize_rows(O.synth_rows(31, 0, 70000, 384), "f32")
    for nq in (3, 24):
        q = O.synth_rows(32, 0, nq, 384)
        wi, ws = coracle.cosine_topk(q, st, 10)
        gi, gs = fast_cpu.fast_topk(torch.from_numpy(st), q, 10)
        assert np.array_equal(np.sort(gi, axis=1), np.sort(wi, axis=1))
        assert np.allclose(gs, ws, rtol=1e-5, atol=1e-7)


def test_oracle_agrees_with_third_party_brute_force_cosine():
    """Independent check of the restated semantics (the oracle is otherwise pinned only to itself): scikit-learn's
    brute-force cosine k-NN and scipy's cosine distance - third-party implementations of "cosine similarity, larger
    is better, k best in descending order" on RAW (un-normalised) fp32 rows, the way a COSINE collection is fed
    (reference "chunking_storing (1).py":29,383-396 and retrieve.py:28-34).  Same index lists on tie-free data,
    similarity = 1 - distance within fp32 accuracy."""
    from scipy.spatial.distance import cdist
    from sklearn.neighbors import NearestNeighbors
    rng = np.random.default_rng(7)
    x = (O.synth_rows(41, 0, 5000, 384) * rng.uniform(0.1, 30.0, size=(5000, 1))).astype(np.float32)   # norms all over the place
    q = (O.synth_rows(42, 0, 9, 384) * rng.uniform(0.5, 4.0, size=(9, 1))).astype(np.float32)
    k = 10
    wi, ws = O.cosine_topk(q, O.normalize_rows(x, "f32"), k)
    nn = NearestNeighbors(n_neighbors=k, metric="cosine", algorithm="brute").fit(x.astype(np.float64))
    dist, ind = nn.kneighbors(q.astype(np.float64))
    assert np.array_equal(ind, wi)
    assert np.allclose(1.0 - dist, ws, rtol=1e-5, atol=1e-6)
    d = cdist(q.astype(np.float64), x.astype(np.float64), metric="cosine")
    order = np.argsort(d, axis=1, kind="stable")[:, :k]
    assert np.array_equal(order, wi)
    assert np.allclose(1.0 - np.take_along_axis(d, order, axis=1), ws, rtol=1e-5, atol=1e-6)


def test_c_oracle_under_address_and_ub_sanitizers():
    """The checker itself is checked: oracle/ragfin_oracle.c + oracle/selftest.c built with -fsanitize=address,undefined
    and run over ragged dims, zero rows, duplicates, k > n, n = 0 and the threaded paths."""
    import subprocess
    odir = os.path.join(os.path.dirname(GOLDEN), os.pardir, "oracle")
    b = subprocess.run(["make", "-C", odir, "selftest"], capture_output=True, text=True)
    if b.returncode != 0:
        pytest.skip("no sanitizer runtime for this compiler: " + b.stderr.strip().splitlines()[-1][:200])
    r = subprocess.run([os.path.join(odir, "_san", "selftest")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ORACLE SELFTEST OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


import _thirdparty as TP   # noqa: E402


@pytest.mark.parametrize("case", TP.cases(), ids=lambda c: c["name"])
def test_oracle_pinned_by_committed_thirdparty_fixtures(coracle, case):
    """tests/golden/thirdparty_topk.json was written by scikit-learn (+ scipy) from the raw rows, without the oracle
    (scripts/make_thirdparty_golden.py): both oracle implementations must reproduce its id lists and similarities."""
    x, q = TP.raw_inputs(case["seed"], case["n"], case["dim"], case["nq"], case["scale_seed"])
    q = q[case["queries"]]
    ids_c, sc_c = coracle.cosine_topk(q, coracle.normalize_rows(x, "f32"), case["k"])
    TP.check(case, ids_c, sc_c)
    if case["n"] <= 5000:                                          # the numpy restatement is slow: small cases only
        ids_p, sc_p = O.cosine_topk(q, O.normalize_rows(x, "f32"), case["k"])
        assert np.array_equal(ids_p, ids_c) and np.array_equal(sc_p.view(np.uint32), sc_c.view(np.uint32))
