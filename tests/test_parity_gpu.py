"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs.  Index sets and scores must be BIT-EXACT (ids equal, fp32 score bit patterns equal)."""
import json
import os

import numpy as np
import pytest

from oracle import ragfin_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def _multi_kernel_paths(monkeypatch):
    """This module pins the multi-kernel paths (K2 scan, K3 / K3r sweeps, large k): handles created here do not take the
    one-kernel search, which has its own parity suite (tests/test_fused_gpu.py) and serves <= 64 queries by default."""
    monkeypatch.setenv("RAGFIN_NO_FUSED", "1")


def _index(x, dtype, capacity=None):
    import ragfin_b200
    idx = ragfin_b200.Index(x.shape[1], dtype, capacity=capacity or max(len(x), 1), device=0)
    if len(x):
        idx.add(x)
    return idx


def _assert_same(got, want, what=""):
    gi, gs = got
    wi, ws = want
    assert np.array_equal(gi, wi), f"{what}: ids differ\n{gi}\n{wi}"
    assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32)), f"{what}: score bits differ\n{gs}\n{ws}"


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("dim", [384, 768, 1024, 100, 33, 8])
def test_ingest_is_bit_exact(coracle, dtype, dim):
    x = O.synth_rows(3, 0, 777, dim, zero_every=17)
    x[5] *= 1000.0
    x[6] *= 1e-6
    idx = _index(x, dtype)
    got = idx.read_rows(0, 777)
    want = coracle.normalize_rows(x, dtype)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("dtype", O.DTYPES)
def test_synthetic_ingest_matches_host_generator(coracle, dtype):
    import ragfin_b200
    idx = ragfin_b200.Index(768, dtype, capacity=3000, device=0)
    idx.add_synthetic(1234, 0, 1000, dup_every=7, zero_every=11)
    idx.add_synthetic(1234, 1000, 2000, dup_every=7, zero_every=11)   # chunked adds continue the same matrix
    want = coracle.normalize_rows(O.synth_rows(1234, 0, 3000, 768, 7, 11), dtype)
    assert np.array_equal(idx.read_rows(0, 3000).view(np.uint32), want.view(np.uint32))
    assert len(idx) == 3000


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("dim", [384, 768, 1024])
@pytest.mark.parametrize("nq,k", [(1, 10), (2, 5), (3, 1), (4, 100), (9, 10)])
def test_search_matches_oracle(coracle, dtype, dim, nq, k):
    n = 20000 if dim == 768 else 6000
    x = O.synth_rows(40 + dim, 0, n, dim, dup_every=97, zero_every=1013)
    q = O.synth_rows(41 + dim, 0, nq, dim)
    idx = _index(x, dtype)
    got = idx.search(q, k)
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), k)
    _assert_same(got, want, f"{dtype} dim={dim} nq={nq} k={k}")
    assert idx.stats()["launches"] > 0


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("dim,n", [(768, 20001), (384, 5000), (100, 3000), (8, 257), (1024, 777), (2048, 1200), (768, 1)])
@pytest.mark.parametrize("nq", [1, 2])
def test_both_scan_kernels_match_the_oracle(coracle, dtype, dim, n, nq):
    """Batches of 1-2 queries run the HBM-bound scan: TMA-fed ring (default) or register-path loads."""
    x = O.synth_rows(45, 0, n, dim, dup_every=53, zero_every=509)
    q = O.synth_rows(46, 0, nq, dim)
    idx = _index(x, dtype)
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), 10)
    for variant in (2, 1, 0):
        idx.set_scan_variant(variant)
        got = idx.search(q, 10)
        assert idx.stats()["path"] == 0
        _assert_same(got, want, f"scan variant {variant} {dtype} dim={dim} n={n} nq={nq}")
    keep = np.zeros(n, bool)
    keep[::3] = True
    stored = coracle.normalize_rows(x, dtype)
    wi, ws = coracle.cosine_topk(q, stored[keep], 5)
    rows = np.flatnonzero(keep)
    wi = np.where(wi >= 0, rows[np.maximum(wi, 0)], -1)
    for variant in (2, 1):
        idx.set_scan_variant(variant)
        _assert_same(idx.search(q, 5, allow=keep), (wi, ws), f"filtered scan variant {variant}")


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("dim", [100, 33, 8, 2048, 1536])
def test_search_odd_and_wide_dims(coracle, dtype, dim):
    x = O.synth_rows(50, 0, 3000, dim)
    q = O.synth_rows(51, 0, 2, dim)
    got = _index(x, dtype).search(q, 7)
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), 7), f"{dtype} dim={dim}")


@pytest.mark.parametrize("n", [1, 2, 16, 31, 32, 33, 255, 257])
def test_small_collections_and_k_above_n(coracle, n):
    """The reference's real collection has 16 rows and is queried with limit up to 1000
    (graph_cons.py:279): every row comes back, padded slots are (-1, -inf)."""
    x = O.synth_rows(60, 0, n, 384)
    q = O.synth_rows(61, 0, 3, 384)
    idx = _index(x, "f32")
    for k in (1, 3, 20, 200):
        got = idx.search(q, k)
        want = coracle.cosine_topk(q, coracle.normalize_rows(x, "f32"), k)
        _assert_same(got, want, f"n={n} k={k}")
        if k > n:
            assert (got[0][:, n:] == -1).all() and np.isneginf(got[1][:, n:]).all()


def test_empty_collection():
    import ragfin_b200
    idx = ragfin_b200.Index(384, "bf16", capacity=8, device=0)
    ids, sc = idx.search(np.ones((2, 384), np.float32), 3)
    assert (ids == -1).all() and np.isneginf(sc).all()


@pytest.mark.parametrize("dtype", O.DTYPES)
def test_heavy_duplicates_take_the_exact_tier(coracle, dtype):
    """300 copies of the best row straddle the candidate guard band: the certificate must fail
    and the exact rescan tier must still return the lowest ids."""
    x = O.synth_rows(70, 0, 5000, 768)
    q = O.synth_rows(71, 0, 2, 768)
    x[100:400] = q[0] * 3.0
    x[4000:4100] = q[1]
    idx = _index(x, dtype)
    got = idx.search(q, 10)
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), 10), dtype)
    assert got[0][0].tolist() == list(range(100, 110))
    assert idx.stats()["queries_rescanned"] == 2


def test_all_rows_identical(coracle):
    x = np.tile(O.synth_rows(80, 0, 1, 384), (3000, 1))
    q = O.synth_rows(81, 0, 1, 384)
    got = _index(x, "bf16").search(q, 5)
    assert got[0][0].tolist() == [0, 1, 2, 3, 4]
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, "bf16"), 5))


def test_zero_rows_and_zero_query(coracle):
    x = O.synth_rows(90, 0, 2000, 384, zero_every=4)
    idx = _index(x, "f16")
    q = np.zeros((1, 384), np.float32)
    got = idx.search(q, 5)
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, "f16"), 5))
    assert got[0][0].tolist() == [0, 1, 2, 3, 4] and not got[1].any()


def test_incremental_add_and_id_base(coracle):
    x = O.synth_rows(95, 0, 9000, 768)
    q = O.synth_rows(96, 0, 4, 768)
    idx = _index(x[:1000], "bf16", capacity=9000)
    idx.add(x[1000:1001])
    idx.add(x[1001:])
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, "bf16"), 10)
    _assert_same(idx.search(q, 10), want)
    idx.set_id_base(5_000_000_000)
    got = idx.search(q, 10)
    assert np.array_equal(got[0], want[0] + 5_000_000_000)


def test_device_path_and_merge(coracle):
    import torch
    import ragfin_b200
    x = O.synth_rows(97, 0, 8000, 768, dup_every=5)
    q = O.synth_rows(98, 0, 5, 768)
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, "f16"), 10)
    qd = torch.from_numpy(q).cuda()
    shards, outs = [], []
    for a, b in ((0, 3000), (3000, 3001), (3001, 8000)):
        s = ragfin_b200.Index(768, "f16", capacity=b - a, device=0)
        s.add(torch.from_numpy(x[a:b]).cuda())
        s.set_id_base(a)
        shards.append(s)
        outs.append(s.search_device(qd, 10))
    ids = torch.cat([o[0] for o in outs], dim=1)
    sc = torch.cat([o[1] for o in outs], dim=1)
    mi, ms = ragfin_b200.merge_topk(ids, sc, 3, 10)
    torch.cuda.synchronize()
    _assert_same((mi.cpu().numpy(), ms.cpu().numpy()), want)


def test_golden_fixtures_on_gpu():
    with open(os.path.join(GOLDEN, "topk_cases.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        x = O.synth_rows(c["seed"], 0, c["n"], c["dim"], c["dup_every"], c["zero_every"])
        q = O.synth_rows(c["seed"] + 1, 0, c["nq"], c["dim"])
        ids, sc = _index(x, c["dtype"]).search(q, c["k"])
        assert ids.tolist() == c["ids"], c["name"]
        assert sc.view(np.uint32).tolist() == c["score_bits"], c["name"]


def test_errors_are_loud():
    import ragfin_b200
    idx = ragfin_b200.Index(384, "f32", capacity=4, device=0)
    with pytest.raises(ragfin_b200.RagfinError):
        idx.add(np.zeros((5, 384), np.float32))          # over capacity
    with pytest.raises(ValueError):
        idx.search(np.zeros((1, 383), np.float32), 3)    # wrong dim
    with pytest.raises(ValueError):
        idx.search(np.zeros((1, 384), np.float32), 0)    # top_k < 1


def test_reference_call_sites_through_the_shim(coracle):
    """Ingest + search exactly as "chunking_storing (1).py":11-29,376-417 and vector_rag_mcp/main.py:48-70
    do, through the pymilvus-shaped shim on the real engine."""
    from ragfin_b200 import milvus_compat as mc
    from ragfin_b200.vector_rag import HashingEncoder, VectorRAG
    with open(os.path.join(GOLDEN, "fin_chunks_collection.json")) as f:
        g = json.load(f)
    F, D = mc.FieldSchema, mc.DataType
    fields = [F("id", D.VARCHAR, max_length=100, is_primary=True), F("text", D.VARCHAR, max_length=4000),
              F("embedding", D.FLOAT_VECTOR, dim=384), F("period", D.VARCHAR, max_length=20),
              F("chunk_type", D.VARCHAR, max_length=30), F("statement_type", D.VARCHAR, max_length=30),
              F("primary_value", D.DOUBLE)]
    mc.connections.connect("default", host="localhost", port="19530")
    if mc.utility.has_collection("fin_chunks"):
        mc.utility.drop_collection("fin_chunks")
    col = mc.Collection("fin_chunks", mc.CollectionSchema(fields, "Financial complete context chunks"))
    col.create_index("embedding", {"index_type": "IVF_FLAT", "metric_type": "COSINE", "params": {"nlist": 128}})
    ch = g["chunks"]
    emb = O.synth_rows(g["seed"], 0, 16, 384)
    col.insert([[c["id"] for c in ch], [c["id"] + " text" for c in ch], emb.tolist(), [c["period"] for c in ch],
                [c["chunk_type"] for c in ch], ["consolidated"] * 16, [1.0] * 16])
    col.flush()
    col.load()
    assert col.num_entities == 16
    q = O.synth_rows(g["seed"] + 1, 0, 5, 384)
    for qi, want in enumerate(g["queries"]):
        hits = col.search(q[qi:qi + 1], "embedding", {"metric_type": "COSINE"}, 3, output_fields=["text", "period", "chunk_type"])[0]
        assert [h.id for h in hits] == want["top3_ids"]
        assert [np.float32(h.score).view(np.uint32).item() for h in hits] == want["top3_score_bits"]
    allhits = col.search(q[:1], "embedding", {"metric_type": "COSINE"}, limit=1000, output_fields=["id"])[0]
    assert len(allhits) == 16
    rag = VectorRAG(HashingEncoder(384), collection=col)
    env = rag.search_vectors("net profit Q1", 3)
    assert env["status"] == "success" and env["result_count"] == 3
    mc.utility.drop_collection("fin_chunks")


def test_chunk_ingest_pipeline_end_to_end(coracle):
    """N2: statements -> chunker -> encoder -> insert / flush / load -> search, as "chunking_storing (1).py":335-417 runs
    it (MiniLM replaced by the hashing stand-in: no weights offline).  Every chunk's own text must retrieve it first."""
    from ragfin_b200 import chunker, milvus_compat as mc
    from ragfin_b200.vector_rag import HashingEncoder, VectorRAG
    F, D = mc.FieldSchema, mc.DataType
    fields = [F("id", D.VARCHAR, max_length=100, is_primary=True), F("text", D.VARCHAR, max_length=4000),
              F("embedding", D.FLOAT_VECTOR, dim=384), F("period", D.VARCHAR, max_length=20),
              F("chunk_type", D.VARCHAR, max_length=30), F("statement_type", D.VARCHAR, max_length=30),
              F("primary_value", D.DOUBLE)]
    if mc.utility.has_collection("fin_chunks"):
        mc.utility.drop_collection("fin_chunks")
    col = mc.Collection("fin_chunks", mc.CollectionSchema(fields, "Financial complete context chunks"))
    col.create_index("embedding", {"index_type": "IVF_FLAT", "metric_type": "COSINE", "params": {"nlist": 128}})
    with open(os.path.join(GOLDEN, "fin_statements.json")) as f:
        chunks = chunker.build_corpus_from_bundle(json.load(f))
    enc = HashingEncoder(384)
    chunker.ingest_chunks(col, chunks, enc.encode)
    assert col.num_entities == 16
    emb = enc.encode([c["text"] for c in chunks])
    want_ids, want_sc = coracle.cosine_topk(emb, coracle.normalize_rows(emb, "f32"), 3)
    res = col.search(emb, "embedding", {"metric_type": "COSINE"}, 3, output_fields=["text", "period", "chunk_type"])
    for i, hits in enumerate(res):
        assert hits[0].id == chunks[i]["id"] and hits[0].entity.get("text") == chunks[i]["text"]
        assert [h.id for h in hits] == [chunks[j]["id"] for j in want_ids[i]]
        assert [np.float32(h.score).view(np.uint32).item() for h in hits] == want_sc[i].view(np.uint32).tolist()
    rag = VectorRAG(enc, collection=col)
    question = "ICICI Bank Limited Q3_FY2024 Balance Sheet Analysis total assets advances"
    top = rag.search(question, 1)[0]
    row = int(coracle.cosine_topk(enc.encode([question]), coracle.normalize_rows(emb, "f32"), 1)[0][0, 0])
    assert top["text"] == chunks[row]["text"] and top["period"] == chunks[row]["period"] == "Q3_FY2024" and top["rank"] == 1
    mc.utility.drop_collection("fin_chunks")


# ------------------------------------------------------------------------------------------------
# tensor-core path (query batches >= 9 rows)
# ------------------------------------------------------------------------------------------------
def _rounded_operands(idx, q, dtype):
    stored = idx.read_rows(0, len(idx))
    qhat = O.normalize_rows(q, "f32")
    if dtype == "f32":      # kind::tf32 reads the top 19 bits of each fp32 operand
        trunc = lambda a: (np.ascontiguousarray(a).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        return trunc(stored), trunc(qhat)
    return stored, O.round_to_storage(qhat, dtype)


@pytest.mark.parametrize("dtype,dim,n,nq", [("bf16", 768, 5000, 130), ("f16", 384, 777, 9), ("bf16", 100, 3000, 128),
                                             ("f32", 768, 2100, 40), ("f16", 1024, 9000, 257)])
def test_gemm_raw_scores_match_matmul(dtype, dim, n, nq):
    import torch
    x = O.synth_rows(5, 0, n, dim)
    q = O.synth_rows(6, 0, nq, dim)
    idx = _index(x, dtype)
    got = idx.debug_gemm_scores(torch.from_numpy(q).cuda()).cpu()
    rows, qs = _rounded_operands(idx, q, dtype)
    ref = (torch.from_numpy(qs).double() @ torch.from_numpy(rows).double().T).float()
    assert (got - ref).abs().max().item() < 2e-6      # exact products, fp32 accumulation


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("nq,k", [(9, 10), (128, 5), (129, 1), (300, 10), (64, 100)])
def test_gemm_search_matches_oracle(coracle, dtype, nq, k):
    x = O.synth_rows(140, 0, 30000, 768, dup_every=97, zero_every=1013)
    q = O.synth_rows(141, 0, nq, 768)
    idx = _index(x, dtype)
    got = idx.search(q, k)
    assert idx.stats()["path"] == 1
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), k), f"gemm {dtype} nq={nq} k={k}")


@pytest.mark.parametrize("dtype,dim", [("bf16", 384), ("f16", 100), ("f32", 1024), ("bf16", 33), ("f32", 8)])
def test_gemm_search_other_dims(coracle, dtype, dim):
    x = O.synth_rows(150, 0, 7001, dim)
    q = O.synth_rows(151, 0, 33, dim)
    idx = _index(x, dtype)
    got = idx.search(q, 10)
    assert idx.stats()["path"] == 1
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), 10), f"gemm {dtype} dim={dim}")


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("n,dim,nq,k", [(70000, 128, 17, 10), (70000, 128, 300, 10), (150000, 64, 130, 100), (40001, 256, 1030, 1),
                                        (260000, 32, 5, 224)])
def test_gemm_append_and_bound_modes_agree_with_the_oracle(coracle, dtype, n, dim, nq, k):
    """Large unfiltered corpora run the sample ("bound") pass and then the append-mode sweep (no lists, exact by
    construction).  Turning append mode off falls back to K' lists seeded by the bound pass; turning the bound pass
    off as well gives the plain list sweep.  All three must return the oracle's hits bit for bit."""
    x = O.synth_rows(190, 0, n, dim, dup_every=89, zero_every=2011)
    q = O.synth_rows(191, 0, nq, dim)
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), k)
    idx = _index(x, dtype)
    appended = idx.search(q, k)
    st = idx.stats()
    assert st["path"] == 1 and st["queries_rescanned"] == 0
    _assert_same(appended, want, f"append {dtype} n={n} nq={nq} k={k}")
    idx.set_append_mode(False)
    seeded = idx.search(q, k)
    _assert_same(seeded, want, f"bound+lists {dtype} n={n} nq={nq} k={k}")
    if k <= 100:    # beyond K' = 128 only append mode reaches the tensor cores; the fallback is the large-k path
        assert idx.stats()["path"] == 1 and idx.stats()["launches"] == st["launches"]
        idx.set_bound_pass(False)
        plain = idx.search(q, k)
        assert idx.stats()["launches"] == st["launches"] - 2 * ((nq + 4095) // 4096), "the bound pass did not run"
        _assert_same(plain, want, f"lists {dtype} n={n} nq={nq} k={k}")


@pytest.mark.parametrize("dtype", O.DTYPES)
@pytest.mark.parametrize("n,dim,nq,k", [(70000, 128, 1, 10), (70000, 128, 2, 1), (70000, 128, 16, 10), (150000, 64, 5, 100),
                                        (40000, 1024, 7, 10), (66000, 100, 3, 10), (35000, 768, 16, 5), (70000, 128, 17, 10)])
def test_gemm_swapped_roles_for_small_batches(coracle, dtype, n, dim, nq, k):
    """Variant 3: batches of <= 16 queries in append mode run gemm_rows_kernel (corpus rows are the MMA's M, the
    queries its N = 16); 17 queries fall back to the standard orientation.  Same bits either way."""
    x = O.synth_rows(196, 0, n, dim, dup_every=61, zero_every=1999)
    q = O.synth_rows(197, 0, nq, dim)
    if nq >= 3:
        x[4000:4030] = q[2] * 2.0        # 30 exact duplicates of one query: ties resolved by row id
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), k)
    idx = _index(x, dtype)
    idx.set_gemm_min_batch(1)
    idx.set_gemm_variant(3)
    got = idx.search(q, k)
    st = idx.stats()
    assert st["path"] == 1 and st["queries_rescanned"] == 0
    _assert_same(got, want, f"swapped roles {dtype} n={n} dim={dim} nq={nq} k={k}")
    idx.set_gemm_variant(1)
    _assert_same(idx.search(q, k), want, "standard orientation")


def test_gemm_bound_pass_with_heavy_duplicates(coracle):
    x = O.synth_rows(192, 0, 70000, 128)
    q = O.synth_rows(193, 0, 20, 128)
    x[100:400] = q[0] * 3.0          # 300 copies of the best row of query 0: certificate fails, exact tier answers
    x[30000:30040] = q[5]            # 40 copies: K' = 32 of them fill the list, ties broken by row id
    x[512::1024] = q[7]              # one copy in every fourth tile
    idx = _index(x, "bf16")
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, "bf16"), 10)
    got = idx.search(q, 10)
    assert idx.stats()["path"] == 1 and idx.stats()["queries_rescanned"] == 0   # append mode needs no certificate
    _assert_same(got, want)
    assert got[0][0].tolist() == list(range(100, 110))
    idx.set_append_mode(False)
    _assert_same(idx.search(q, 10), want)


def test_gemm_append_overflow_goes_to_the_exact_tier(coracle):
    """More duplicates of the best row than one query may rescore: the buffer overflows, tier 2 answers exactly."""
    x = O.synth_rows(194, 0, 70000, 64)
    q = O.synth_rows(195, 0, 6, 64)
    x[1000:4000] = q[2] * 0.5
    idx = _index(x, "f16")
    got = idx.search(q, 10)
    assert idx.stats()["path"] == 1 and idx.stats()["queries_rescanned"] == 1
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, "f16"), 10))
    assert got[0][2].tolist() == list(range(1000, 1010))


def test_gemm_and_scan_paths_agree_and_knob_works(coracle):
    x = O.synth_rows(160, 0, 20000, 768, dup_every=50)
    q = O.synth_rows(161, 0, 20, 768)
    idx = _index(x, "bf16")
    a = idx.search(q, 10)
    assert idx.stats()["path"] == 1
    idx.set_gemm_min_batch(1 << 30)
    b = idx.search(q, 10)
    assert idx.stats()["path"] == 0
    _assert_same(a, b, "gemm vs scan")
    _assert_same(a, coracle.cosine_topk(q, coracle.normalize_rows(x, "bf16"), 10))


def test_gemm_heavy_duplicates_take_the_exact_tier(coracle):
    x = O.synth_rows(170, 0, 6000, 768)
    q = O.synth_rows(171, 0, 12, 768)
    x[100:400] = q[0] * 3.0
    x[4000:4100] = q[5]
    idx = _index(x, "bf16")
    got = idx.search(q, 10)
    assert idx.stats()["path"] == 1 and idx.stats()["queries_rescanned"] == 2
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, "bf16"), 10))
    assert got[0][0].tolist() == list(range(100, 110))


def test_gemm_small_collection_and_k_above_n(coracle):
    x = O.synth_rows(180, 0, 16, 384)
    q = O.synth_rows(181, 0, 40, 384)
    idx = _index(x, "f32")
    got = idx.search(q, 20)
    assert idx.stats()["path"] == 1
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, "f32"), 20))


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("dtype,nq", [("bf16", 300), ("f16", 129), ("f32", 513), ("bf16", 40)])
def test_gemm_cluster_multicast_variants(coracle, cluster, dtype, nq):
    """Thread-block clusters with TMA multicast of the corpus tile: same bits for every cluster size, including
    query-tile counts that do not divide the cluster (padding tiles)."""
    x = O.synth_rows(190, 0, 26000, 768, dup_every=211)
    q = O.synth_rows(191, 0, nq, 768)
    idx = _index(x, dtype)
    idx.set_gemm_cluster(cluster)
    got = idx.search(q, 10)
    assert idx.stats()["path"] == 1
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), 10), f"cluster={cluster} {dtype} nq={nq}")


# ------------------------------------------------------------------------------------------------
# large-k path (k > 224 on corpora of more than 256 rows; Milvus allows limit up to 16384)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,n,k,nq", [("f32", 3000, 300, 2), ("bf16", 20000, 1000, 3), ("f16", 9000, 5000, 1),
                                          ("bf16", 700, 1000, 2), ("f32", 40000, 16384, 1)])
def test_large_k_matches_oracle(coracle, dtype, n, k, nq):
    x = O.synth_rows(200, 0, n, 384, dup_every=37, zero_every=501)
    q = O.synth_rows(201, 0, nq, 384)
    idx = _index(x, dtype)
    got = idx.search(q, k)
    assert idx.stats()["path"] == 2
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), k), f"large-k {dtype} n={n} k={k}")
    if k > n:
        assert (got[0][:, n:] == -1).all() and np.isneginf(got[1][:, n:]).all()


@pytest.mark.parametrize("dtype,dim,n,k,nq", [("bf16", 768, 60000, 1000, 1), ("bf16", 768, 60000, 1000, 16), ("f16", 384, 50000, 300, 19),
                                              ("f32", 384, 30000, 1000, 5), ("bf16", 100, 70000, 16384, 2), ("f32", 768, 9000, 9000, 3)])
def test_large_k_batched_pipeline_matches_oracle(coracle, dtype, dim, n, k, nq):
    """k > 256 on corpora of >= 4096 rows: tensor-core sweep dumping approximate scores of 16 queries at a time, radix select of
    the k-th approximate score, compaction of every row within 2 eps, exact rescore, rank sort - batched over the queries, one
    stream synchronisation per call.  Same bits as the oracle and as the one-query exact path."""
    import ragfin_b200
    x = O.synth_rows(210, 0, n, dim, dup_every=41, zero_every=997)
    q = O.synth_rows(211, 0, nq, dim)
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), k)
    idx = _index(x, dtype)
    got = idx.search(q, k)
    st = idx.stats()
    assert st["path"] == 2 and st["queries_rescanned"] == 0 and st["launches"] == 10 * ((nq + 15) // 16), st
    _assert_same(got, want, f"batched large-k {dtype} dim={dim} n={n} k={k} nq={nq}")
    idx.close()


def test_large_k_batched_pipeline_filters_and_duplicates(coracle, monkeypatch):
    import ragfin_b200
    n, dim, k = 40000, 128, 500
    x = O.synth_rows(220, 0, n, dim)
    x[1000:9000] = x[999]                       # 8000 identical rows at the top of one query: more candidates than the buffer
    q = np.concatenate([x[999:1000] * 3.0, O.synth_rows(221, 0, 2, dim)])
    idx = _index(x, "bf16")
    stored = coracle.normalize_rows(x, "bf16")
    got = idx.search(q, k)
    st = idx.stats()
    assert st["path"] == 2 and st["queries_rescanned"] == 1, st       # the flagged query took the exact one-query path
    _assert_same(got, coracle.cosine_topk(q, stored, k), "duplicates")
    assert list(got[0][0][:3]) == [999, 1000, 1001]
    allow = np.random.default_rng(5).random(n) < 0.3
    ids, sc = idx.search(q[1:], k, allow=allow)
    rows = np.flatnonzero(allow)
    wi, ws = coracle.cosine_topk(q[1:], stored[rows], k)
    assert np.array_equal(ids, rows[wi]) and np.array_equal(sc.view(np.uint32), ws.view(np.uint32))
    monkeypatch.setenv("RAGFIN_NO_BIGK_BATCHED", "1")           # the one-query exact path answers the same bits
    idx2 = _index(x, "bf16")
    _assert_same(idx2.search(q[1:], k), coracle.cosine_topk(q[1:], stored, k), "one-query path")
    assert idx2.stats()["launches"] == 40


def test_hybrid_limit_1000_through_the_shim_on_a_larger_collection(coracle):
    """graph_cons.py:275-281 asks limit=1000; on a collection larger than that it must return exactly 1000 ranked hits."""
    from ragfin_b200 import milvus_compat as mc
    F, D = mc.FieldSchema, mc.DataType
    mc.utility.drop_collection("big")
    col = mc.Collection("big", mc.CollectionSchema([F("id", D.VARCHAR, max_length=100, is_primary=True),
                                                     F("embedding", D.FLOAT_VECTOR, dim=384)]), storage_dtype="bf16")
    x = O.synth_rows(210, 0, 5000, 384)
    col.insert([[f"c{i}" for i in range(5000)], x])
    col.load()
    q = O.synth_rows(211, 0, 1, 384)
    hits = col.search(q, "embedding", {"metric_type": "COSINE"}, limit=1000, output_fields=["id"])[0]
    wi, ws = coracle.cosine_topk(q, coracle.normalize_rows(x, "bf16"), 1000)
    assert [h.entity.get("id") for h in hits] == [f"c{i}" for i in wi[0]]
    assert np.array_equal(np.array([h.score for h in hits], np.float32).view(np.uint32), ws[0].view(np.uint32))
    mc.utility.drop_collection("big")


def test_save_load_roundtrip_is_bit_identical(coracle, tmp_path):
    import ragfin_b200
    x = O.synth_rows(220, 0, 12345, 768, dup_every=17)
    q = O.synth_rows(221, 0, 12, 768)
    idx = _index(x, "bf16")
    idx.set_id_base(777)
    want = idx.search(q, 10)
    path = str(tmp_path / "corpus.ragfin")
    idx.save(path)
    idx.close()
    back = ragfin_b200.Index.load(path, capacity=20000, device=0)
    assert len(back) == 12345 and back.dim == 768 and back.dtype == "bf16"
    _assert_same(back.search(q, 10), want, "reloaded")
    assert np.array_equal(back.read_rows(0, 100).view(np.uint32), coracle.normalize_rows(x[:100], "bf16").view(np.uint32))
    back.add(x[:5])                    # still appendable up to the new capacity
    assert len(back) == 12350
    with pytest.raises(ragfin_b200.RagfinError):
        ragfin_b200.Index.load(str(tmp_path / "missing.ragfin"))
    (tmp_path / "junk.ragfin").write_bytes(b"not a matrix" * 10)
    with pytest.raises(ragfin_b200.RagfinError):
        ragfin_b200.Index.load(str(tmp_path / "junk.ragfin"))


def test_collection_save_and_reload_through_the_shim(tmp_path):
    from ragfin_b200 import milvus_compat as mc
    F, D = mc.FieldSchema, mc.DataType
    mc.utility.drop_collection("persist")
    col = mc.Collection("persist", mc.CollectionSchema([F("id", D.VARCHAR, max_length=100, is_primary=True),
                                                         F("embedding", D.FLOAT_VECTOR, dim=384), F("period", D.VARCHAR, max_length=20)]))
    x = O.synth_rows(230, 0, 300, 384)
    col.insert([[f"c{i}" for i in range(300)], x, [f"Q{i % 4 + 1}_FY2024" for i in range(300)]])
    col.flush()
    q = O.synth_rows(231, 0, 2, 384)
    before = [[(h.id, h.score, h.entity.period) for h in hits] for hits in col.search(q, "embedding", {"metric_type": "COSINE"}, 5, output_fields=["period"])]
    mc.save_collection("persist", str(tmp_path))
    mc.utility.drop_collection("persist")
    col2 = mc.load_collection("persist", str(tmp_path))
    assert col2.num_entities == 300 and mc.Collection("persist").num_entities == 300
    after = [[(h.id, h.score, h.entity.period) for h in hits] for hits in col2.search(q, "embedding", {"metric_type": "COSINE"}, 5, output_fields=["period"])]
    assert before == after
    assert col2.query(expr='id in ["c7"]', output_fields=["period"]) == [{"id": "c7", "period": "Q4_FY2024"}]
    mc.utility.drop_collection("persist")


@pytest.mark.parametrize("cluster", [1, 4])
@pytest.mark.parametrize("dtype,dim,nq", [("bf16", 768, 300), ("f16", 384, 40), ("bf16", 100, 129)])
def test_gemm_a_stationary_variant(coracle, cluster, dtype, dim, nq):
    """Experimental tcgen05 variant with the query tile resident in tensor memory (A from TMEM)."""
    x = O.synth_rows(240, 0, 21000, dim, dup_every=211)
    q = O.synth_rows(241, 0, nq, dim)
    idx = _index(x, dtype)
    idx.set_gemm_variant(2)
    idx.set_gemm_cluster(cluster)
    got = idx.search(q, 10)
    assert idx.stats()["path"] == 1
    _assert_same(got, coracle.cosine_topk(q, coracle.normalize_rows(x, dtype), 10), f"astat {dtype} dim={dim} nq={nq}")


def test_more_queries_than_one_pipeline_pass(coracle):
    """nq above the 4096-query workspace bound is processed in several passes."""
    x = O.synth_rows(250, 0, 3000, 128)
    q = O.synth_rows(251, 0, 4096 + 130, 128)
    idx = _index(x, "f16")
    got = idx.search(q, 3)
    want = coracle.cosine_topk(q, coracle.normalize_rows(x, "f16"), 3)
    _assert_same(got, want, "nq=4226")


# ------------------------------------------------------------------------------------------------
# scalar-filtered search (row bitmask consumed in the top-k epilogues)
# ------------------------------------------------------------------------------------------------
def _oracle_filtered(coracle, q, stored, allow, k):
    rows = np.flatnonzero(allow)
    ids, sc = coracle.cosine_topk(q, stored[rows], k) if len(rows) else (np.full((len(q), k), -1, np.int64), np.full((len(q), k), -np.inf, np.float32))
    return np.where(ids >= 0, rows[np.clip(ids, 0, None)] if len(rows) else -1, -1), sc


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("nq,k,keep", [(1, 10, 0.5), (3, 5, 0.01), (40, 10, 0.3), (2, 300, 0.2), (2, 10, 0.0), (300, 10, 0.9)])
def test_filtered_search_matches_oracle_on_allowed_rows(coracle, dtype, nq, k, keep):
    n = 12000
    x = O.synth_rows(260, 0, n, 384, dup_every=41)
    q = O.synth_rows(261, 0, nq, 384)
    allow = np.random.default_rng(5).random(n) < keep
    idx = _index(x, dtype)
    got = idx.search(q, k, allow=allow)
    want = _oracle_filtered(coracle, q, coracle.normalize_rows(x, dtype), allow, k)
    _assert_same(got, want, f"filtered {dtype} nq={nq} k={k} keep={keep}")
    valid = got[0][got[0] >= 0]
    assert allow[valid].all()


def test_filtered_search_through_the_shim():
    from ragfin_b200 import milvus_compat as mc
    F, D = mc.FieldSchema, mc.DataType
    mc.utility.drop_collection("filt")
    col = mc.Collection("filt", mc.CollectionSchema([F("id", D.VARCHAR, max_length=100, is_primary=True),
                                                      F("embedding", D.FLOAT_VECTOR, dim=384), F("period", D.VARCHAR, max_length=20)]))
    x = O.synth_rows(270, 0, 2000, 384)
    col.insert([[f"c{i}" for i in range(2000)], x, [f"Q{i % 4 + 1}_FY2024" for i in range(2000)]])
    col.load()
    q = O.synth_rows(271, 0, 1, 384)
    hits = col.search(q, "embedding", {"metric_type": "COSINE"}, 7, expr='period == "Q3_FY2024"', output_fields=["period"])[0]
    allhits = col.search(q, "embedding", {"metric_type": "COSINE"}, 2000, output_fields=["period"])[0]
    assert len(hits) == 7 and [h.id for h in hits] == [h.id for h in allhits if h.entity.period == "Q3_FY2024"][:7]
    mc.utility.drop_collection("filt")


def test_concurrent_searches_on_one_handle(coracle):
    """FastMCP runs sync tools on worker threads: concurrent search calls on one collection handle must
    serialise inside the library and each return its own, correct hits."""
    import threading
    x = O.synth_rows(210, 0, 80000, 128, dup_every=71)
    idx = _index(x, "bf16")
    stored = coracle.normalize_rows(x, "bf16")
    qs = [O.synth_rows(300 + t, 0, nq, 128) for t, nq in enumerate((1, 2, 7, 40))]
    want = [coracle.cosine_topk(q, stored, 10) for q in qs]
    errors = []

    def worker(t):
        try:
            for _ in range(15):
                got = idx.search(qs[t], 10)
                if not (np.array_equal(got[0], want[t][0]) and np.array_equal(got[1].view(np.uint32), want[t][1].view(np.uint32))):
                    errors.append(f"thread {t}: wrong hits")
                    return
        except Exception as e:   # noqa: BLE001
            errors.append(f"thread {t}: {e!r}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


import _thirdparty as TP   # noqa: E402


@pytest.mark.parametrize("case", TP.cases(), ids=lambda c: c["name"])
def test_engine_reproduces_committed_thirdparty_fixtures(case):
    """tests/golden/thirdparty_topk.json (written by scikit-learn + scipy from the raw rows, scripts/make_thirdparty_golden.py;
    the oracle takes no part): the engine with fp32 storage returns the same id lists, similarities within 1e-5 relative
    (north_star's fp32 tolerance); with bf16 / fp16 storage the fp32 answer is recalled completely once the candidates are
    rescored in fp32 (recall@k = 1.0 against an fp32 rescore: the 16-bit top-(k + margin) contains the fp32 top-k)."""
    x, q = TP.raw_inputs(case["seed"], case["n"], case["dim"], case["nq"], case["scale_seed"])
    q = q[case["queries"]]
    k = case["k"]
    idx = _index(x, "f32")
    ids, sc = idx.search(q, k)
    TP.check(case, ids, sc)
    idx.close()
    if case["strict"] and case["n"] >= 2000:
        want = np.asarray(case["ids"])
        for dtype, tol in (("bf16", 4e-3), ("f16", 5e-4)):         # stated tolerance of a 16-bit stored score vs fp32
            idx = _index(x, dtype)
            wide_ids, wide_sc = idx.search(q, min(4 * k, case["n"]))
            for r in range(len(q)):
                assert set(want[r].tolist()) <= set(wide_ids[r].tolist()), (case["name"], dtype, r)
                pos = {int(i): j for j, i in enumerate(wide_ids[r])}
                got = np.array([wide_sc[r][pos[int(i)]] for i in want[r]])
                assert np.allclose(got, np.asarray(case["sims"])[r], rtol=0, atol=tol), (case["name"], dtype)
            idx.close()


def test_reserve_grows_on_the_device_bit_identically(coracle):
    """ragfin_reserve: new allocation + device-to-device copy; stored bits, ids and later searches unchanged."""
    x = O.synth_rows(61, 0, 3000, 384, dup_every=13)
    q = O.synth_rows(62, 0, 5, 384)
    for dtype in O.DTYPES:
        idx = _index(x[:1000], dtype, capacity=1000)
        before = idx.search(q, 10)
        with pytest.raises(Exception):
            idx.add(x[1000:1001])                       # full
        idx.reserve(3000)
        assert len(idx) == 1000
        _assert_same(idx.search(q, 10), before, "after reserve")
        idx.add(x[1000:])
        want_rows = coracle.normalize_rows(x, dtype)
        assert np.array_equal(idx.read_rows(0, 3000).view(np.uint32), want_rows.view(np.uint32))
        _assert_same(idx.search(q, 10), coracle.cosine_topk(q, want_rows, 10), f"{dtype} grown")
        idx.close()


def test_device_feed_through_the_shim_and_the_chunk_pipeline(coracle):
    """N2 without the host round trip ("chunking_storing (1).py":379-396 with an encoder that lives on the GPU):
    Collection.insert and ingest_chunks accept CUDA tensors, which reach K1 through ragfin_add(src_is_device=1); the
    collection grows on the device.  Same bits as the host feed."""
    import torch
    from ragfin_b200 import milvus_compat as mc
    from ragfin_b200.chunker import FIELD_ORDER, ingest_chunks
    F, D = mc.FieldSchema, mc.DataType
    fields = [F("id", D.VARCHAR, max_length=100, is_primary=True), F("text", D.VARCHAR, max_length=4000),
              F("embedding", D.FLOAT_VECTOR, dim=384), F("period", D.VARCHAR, max_length=20),
              F("chunk_type", D.VARCHAR, max_length=30), F("statement_type", D.VARCHAR, max_length=30),
              F("primary_value", D.DOUBLE)]
    x = O.synth_rows(71, 0, 300, 384)
    chunks = [dict(id=f"c{i}", text=f"text {i}", period="Q1_FY2024", chunk_type="t", statement_type="s", primary_value=float(i))
              for i in range(300)]

    cols = {}
    for feed in ("host", "device"):
        name = f"feed_{feed}"
        mc.utility.drop_collection(name)
        col = mc.Collection(name, mc.CollectionSchema(fields, "x"), storage_dtype="bf16", initial_capacity=64)
        col.create_index("embedding", {"index_type": "IVF_FLAT", "metric_type": "COSINE", "params": {"nlist": 128}})
        seen = []

        def encode(texts, feed=feed, seen=seen):
            rows = x[[int(t.split()[1]) for t in texts]]
            seen.append(len(texts))
            return torch.from_numpy(rows).cuda() if feed == "device" else rows.tolist()

        for r0 in range(0, 300, 100):                    # three ingests: the collection grows 64 -> 128 -> 256 -> 512
            ingest_chunks(col, chunks[r0:r0 + 100], encode)
        assert col.num_entities == 300 and seen == [100, 100, 100]
        cols[feed] = col
    a, b = cols["host"]._st.index, cols["device"]._st.index
    want = coracle.normalize_rows(x, "bf16")
    assert np.array_equal(a.read_rows(0, 300).view(np.uint32), want.view(np.uint32))
    assert np.array_equal(b.read_rows(0, 300).view(np.uint32), want.view(np.uint32))
    q = O.synth_rows(72, 0, 3, 384)
    ra = cols["host"].search(q, "embedding", {"metric_type": "COSINE"}, 5, output_fields=["text"])
    rb = cols["device"].search(q, "embedding", {"metric_type": "COSINE"}, 5, output_fields=["text"])
    wi, ws = coracle.cosine_topk(q, want, 5)
    for qi in range(3):
        assert [h.id for h in ra[qi]] == [h.id for h in rb[qi]] == [f"c{i}" for i in wi[qi]]
        assert [h.score for h in ra[qi]] == [h.score for h in rb[qi]] == [float(v) for v in ws[qi]]
    for n in ("feed_host", "feed_device"):
        mc.utility.drop_collection(n)


# ---- variants first run on a GPU at the start of round 2 (profiles/r02): 2-SM MMA pairs, views -------------------------
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


@pytest.mark.parametrize("dtype,dim", [("bf16", 768), ("f16", 768), ("bf16", 1024), ("bf16", 2048), ("f16", 384)])
def test_tensor_core_error_stays_inside_the_allowance(dtype, dim):
    """The accumulation allowance of the tensor-core paths ((ld + 64) * 2^-22 * 17/16 + 2^-21, csrc/ragfin_api.cu
    eps_gemm_const) against the measured |tensor-core score - exact score| on the operands that maximise it: every product
    positive, so the partial sums are as large as they can be (|x| . |q| ~ 0.6-0.8), plus random-sign operands.  The
    tensor core multiplies the STORED rows by the 16-bit-rounded query, so the exact reference uses those operands (the
    query rounding is a separate, rigorously bounded term).  Required: max error <= allowance / 4."""
    import torch
    ld = (dim + 7) // 8 * 8
    allowance = (ld + 64) * 2.0 ** -22 * 1.0625 + 2.0 ** -21
    worst = 0.0
    for positive in (True, False):
        x = O.synth_rows(510, 0, 4000, dim)
        q = O.synth_rows(511, 0, 130, dim)
        if positive:
            x, q = np.abs(x), np.abs(q)
        idx = _index(x, dtype)
        got = idx.debug_gemm_scores(torch.from_numpy(q).cuda()).cpu().numpy().astype(np.float64)
        stored = idx.read_rows(0, len(x)).astype(np.float64)
        q16 = O.round_to_storage(O.normalize_rows(q, "f32"), dtype).astype(np.float64)
        worst = max(worst, float(np.abs(got - q16 @ stored.T).max()))
        idx.close()
    assert worst <= allowance / 4, (worst, allowance)
