"""CPU suite for the reference-shaped host layer (pymilvus-compatible shim + VectorRAG mirror).
The engine is replaced by an oracle-backed stand-in through the shim's index_factory hook, so the
host logic (columns, hits, query expressions, envelopes) is what is under test here."""
import json
import os

import numpy as np
import pytest

from oracle import ragfin_oracle as O
from ragfin_b200 import milvus_compat as mc
from ragfin_b200.vector_rag import HashingEncoder, VectorRAG, merge_hybrid, validate_search_request

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class OracleIndex:
    """Stand-in with the `Index` surface the shim uses (add / search / close)."""

    def __init__(self, dim, dtype, capacity, device):
        self.dim, self.dtype, self.capacity = dim, dtype, capacity
        self.rows = np.zeros((0, dim), np.float32)
        self.closed = False

    def add(self, rows):
        assert len(self.rows) + len(rows) <= self.capacity
        self.rows = np.concatenate([self.rows, O.normalize_rows(rows, self.dtype)])

    def reserve(self, capacity):
        self.capacity = max(self.capacity, capacity)

    def search(self, q, k, allow=None):
        if allow is None:
            return O.cosine_topk(q, self.rows, k)
        rows = np.flatnonzero(allow)
        ids, sc = O.cosine_topk(q, self.rows[rows], k)
        return np.where(ids >= 0, rows[np.clip(ids, 0, None)] if len(rows) else -1, -1), sc

    def close(self):
        self.closed = True


def reference_fields(dim=384):
    F, D = mc.FieldSchema, mc.DataType
    return [F("id", D.VARCHAR, max_length=100, is_primary=True), F("text", D.VARCHAR, max_length=4000),
            F("embedding", D.FLOAT_VECTOR, dim=dim), F("period", D.VARCHAR, max_length=20),
            F("chunk_type", D.VARCHAR, max_length=30), F("statement_type", D.VARCHAR, max_length=30),
            F("primary_value", D.DOUBLE)]


def build_collection(name="fin_chunks", factory=OracleIndex, **kw):
    """The reference's ingest script, "chunking_storing (1).py":11-29 and 376-396, on the golden chunk metadata."""
    with open(os.path.join(GOLDEN, "fin_chunks_collection.json")) as f:
        g = json.load(f)
    mc.connections.connect("default", host="localhost", port="19530")
    if mc.utility.has_collection(name):
        mc.utility.drop_collection(name)
    col = mc.Collection(name, mc.CollectionSchema(reference_fields(), "Financial complete context chunks"),
                        index_factory=factory, **kw)
    col.create_index("embedding", {"index_type": "IVF_FLAT", "metric_type": "COSINE", "params": {"nlist": 128}})
    chunks = g["chunks"]
    texts = [f"{c['period']} {c['chunk_type']} text of {c['id']}" for c in chunks]
    emb = O.synth_rows(g["seed"], 0, 16, 384)
    data = [[c["id"] for c in chunks], texts, emb.tolist(), [c["period"] for c in chunks],
            [c["chunk_type"] for c in chunks], ["consolidated"] * 16, [float(i) for i in range(16)]]
    mr = col.insert(data)
    assert mr.insert_count == 16
    assert col.num_entities == 0          # pymilvus counts flushed rows
    col.flush()
    col.load()
    return col, g


def test_ingest_and_search_like_the_reference():
    col, g = build_collection()
    assert col.num_entities == 16
    q = O.synth_rows(g["seed"] + 1, 0, 5, 384)
    for qi, want in enumerate(g["queries"]):
        res = col.search(q[qi:qi + 1], "embedding", {"metric_type": "COSINE"}, 3,
                         output_fields=["text", "period", "chunk_type"])
        hits = res[0]
        assert [h.id for h in hits] == want["top3_ids"]
        assert [np.float32(h.score).view(np.uint32).item() for h in hits] == want["top3_score_bits"]
        assert hits[0].distance == hits[0].score and hits[0].entity.period.startswith("Q")
        assert hits[0].entity.get("chunk_type") == hits[0].entity.chunk_type
        with pytest.raises(mc.MilvusException):
            hits[0].entity.statement_type     # not requested in output_fields


def test_limit_keyword_and_limit_above_n_returns_all_rows():
    col, g = build_collection()
    q = O.synth_rows(g["seed"] + 1, 0, 1, 384)
    res = col.search(q, "embedding", {"metric_type": "COSINE"}, limit=1000, output_fields=["id", "text", "period", "chunk_type"])
    assert len(res) == 1 and len(res[0]) == 16
    sc = [h.score for h in res[0]]
    assert sc == sorted(sc, reverse=True)
    assert sorted(h.entity.get("id") for h in res[0]) == sorted(c["id"] for c in g["chunks"])


def test_search_argument_errors():
    col, g = build_collection()
    q = np.zeros((1, 384), np.float32)
    for bad in (0, 16385, "3"):
        with pytest.raises(mc.MilvusException):
            col.search(q, "embedding", {"metric_type": "COSINE"}, bad)
    with pytest.raises(mc.MilvusException):
        col.search(q, "embedding", {"metric_type": "L2"}, 3)
    with pytest.raises(mc.MilvusException):
        col.search(q, "vector", {"metric_type": "COSINE"}, 3)
    with pytest.raises(mc.MilvusException):
        col.search(np.zeros((1, 383), np.float32), "embedding", {"metric_type": "COSINE"}, 3)
    with pytest.raises(mc.MilvusException):
        col.search(q, "embedding", {"metric_type": "COSINE"}, 3, output_fields=["nope"])
    with pytest.raises(mc.MilvusException):
        col.create_index("embedding", {"metric_type": "L2"})


def test_query_expressions():
    col, g = build_collection()
    rows = col.query(expr="", limit=3, output_fields=["id", "period", "chunk_type", "statement_type"])   # test_vector.py:35-39
    assert [r["id"] for r in rows] == [c["id"] for c in g["chunks"][:3]]
    ids = [g["chunks"][5]["id"], g["chunks"][2]["id"], "missing"]
    expr = f"id in {str(ids)}".replace("'", '"')                                                        # graph_cons.py:306-311
    rows = col.query(expr=expr, output_fields=["id", "text", "period", "chunk_type"])
    assert [r.get("id") for r in rows] == [g["chunks"][2]["id"], g["chunks"][5]["id"]]
    rows = col.query(expr='period == "Q2_FY2024"', output_fields=["chunk_type"])
    assert len(rows) == 4
    with pytest.raises(mc.MilvusException):
        col.query(expr="")
    with pytest.raises(mc.MilvusException):
        col.query(expr="id like 'x%'")


def test_open_existing_and_missing_collection():
    build_collection("fin_chunks")
    again = mc.Collection("fin_chunks")
    again.load()
    assert again.num_entities == 16
    with pytest.raises(mc.SchemaNotReadyException):
        mc.Collection("does_not_exist")
    mc.utility.drop_collection("fin_chunks")
    assert not mc.utility.has_collection("fin_chunks")


def test_growth_on_the_device_and_duplicate_pk():
    made = []

    def factory(*a):
        made.append(OracleIndex(*a))
        return made[-1]

    mc.utility.drop_collection("grow")
    col = mc.Collection("grow", mc.CollectionSchema(reference_fields(64)), index_factory=factory, initial_capacity=4)
    x = O.synth_rows(5, 0, 11, 64)
    for a, b in ((0, 3), (3, 4), (4, 11)):
        col.insert([[f"c{i}" for i in range(a, b)], ["t"] * (b - a), x[a:b], ["p"] * (b - a), ["k"] * (b - a),
                    ["s"] * (b - a), [0.0] * (b - a)])
        col.flush()
    assert len(made) == 1 and not made[0].closed and made[0].capacity == 16      # grown in place (Index.reserve), never rebuilt
    assert np.array_equal(made[0].rows, O.normalize_rows(x, "f32"))
    with pytest.raises(mc.MilvusException):
        col.insert([["c3"], ["t"], x[:1], ["p"], ["k"], ["s"], [0.0]])
    res = col.search(x[7:8] * 3.0, "embedding", {"metric_type": "COSINE"}, 1)
    assert res[0][0].id == "c7" and abs(res[0][0].score - 1.0) < 1e-6


def test_vector_rag_mirror_and_tool_envelope():
    col, g = build_collection()
    rag = VectorRAG(HashingEncoder(384), collection=col)
    ctx = rag.search("net profit Q1", top_k=3)                                  # test_vector.py:97-100
    assert [c["rank"] for c in ctx] == [1, 2, 3]
    assert set(ctx[0]) == {"rank", "text", "period", "chunk_type", "statement_type", "primary_value", "score"}
    assert isinstance(ctx[0]["score"], float) and ctx[0]["score"] >= ctx[1]["score"] >= ctx[2]["score"]
    env = rag.search_vectors("net profit Q1", 3)
    assert env["status"] == "success" and env["result_count"] == 3 and env["results"] == ctx
    bad = rag.search_vectors("net profit Q1", 0)
    assert bad["status"] == "error" and "limit" in bad["message"] and bad["query"] == "net profit Q1"
    assert rag.get_collection_stats() == {"status": "success", "collection_name": "fin_chunks", "total_chunks": 16}
    tup = rag.retrieve_contexts("net profit Q1", top_k=5)
    assert len(tup) == 5 and tup[0][0] == ctx[0]["text"]
    hv = rag.hybrid_vector_chunks("net profit Q1")
    assert len(hv) == 16 and hv[0]["id"] in [c["id"] for c in g["chunks"]]
    merged = merge_hybrid(hv[:2], [{"id": hv[1]["id"], "score": 1.0}, {"id": "g1", "score": 1.0}])
    assert [m["id"] for m in merged] == [hv[0]["id"], hv[1]["id"], "g1"]


def test_rest_request_bounds():
    assert validate_search_request("net profit Q1") == 3
    for q, k in (("abc", 3), ("net profit", 0), ("net profit", 21)):
        with pytest.raises(ValueError):
            validate_search_request(q, k)


def test_filtered_search_expr():
    col, g = build_collection()
    q = O.synth_rows(g["seed"] + 1, 0, 1, 384)
    hits = col.search(q, "embedding", {"metric_type": "COSINE"}, 10, expr='period == "Q2_FY2024"', output_fields=["period"])[0]
    assert len(hits) == 4 and all(h.entity.period == "Q2_FY2024" for h in hits)
    allhits = col.search(q, "embedding", {"metric_type": "COSINE"}, 16, output_fields=["period"])[0]
    assert [h.id for h in hits] == [h.id for h in allhits if h.entity.period == "Q2_FY2024"]
    ids = [g["chunks"][3]["id"], g["chunks"][9]["id"]]
    hits = col.search(q, "embedding", {"metric_type": "COSINE"}, 5, expr=f"id in {ids}".replace("'", '"'))[0]
    assert sorted(h.id for h in hits) == sorted(ids)
    assert col.search(q, "embedding", {"metric_type": "COSINE"}, 5, expr='period == "none"')[0] == []
    with pytest.raises(mc.MilvusException):
        col.search(q, "embedding", {"metric_type": "COSINE"}, 5, expr="period like 'Q%'")


def test_failed_insert_leaves_the_collection_unchanged():
    """A Milvus insert is all-or-nothing: an over-long VARCHAR in a LATER column, a duplicate primary key or a wrong
    embedding shape must not leave earlier columns one row longer (every later hit would carry the wrong fields)."""
    col, g = build_collection("atomic_ins")
    st = col._st
    before = {k: list(v) for k, v in st.columns.items()}
    n0, pk0 = st.n_inserted, dict(st.pk_to_row)
    e = O.synth_rows(5, 0, 1, 384)
    bad = [
        [["a"], ["hello"], e.tolist(), ["p" * 21], ["t"], ["s"], [1.0]],                 # period exceeds max_length 20
        [[g["chunks"][0]["id"]], ["hello"], e.tolist(), ["ok"], ["t"], ["s"], [1.0]],    # duplicate primary key
        [["a", "a"], ["x", "y"], np.concatenate([e, e]).tolist(), ["ok", "ok"], ["t", "t"], ["s", "s"], [1.0, 2.0]],   # duplicate inside the insert
        [["a"], ["hello"], e[:, :100].tolist(), ["ok"], ["t"], ["s"], [1.0]],            # wrong embedding width
    ]
    for data in bad:
        with pytest.raises(mc.MilvusException):
            col.insert(data)
        assert st.columns == before and st.n_inserted == n0 and st.pk_to_row == pk0 and not st.pending
    mr = col.insert([["a"], ["hello"], e.tolist(), ["ok"], ["t"], ["s"], [1.0]])
    col.flush()
    assert mr.primary_keys == ["a"] and col.num_entities == 17
    assert col.query('id in ["a"]', output_fields=["text", "period"]) == [{"id": "a", "text": "hello", "period": "ok"}]


def test_unsupported_expressions_are_rejected_not_swallowed():
    col, g = build_collection("expr_strict")
    for expr in ['chunk_type == "x" and period == "Q1"', "period == Q1", 'id in ["a" "b"]', "primary_value > 3",
                 'not period == "x"', 'id in ["a",]', 'period == "x" or period == "y"']:
        with pytest.raises(mc.MilvusException):
            col.query(expr)
        with pytest.raises(mc.MilvusException):
            col.search(O.synth_rows(1, 0, 1, 384), "embedding", {"metric_type": "COSINE"}, 3, expr=expr)
    pid = g["chunks"][3]["period"]
    rows = col.query(f'period == "{pid}"', output_fields=["period"])
    assert rows and all(r["period"] == pid for r in rows)
    assert col.query("primary_value == 2.0", output_fields=["primary_value"])[0]["primary_value"] == 2.0
    assert col.query("id in []") == []


def test_collection_grows_without_a_host_copy_of_the_embeddings():
    """Capacity doubling goes through Index.reserve (device-side growth): the shim holds no second copy of the rows and
    nothing is re-ingested; results equal one big insert."""
    calls = []

    class Counting(OracleIndex):
        def add(self, rows):
            calls.append(len(rows))
            super().add(rows)

        def reserve(self, capacity):
            calls.append(("reserve", capacity))
            super().reserve(capacity)

    name = "grow"
    if mc.utility.has_collection(name):
        mc.utility.drop_collection(name)
    col = mc.Collection(name, mc.CollectionSchema(reference_fields(64), "g"), index_factory=Counting, initial_capacity=8)
    x = O.synth_rows(9, 0, 40, 64)
    for r0 in range(0, 40, 10):
        col.insert([[f"k{r0 + i}" for i in range(10)], ["t"] * 10, x[r0:r0 + 10], ["p"] * 10, ["c"] * 10, ["s"] * 10, [0.0] * 10])
        col.flush()
    assert not hasattr(col._st, "raw")
    assert [c for c in calls if not isinstance(c, tuple)] == [10, 10, 10, 10]        # every row ingested exactly once
    assert [c for c in calls if isinstance(c, tuple)] == [("reserve", 32), ("reserve", 64)]
    res = col.search(x[17:18], "embedding", {"metric_type": "COSINE"}, 3)
    assert res[0][0].id == "k17"


def test_scalar_filter_index_stays_current_and_equals_a_column_scan():
    """`field == literal` / `field in [...]` on a scalar field are answered from a value -> rows index built on first use and
    kept current by later inserts (a filtered search costs the matching rows, not a pass over the column).  Checked against
    a plain scan of the columns: ascending rows, literals listed twice, numeric fields, rows inserted after the first
    filter, unflushed rows invisible to query(), and a failed insert leaving the index untouched."""
    import random
    mc.connections.connect("default", host="localhost", port="19530")
    if mc.utility.has_collection("scalars"):
        mc.utility.drop_collection("scalars")
    F, D = mc.FieldSchema, mc.DataType
    fields = [F("id", D.VARCHAR, max_length=20, is_primary=True), F("embedding", D.FLOAT_VECTOR, dim=8),
              F("period", D.VARCHAR, max_length=8), F("bucket", D.INT64)]
    col = mc.Collection("scalars", mc.CollectionSchema(fields, ""), index_factory=OracleIndex)
    rng = random.Random(7)
    periods, total = ["Q1", "Q2", "Q3", "Q4"], 0

    def insert(n):
        nonlocal total
        rows = [[f"r{total + i}" for i in range(n)], O.synth_rows(40 + total, 0, n, 8).tolist(),
                [rng.choice(periods) for _ in range(n)], [rng.randrange(5) for _ in range(n)]]
        total += n
        col.insert(rows)

    def scan(pred):
        st = col._st
        return [st.columns["id"][r] for r in range(col.num_entities) if pred(st.columns["period"][r], st.columns["bucket"][r])]

    insert(300)
    col.flush()
    assert [r["id"] for r in col.query(expr='period == "Q3"')] == scan(lambda p, b: p == "Q3")
    assert "period" in col._st.scalar_index and "bucket" not in col._st.scalar_index
    insert(200)                                                          # pending: indexed, but not yet visible to query()
    assert [r["id"] for r in col.query(expr='period in ["Q1", "Q4", "Q1"]')] == scan(lambda p, b: p in ("Q1", "Q4"))
    assert len(col.query(expr='period in ["Q1", "Q4", "Q1"]')) < sum(p in ("Q1", "Q4") for p in col._st.columns["period"])
    col.flush()
    assert [r["id"] for r in col.query(expr='period in ["Q1", "Q4"]')] == scan(lambda p, b: p in ("Q1", "Q4"))
    assert [r["id"] for r in col.query(expr="bucket in [0, 3]")] == scan(lambda p, b: b in (0, 3))
    assert [r["id"] for r in col.query(expr="bucket == 4")] == scan(lambda p, b: b == 4)
    assert col.query(expr='period == "Q9"') == [] and col.query(expr="bucket in []") == []
    before = {f: {v: list(r) for v, r in inv.items()} for f, inv in col._st.scalar_index.items()}
    with pytest.raises(mc.MilvusException):
        col.insert([["x1", "r0"], O.synth_rows(1, 0, 2, 8).tolist(), ["Q1", "Q2"], [1, 2]])      # duplicate primary key
    assert col._st.scalar_index == before
    q = O.synth_rows(99, 0, 1, 8)
    hits = col.search(q, "embedding", {"metric_type": "COSINE"}, 500, expr="bucket == 2", output_fields=["bucket"])[0]
    assert sorted(h.id for h in hits) == sorted(scan(lambda p, b: b == 2)) and all(h.entity.bucket == 2 for h in hits)


def test_hit_entity_is_a_view_of_the_requested_fields_only():
    """hit.entity resolves the requested output fields when they are read (rows are append-only): requested fields by
    attribute and .get(), anything else raises / returns the default exactly as a pymilvus entity does, values survive
    later inserts and a drop of the collection, round_decimal rounds the score, k > N returns N hits."""
    import copy
    col, g = build_collection()
    q = O.synth_rows(g["seed"] + 1, 0, 2, 384)
    res = col.search(q, "embedding", {"metric_type": "COSINE"}, 50, output_fields=["id", "period"])
    assert len(res) == 2 and len(res[0]) == 16 and len(res[1]) == 16            # padded slots never become hits
    h = res[0][0]
    row = [c["id"] for c in g["chunks"]].index(h.id)
    assert h.entity.id == h.id and h.entity.period == g["chunks"][row]["period"] == h.entity.get("period")
    assert h.entity.get("text") is None and h.entity.get("text", "dflt") == "dflt"
    with pytest.raises(mc.MilvusException):
        h.entity.text
    assert h.entity.to_dict() == {"id": h.id, "period": g["chunks"][row]["period"]} == h.entity.fields
    assert h.to_dict() == {"id": h.id, "distance": h.distance, "entity": h.entity.to_dict()} and h.score == h.distance
    assert copy.copy(h.entity).period == h.entity.period and "period" in repr(h)
    assert mc.Entity({"a": 1, "b": "x"}).b == "x" and mc.Entity({"a": 1}).get("zz", 5) == 5     # detached record
    scores = [x.distance for x in res[0]]
    assert scores == sorted(scores, reverse=True) and res[0].ids == [x.id for x in res[0]] and res[0].distances == scores
    rounded = col.search(q[:1], "embedding", {"metric_type": "COSINE"}, 3, round_decimal=2)[0]
    assert [x.distance for x in rounded] == [round(s_, 2) for s_ in scores[:3]]
    col.insert([["late"], ["late text"], O.synth_rows(77, 0, 1, 384).tolist(), ["Q9"], ["t"], ["s"], [1.0]])
    col.flush()
    assert h.entity.period == g["chunks"][row]["period"]
    mc.utility.drop_collection(col.name)
    assert h.entity.period == g["chunks"][row]["period"]


def test_concurrent_inserts_and_searches_on_one_collection():
    """FastMCP runs its sync tools on worker threads and an ingest may run beside them: inserts, flushes,
    plain and filtered searches and scalar queries from several threads on ONE collection.  Every search must see a
    consistent prefix of the rows (pk, columns, scalar index and embeddings in step), and the end state must equal a
    sequential ingest."""
    import threading
    mc.connections.connect("default", host="localhost", port="19530")
    if mc.utility.has_collection("mt"):
        mc.utility.drop_collection("mt")
    F, D = mc.FieldSchema, mc.DataType
    fields = [F("id", D.VARCHAR, max_length=20, is_primary=True), F("embedding", D.FLOAT_VECTOR, dim=16), F("tag", D.VARCHAR, max_length=4)]
    col = mc.Collection("mt", mc.CollectionSchema(fields, ""), index_factory=OracleIndex)
    n_writers, per_writer, block = 3, 20, 7
    x = O.synth_rows(500, 0, n_writers * per_writer * block, 16)
    q = O.synth_rows(501, 0, 2, 16)
    errors, stop = [], threading.Event()

    def writer(w):
        try:
            for b in range(per_writer):
                r0 = (w * per_writer + b) * block
                col.insert([[f"r{r0 + i}" for i in range(block)], x[r0:r0 + block].tolist(), [f"t{(r0 + i) % 3}" for i in range(block)]])
                if b % 4 == 3:
                    col.flush()
        except Exception as e:   # noqa: BLE001
            errors.append(f"writer {w}: {e!r}")

    def reader(t):
        try:
            while not stop.is_set():
                for hits in col.search(q, "embedding", {"metric_type": "COSINE"}, 5, output_fields=["id", "tag"]):
                    for h in hits:
                        row = int(h.id[1:])
                        if h.entity.id != h.id or h.entity.tag != f"t{row % 3}":
                            errors.append(f"reader {t}: hit {h.id} carries the columns of another row")
                    if [h.distance for h in hits] != sorted((h.distance for h in hits), reverse=True):
                        errors.append(f"reader {t}: hits out of order")
                for h in col.search(q[:1], "embedding", {"metric_type": "COSINE"}, 4, expr='tag == "t1"', output_fields=["tag"])[0]:
                    if h.entity.tag != "t1" or int(h.id[1:]) % 3 != 1:
                        errors.append(f"reader {t}: filtered hit {h.id} has tag {h.entity.tag}")
                for r in col.query(expr='tag in ["t0", "t2"]', output_fields=["tag"]):
                    if r["tag"] != f"t{int(r['id'][1:]) % 3}":
                        errors.append(f"reader {t}: query row {r}")
        except Exception as e:   # noqa: BLE001
            errors.append(f"reader {t}: {e!r}")

    readers = [threading.Thread(target=reader, args=(t,)) for t in range(3)]
    writers = [threading.Thread(target=writer, args=(w,)) for w in range(n_writers)]
    for th in readers + writers:
        th.start()
    for th in writers:
        th.join()
    stop.set()
    for th in readers:
        th.join()
    assert not errors, errors[:5]
    col.flush()
    n = len(x)
    assert col.num_entities == n and len(col._st.pk_to_row) == n
    # end state == sequential semantics: the row a pk maps to holds that pk's embedding (blocks may interleave between writers)
    order = [int(pk[1:]) for pk in col._st.columns["id"]]
    assert sorted(order) == list(range(n))
    want_ids, want_sc = O.cosine_topk(q, O.normalize_rows(x[order], "f32"), 5)
    got = col.search(q, "embedding", {"metric_type": "COSINE"}, 5)
    assert [[h.id for h in hits] for hits in got] == [[f"r{order[i]}" for i in row] for row in want_ids]
    assert [[h.distance for h in hits] for hits in got] == [[float(s) for s in row] for row in want_sc]
    assert sorted(r["id"] for r in col.query(expr='tag == "t2"')) == sorted(f"r{i}" for i in range(n) if i % 3 == 2)
