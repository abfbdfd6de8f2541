"""The drop-in claim, checked with the reference's OWN code: its scripts are executed unmodified from /root/reference with
`pymilvus` bound to `ragfin_b200.milvus_compat` (the one-line import switch of INTEGRATION.md, done here through
sys.modules), and every other external service stubbed (MiniLM -> the hashing stand-in encoder, Gemini, FastMCP, dotenv,
the Neo4j driver).  What runs is the reference's ingest script ("chunking_storing (1).py": schema, create_index, chunk
building from extract_data/, insert / flush / load, search_financial_query), retrieve.py's SimpleRAG, the MCP server module
(VectorRAG.search, the search_vectors / get_collection_stats / health_check tools), test_vector.py's collection check and
graph_cons.py's hybrid query (search limit=1000 + `id in [...]` query).  Results are compared with the oracle's exact
cosine top-k over the same stand-in embeddings.

Runs in the build container only (/root/reference is not on the GPU box) and on CPU: the collection's engine is replaced by
an oracle-backed index - this suite checks the SURFACE the reference needs; the engine behind the same surface is checked by
the GPU suite (tests/test_config0_gpu.py, tests/test_parity_gpu.py)."""
import contextlib
import io
import os
import runpy
import sys
import time
import types

import pytest

from oracle import ragfin_oracle as O
from ragfin_b200 import milvus_compat as mc
from ragfin_b200.vector_rag import HashingEncoder
from test_shim_cpu import OracleIndex

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "chunking_storing (1).py")), reason="reference tree not mounted")

class _Answer:
    text = " stubbed answer "


class _GenerativeModel:
    prompts = []

    def __init__(self, name):
        self.name = name

    def generate_content(self, prompt):
        _GenerativeModel.prompts.append(prompt)
        return _Answer()


class _Session:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, *a, **kw):
        return types.SimpleNamespace(data=lambda: [], single=lambda: None)


def _stub_modules(monkeypatch):
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = lambda name, **kw: HashingEncoder(384)
    genai = types.ModuleType("google.generativeai")
    genai.configure = lambda **kw: None
    genai.GenerativeModel = _GenerativeModel
    google = types.ModuleType("google")
    google.generativeai = genai
    fastmcp = types.ModuleType("fastmcp")

    class FastMCP:
        def __init__(self, name):
            self.name, self.tools = name, {}

        def tool(self, *a, **kw):
            def deco(fn):
                self.tools[fn.__name__] = fn
                return fn
            return deco

        def run(self, *a, **kw):
            raise AssertionError("the server loop must not start in a test")
    fastmcp.FastMCP = FastMCP
    dotenv = types.ModuleType("dotenv")
    dotenv.load_dotenv = lambda *a, **kw: False
    neo4j = types.ModuleType("neo4j")
    neo4j.GraphDatabase = types.SimpleNamespace(driver=lambda uri, auth=None: types.SimpleNamespace(session=lambda: _Session(), close=lambda: None))
    for name, mod in (("pymilvus", mc), ("sentence_transformers", st), ("google", google), ("google.generativeai", genai),
                      ("fastmcp", fastmcp), ("dotenv", dotenv), ("neo4j", neo4j)):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.setattr(mc, "_default_index_factory", lambda dim, dtype, cap, dev: OracleIndex(dim, dtype, cap, dev))
    monkeypatch.setattr(time, "sleep", lambda s: None)
    monkeypatch.setattr(sys, "dont_write_bytecode", True)                # nothing is written next to the reference's sources


@pytest.fixture
def ingested(monkeypatch):
    """The reference's ingest script, run as it is, from the reference's own working directory."""
    _stub_modules(monkeypatch)
    monkeypatch.chdir(REF)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        ns = runpy.run_path(os.path.join(REF, "chunking_storing (1).py"), run_name="reference_ingest")
    yield ns, out.getvalue()
    mc.utility.drop_collection("fin_chunks")


def _oracle_topk(ns, query, k):
    enc = HashingEncoder(384)
    chunks = ns["all_chunks"]
    x = O.normalize_rows(enc.encode([c["text"] for c in chunks]), "f32")
    ids, sc = O.cosine_topk(enc.encode([query]), x, k)
    return [int(i) for i in ids[0] if i >= 0], [float(s) for s in sc[0][: min(k, len(chunks))]]


def test_reference_ingest_script_runs_unmodified_against_the_shim(ingested):
    ns, printed = ingested
    chunks = ns["all_chunks"]
    assert len(chunks) == 16 and "Inserted 16 chunks into Milvus" in printed
    col = ns["collection"]
    assert isinstance(col, mc.Collection) and col.num_entities == 16
    # the three searches at the end of the script printed their hits: score, chunk type, period
    assert printed.count("Query: '") == 3 and printed.count("1. Score: ") == 3
    q = "What was ICICI's Q1 net profit and profitability?"
    rows, scores = _oracle_topk(ns, q, 3)
    want_first = f"1. Score: {scores[0]:.3f} | Type: {chunks[rows[0]]['chunk_type']} | Period: {chunks[rows[0]]['period']}"
    assert want_first in printed
    # search_financial_query is the script's own function: call it once more, for top-5 (BASELINE config 0)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        ns["search_financial_query"]("How did retail banking perform in Q2?", top_k=5)
    rows, scores = _oracle_topk(ns, "How did retail banking perform in Q2?", 5)
    lines = [ln for ln in out.getvalue().splitlines() if ". Score: " in ln]
    assert lines == [f"{i + 1}. Score: {s:.3f} | Type: {chunks[r]['chunk_type']} | Period: {chunks[r]['period']}"
                     for i, (r, s) in enumerate(zip(rows, scores))]


def test_reference_retrieve_and_mcp_server_modules_run_unmodified(ingested, monkeypatch):
    ns, _ = ingested
    chunks = ns["all_chunks"]
    monkeypatch.syspath_prepend(REF)
    monkeypatch.delitem(sys.modules, "retrieve", raising=False)
    import retrieve                                                   # /root/reference/retrieve.py
    rag = retrieve.SimpleRAG("no-key")
    _GenerativeModel.prompts.clear()
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        answer = rag.search_and_answer("What was the EPS for Q4 FY2024?", top_k=5)
    assert answer == "stubbed answer"
    rows, scores = _oracle_topk(ns, "What was the EPS for Q4 FY2024?", 5)
    heads = [ln.strip() for ln in out.getvalue().splitlines() if "(Score: " in ln]
    assert heads == [f"{i + 1}. [{chunks[r]['period']} - {chunks[r]['chunk_type']}] (Score: {s:.3f})" for i, (r, s) in enumerate(zip(rows, scores))]
    assert all(f"Context {i + 1}: {chunks[r]['text']}" in _GenerativeModel.prompts[-1] for i, r in enumerate(rows))

    # the MCP server module: importing it builds VectorRAG over the loaded collection and registers the tools
    with contextlib.redirect_stdout(io.StringIO()):
        srv = runpy.run_path(os.path.join(REF, "vector_rag_mcp", "main.py"), run_name="reference_mcp_server")
    hits = srv["rag"].search("net profit Q1", top_k=3)                # test_vector.py:97-100
    rows, scores = _oracle_topk(ns, "net profit Q1", 3)
    assert [h["rank"] for h in hits] == [1, 2, 3]
    assert [h["text"] for h in hits] == [chunks[r]["text"] for r in rows]
    assert [h["score"] for h in hits] == scores and all(isinstance(h["score"], float) for h in hits)
    assert all(h[f] == chunks[r][f] for h, r in zip(hits, rows) for f in ("period", "chunk_type", "statement_type", "primary_value"))
    env = srv["search_vectors"]("net profit Q1", 3)
    assert env["status"] == "success" and env["result_count"] == 3 and env["results"] == hits and env["query"] == "net profit Q1"
    bad = srv["search_vectors"]("net profit Q1", 0)                   # an engine error becomes the tool's error envelope
    assert bad["status"] == "error" and "limit" in bad["message"]
    stats = srv["get_collection_stats"]()
    assert stats["status"] == "success" and stats["total_chunks"] == 16 and stats["collection_name"] == "fin_chunks"
    health = srv["health_check"]()
    assert health["status"] == "healthy" and health["total_chunks"] == 16
    answered = srv["answer_question"]("What were the total assets in Q3 FY2024?", 3)
    assert answered["status"] == "success" and answered["context_count"] == 3


def test_reference_collection_check_and_hybrid_query_run_unmodified(ingested, monkeypatch):
    ns, _ = ingested
    chunks = ns["all_chunks"]
    # test_vector.py's first block: connect, load, num_entities, query(expr="", limit=3)  (the HTTP checks behind it fail and are
    # caught by the script itself: no server is listening)
    requests_stub = types.ModuleType("requests")

    def _no_server(*a, **kw):
        raise ConnectionError("no server in a unit test")
    requests_stub.get = requests_stub.post = _no_server
    requests_stub.exceptions = types.SimpleNamespace(ConnectionError=ConnectionError, RequestException=Exception)
    monkeypatch.setitem(sys.modules, "requests", requests_stub)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        try:
            runpy.run_path(os.path.join(REF, "test_vector.py"), run_name="reference_test_vector")
        except (ConnectionError, SystemExit):
            pass                                                      # whatever the script does about its HTTP part
    printed = out.getvalue()
    assert "Milvus Connected" in printed and "Total chunks: 16" in printed and "Milvus Error" not in printed
    for i, c in enumerate(chunks[:3], 1):
        assert f"{i}. {c['id']}" in printed and f"Period: {c['period']}" in printed

    # graph_cons.py: the hybrid query's vector half (limit=1000 -> all 16 chunks, ranked) and its `id in [...]` lookup
    monkeypatch.syspath_prepend(REF)
    monkeypatch.delitem(sys.modules, "graph_cons", raising=False)
    with contextlib.redirect_stdout(io.StringIO()):
        import graph_cons
        hy = graph_cons.FinancialHybridRAG("bolt://nowhere", "u", "p", "fin_chunks")
        loaded = hy.load_chunks_from_milvus()                         # query(expr="", limit=1000, output_fields=[6 fields])
    assert [c["id"] for c in loaded] == [c["id"] for c in chunks] and loaded[5]["primary_value"] == chunks[5]["primary_value"]
    graph_ids = [chunks[11]["id"], chunks[2]["id"]]
    hy.graph_search = lambda question: [{"source_chunk": graph_ids[0]}, {"source_chunk": graph_ids[1]}, {"other": 1}]
    question = "How did retail banking perform in Q3 FY2024?"
    with contextlib.redirect_stdout(io.StringIO()):
        merged = hy.hybrid_query_simple(question)
    rows, scores = _oracle_topk(ns, question, 1000)
    assert len(rows) == 16                                            # limit above N returns every chunk
    assert [m["id"] for m in merged] == [chunks[r]["id"] for r in rows]          # graph chunks are duplicates of vector hits here
    assert [m["score"] for m in merged] == scores
    assert all(m["text"] == chunks[r]["text"] and m["period"] == chunks[r]["period"] for m, r in zip(merged, rows))
    got = hy.vector_store.query(expr=f"id in {str(graph_ids)}".replace("'", '"'),                  # graph_cons.py:304-311
                                output_fields=["id", "text", "period", "chunk_type"])
    assert sorted(c.get("id") for c in got) == sorted(graph_ids)
