"""BASELINE.json config 0 on CPU: the reference's 16 chunks queried with the 40 questions of its qa_subset.json, exact
top-5 cosine, one question at a time as retrieve.py:26-47 does.  Here the shim runs on the oracle-backed stand-in
(host logic + fixture consistency); tests/test_config0_gpu.py runs the same calls on the engine."""
import json
import os

import numpy as np

from ragfin_b200 import chunker, milvus_compat as mc
from ragfin_b200.vector_rag import HashingEncoder, VectorRAG
from test_shim_cpu import OracleIndex, reference_fields

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_qa():
    with open(os.path.join(GOLDEN, "qa_subset_top5.json")) as f:
        return json.load(f)


def build_fin_chunks(index_factory=None):
    """"chunking_storing (1).py":11-29,335-396: statements -> 16 chunks -> encode -> insert / flush / load."""
    if mc.utility.has_collection("fin_chunks"):
        mc.utility.drop_collection("fin_chunks")
    kw = {"index_factory": index_factory} if index_factory else {}
    col = mc.Collection("fin_chunks", mc.CollectionSchema(reference_fields(), "Financial complete context chunks"), **kw)
    col.create_index("embedding", {"index_type": "IVF_FLAT", "metric_type": "COSINE", "params": {"nlist": 128}})
    with open(os.path.join(GOLDEN, "fin_statements.json")) as f:
        chunks = chunker.build_corpus_from_bundle(json.load(f))
    enc = HashingEncoder(384)
    chunker.ingest_chunks(col, chunks, enc.encode)
    return col, chunks, enc


def check_all_questions(col, chunks, enc, qa):
    """Every question through the reference's three call shapes; ids and fp32 score bits equal to the golden top-5."""
    rag = VectorRAG(enc, collection=col)
    by_text = {c["text"]: c["id"] for c in chunks}
    hit = total = 0
    for item in qa["questions"]:
        contexts = rag.search(item["question"], 5)                                   # vector_rag_mcp/main.py:48-70
        assert [c["rank"] for c in contexts] == [1, 2, 3, 4, 5]
        assert [by_text[c["text"]] for c in contexts] == item["top5_ids"], item["id"]
        assert [np.float32(c["score"]).view(np.uint32).item() for c in contexts] == item["top5_score_bits"], item["id"]
        tuples = rag.retrieve_contexts(item["question"], 5)                          # retrieve.py:26-47
        assert [by_text[t[0]] for t in tuples] == item["top5_ids"]
        assert all(t[1] in ("Q1_FY2024", "Q2_FY2024", "Q3_FY2024", "Q4_FY2024") for t in tuples)
        hit += sum(1 for e in item["expected_relevant_chunks"] if e in item["top5_ids"])
        total += len(item["expected_relevant_chunks"])
    # all 40 at once (one search call, nq = 40): same lists
    emb = enc.encode([item["question"] for item in qa["questions"]])
    res = col.search(emb, "embedding", {"metric_type": "COSINE"}, 5, output_fields=["period", "chunk_type"])
    assert [[h.id for h in hits] for hits in res] == [item["top5_ids"] for item in qa["questions"]]
    assert [[np.float32(h.score).view(np.uint32).item() for h in hits] for hits in res] == \
           [item["top5_score_bits"] for item in qa["questions"]]
    assert total == qa["expected_chunks_total"] and abs(hit / total - qa["standin_chunk_recall_at_5"]) < 1e-12


def test_qa_subset_top5_on_the_oracle_backed_shim():
    qa = load_qa()
    assert len(qa["questions"]) == 40 and qa["k"] == 5
    col, chunks, enc = build_fin_chunks(OracleIndex)
    assert col.num_entities == 16
    ids = {c["id"] for c in chunks}
    assert all(set(item["expected_relevant_chunks"]) <= ids for item in qa["questions"])   # the reference's ids are ours
    check_all_questions(col, chunks, enc, qa)
    mc.utility.drop_collection("fin_chunks")
