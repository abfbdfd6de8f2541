"""Writes tests/golden/*.json from the numpy oracle (run in the authoring container).

The reference cannot produce score/ranking vectors here (its arithmetic lives in an external
Milvus server; see oracle/ragfin_oracle.py), so these fixtures pin the ORACLE, not the reference:
"parity unpinned".  fin_chunks_collection.json takes the 16 chunk ids / periods / types from the
reference's FinRag_knowledge_graph/chunks.json (metadata only, no text) when /root/reference is
mounted, with deterministic stand-in embeddings because the MiniLM encoder is not available.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ragfin_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

cases = []
for name, seed, n, dim, nq, k, dtype, dup, zero in [
    ("f32_384_top3", 100, 16, 384, 5, 3, "f32", 0, 0),          # the reference's real shape (16 x 384, k=3)
    ("f32_384_top5_dups", 110, 512, 384, 4, 5, "f32", 7, 0),
    ("bf16_768_top10", 120, 2000, 768, 4, 10, "bf16", 0, 0),
    ("f16_768_top100", 130, 1500, 768, 2, 100, "f16", 11, 29),
    ("f32_1024_top10", 140, 1000, 1024, 3, 10, "f32", 0, 13),
    ("bf16_100_top20_k_gt_n", 150, 12, 100, 2, 20, "bf16", 0, 0),
    ("f16_33_top1", 160, 300, 33, 3, 1, "f16", 3, 0),
]:
    x = O.synth_rows(seed, 0, n, dim, dup, zero)
    q = O.synth_rows(seed + 1, 0, nq, dim)
    ids, sc = O.cosine_topk(q, O.normalize_rows(x, dtype), k)
    cases.append(dict(name=name, seed=seed, n=n, dim=dim, nq=nq, k=k, dtype=dtype, dup_every=dup, zero_every=zero,
                      ids=ids.tolist(), score_bits=sc.view(np.uint32).tolist()))
with open(os.path.join(OUT, "topk_cases.json"), "w") as f:
    json.dump({"generator": "scripts/make_golden.py", "oracle": "oracle/ragfin_oracle.py", "cases": cases}, f)

ref_chunks = "/root/reference/FinRag_knowledge_graph/chunks.json"
# insertion order of "chunking_storing (1).py":341-396: per quarter profitability, balance sheet, ratios, segment
order_types = ["profitability_analysis", "balance_sheet_health", "key_ratios", "segment_performance"]
chunks = []
if os.path.exists(ref_chunks):
    by_id = {c["id"]: c for c in json.load(open(ref_chunks))}
    for qn in (1, 2, 3, 4):
        for t in order_types:
            c = by_id[f"icici_q{qn}_fy2024_{t}"]
            chunks.append({"id": c["id"], "period": c["period"], "chunk_type": c["type"]})
else:
    for qn in (1, 2, 3, 4):
        for t in order_types:
            chunks.append({"id": f"icici_q{qn}_fy2024_{t}", "period": f"Q{qn}_FY2024", "chunk_type": t})
seed = 2023
st = O.normalize_rows(O.synth_rows(seed, 0, 16, 384), "f32")
q = O.synth_rows(seed + 1, 0, 5, 384)
ids, sc = O.cosine_topk(q, st, 3)
queries = [{"top3_ids": [chunks[i]["id"] for i in ids[qi]], "top3_score_bits": sc[qi].view(np.uint32).tolist()}
           for qi in range(5)]
with open(os.path.join(OUT, "fin_chunks_collection.json"), "w") as f:
    json.dump({"generator": "scripts/make_golden.py", "dim": 384, "seed": seed, "chunks": chunks, "queries": queries}, f, indent=1)
print("wrote", os.listdir(OUT))
