"""Fixture for BASELINE.json config 0 ("ICICI chunks queried with qa_subset.json, exact top-5 cosine via retrieve.py"):
   tests/golden/qa_subset_top5.json   the 40 questions of the reference's qa_subset.json (id, category, question,
                                      expected_relevant_chunks - qa_subset.json:10-371) and, for each, the oracle's
                                      top-5 (chunk ids + fp32 score bit patterns) over the 16 chunks.
The reference embeds with SentenceTransformer('all-MiniLM-L6-v2') (retrieve.py:14,27), whose weights are not available
offline, so the embeddings come from the deterministic stand-in `HashingEncoder(384)`: what this fixture pins is the
ENGINE (same top-5 and score bits as the oracle for the reference's own questions, texts and call shape), not the
encoder's semantics.  The chunk-level recall of the stand-in against `expected_relevant_chunks` is recorded as
information.  Run: python scripts/make_qa_golden.py   (needs /root/reference; the fixture travels, the reference does not)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, HERE)
from oracle import ragfin_oracle as O                      # noqa: E402
from ragfin_b200 import chunker                            # noqa: E402
from ragfin_b200.vector_rag import HashingEncoder          # noqa: E402

REF = "/root/reference"
GOLDEN = os.path.join(HERE, "tests", "golden")
with open(os.path.join(REF, "qa_subset.json")) as f:
    qa = json.load(f)
with open(os.path.join(GOLDEN, "fin_statements.json")) as f:
    chunks = chunker.build_corpus_from_bundle(json.load(f))     # insertion order of "chunking_storing (1).py":335-396
enc = HashingEncoder(384)
stored = O.normalize_rows(enc.encode([c["text"] for c in chunks]), "f32")
ids = [c["id"] for c in chunks]
out, hit, total = [], 0, 0
for item in qa["questions"]:
    wi, ws = O.cosine_topk(enc.encode([item["question"]]), stored, 5)
    top = [ids[j] for j in wi[0]]
    exp = item["expected_relevant_chunks"]
    hit += sum(1 for e in exp if e in top)
    total += len(exp)
    out.append({"id": item["id"], "category": item["category"], "question": item["question"],
                "expected_relevant_chunks": exp, "top5_ids": top,
                "top5_score_bits": ws[0].view(np.uint32).tolist()})
doc = {"source": "reference qa_subset.json (questions, expected chunks); oracle top-5 over HashingEncoder(384) embeddings of the 16 chunks",
       "encoder": "HashingEncoder(384) stand-in (MiniLM-L6-v2 weights unavailable offline)", "k": 5,
       "standin_chunk_recall_at_5": hit / total, "expected_chunks_total": total, "questions": out}
with open(os.path.join(GOLDEN, "qa_subset_top5.json"), "w") as f:
    json.dump(doc, f, ensure_ascii=False, indent=1)
print(len(out), "questions; stand-in chunk recall@5 =", round(hit / total, 3), f"({hit}/{total})")
