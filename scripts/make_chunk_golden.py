"""Fixtures for the chunker, taken from the reference in THIS container (it does not travel):
   tests/golden/fin_statements.json    the reference's input statements (extract_data/icici_q*_2023/*.json) bundled as
                                       {quarter directory: {file name: document}}
   tests/golden/reference_chunks.json  the reference's own output (FinRag_knowledge_graph/chunks.json): the golden texts
Run: python scripts/make_chunk_golden.py   (needs /root/reference)."""
import json
import os

REF = "/root/reference"
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(HERE, "tests", "golden")
src_data = os.path.join(REF, "extract_data")
bundle = {}
for q in sorted(os.listdir(src_data)):
    bundle[q] = {}
    for f in sorted(os.listdir(os.path.join(src_data, q))):
        if f.endswith(".json"):
            with open(os.path.join(src_data, q, f)) as fh:
                bundle[q][f] = json.load(fh)
with open(os.path.join(dst, "fin_statements.json"), "w") as f:
    json.dump(bundle, f, ensure_ascii=False, separators=(",", ":"))
with open(os.path.join(REF, "FinRag_knowledge_graph", "chunks.json")) as f:
    chunks = json.load(f)
with open(os.path.join(dst, "reference_chunks.json"), "w") as f:
    json.dump(chunks, f, ensure_ascii=False, indent=1)
print(len(chunks), "reference chunks;", sum(len(v) for v in bundle.values()), "input documents")
