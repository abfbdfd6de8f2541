"""Fixtures for the chunker, taken from the reference in THIS container (it does not travel):
   tests/golden/extract_data/icici_q*_2023/*.json   the reference's input statements (extract_data/)
   tests/golden/reference_chunks.json               the reference's own output (FinRag_knowledge_graph/chunks.json)
Run: python scripts/make_chunk_golden.py   (needs /root/reference)."""
import json
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(HERE, "tests", "golden")
src_data = os.path.join(REF, "extract_data")
for q in sorted(os.listdir(src_data)):
    os.makedirs(os.path.join(dst, "extract_data", q), exist_ok=True)
    for f in sorted(os.listdir(os.path.join(src_data, q))):
        if f.endswith(".json"):
            shutil.copyfile(os.path.join(src_data, q, f), os.path.join(dst, "extract_data", q, f))
with open(os.path.join(REF, "FinRag_knowledge_graph", "chunks.json")) as f:
    chunks = json.load(f)
with open(os.path.join(dst, "reference_chunks.json"), "w") as f:
    json.dump(chunks, f, ensure_ascii=False, indent=1)
print(len(chunks), "reference chunks;", sum(len(os.listdir(os.path.join(dst, "extract_data", q))) for q in os.listdir(os.path.join(dst, "extract_data"))), "input files")
