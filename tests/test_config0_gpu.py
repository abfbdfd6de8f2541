"""BASELINE.json config 0 on the engine: the reference's 16 chunks, the 40 questions of its qa_subset.json, exact top-5
cosine through the pymilvus-shaped shim and the VectorRAG mirror; chunk ids and fp32 score bits equal to the committed
oracle top-5 (tests/golden/qa_subset_top5.json, written by scripts/make_qa_golden.py)."""
import pytest

from ragfin_b200 import milvus_compat as mc
from test_config0_cpu import build_fin_chunks, check_all_questions, load_qa

pytestmark = pytest.mark.gpu


def test_qa_subset_top5_on_the_engine():
    qa = load_qa()
    col, chunks, enc = build_fin_chunks()
    assert col.num_entities == 16
    check_all_questions(col, chunks, enc, qa)
    mc.utility.drop_collection("fin_chunks")
