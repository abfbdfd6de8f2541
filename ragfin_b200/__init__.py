"""ragfin_b200: B200-native exact cosine top-k for rag-fin's vector-RAG hot path.

The compute path is libragfin.so (hand-written sm_100a CUDA behind the C ABI in
include/ragfin.h); this package is the reference-shaped host layer above it.
"""
from ._lib import RagfinError, SO_PATH, SYMBOLS  # noqa: F401
from .engine import Index, merge_topk, PackedHits, MAX_TOPK  # noqa: F401

__all__ = ["Index", "merge_topk", "RagfinError", "MAX_TOPK"]
