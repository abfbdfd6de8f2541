"""Callers of the hot path, mirrored from the reference so that its own usage reads unchanged.

  VectorRAG.search            vector_rag_mcp/main.py:48-70     (list of ranked context dicts)
  search_vectors tool dict    vector_rag_mcp/main.py:134-146   ({"status","query","results","result_count"})
  get_collection_stats        vector_rag_mcp/main.py:157-169
  SimpleRAG contexts          retrieve.py:26-47
  hybrid vector hits + merge  graph_cons.py:272-293, 326-340
  SearchRequest bounds        adapters/vectorrag_adapter.py:24-26 (query >= 5 chars, 1 <= top_k <= 20)

The LLM answer step (Gemini) and the MCP / REST servers are out of scope.  The
encoder is injected: anything with `.encode(list[str]) -> float32 [n, dim]` (the reference uses
SentenceTransformer('all-MiniLM-L6-v2'), not available offline).  `HashingEncoder` is a
deterministic stand-in for tests and demos.
"""
from __future__ import annotations

import hashlib
from typing import Any, Dict, List, Optional

import numpy as np

from .milvus_compat import Collection, connections

SEARCH_OUTPUT_FIELDS = ["text", "period", "chunk_type", "statement_type", "primary_value"]


class HashingEncoder:
    """Deterministic bag-of-tokens embedding (signed feature hashing). Stand-in for MiniLM-L6 (384-d)."""

    def __init__(self, dim: int = 384):
        self.dim = dim

    def encode(self, texts, **kwargs) -> np.ndarray:
        if isinstance(texts, str):
            texts = [texts]
        out = np.zeros((len(texts), self.dim), dtype=np.float32)
        for i, t in enumerate(texts):
            for tok in str(t).lower().split():
                h = int.from_bytes(hashlib.blake2b(tok.encode(), digest_size=8).digest(), "little")
                out[i, h % self.dim] += 1.0 if (h >> 63) else -1.0
            n = float(np.linalg.norm(out[i]))
            if n > 0:
                out[i] /= n
        return out


def validate_search_request(query: str, top_k: int = 3) -> int:
    """adapters/vectorrag_adapter.py:24-26: query min_length 5, top_k in [1, 20], default 3."""
    if not isinstance(query, str) or len(query) < 5:
        raise ValueError("query must be a string of at least 5 characters")
    if not isinstance(top_k, int) or not 1 <= top_k <= 20:
        raise ValueError("top_k must be an integer in [1, 20]")
    return top_k


class VectorRAG:
    """vector_rag_mcp/main.py:36-70 without the Gemini client."""

    def __init__(self, encoder, collection_name: str = "fin_chunks", host: str = "localhost", port: str = "19530",
                 collection: Optional[Collection] = None):
        self.similarity_model = encoder
        if collection is None:
            connections.connect("default", host=host, port=port)
            collection = Collection(collection_name)
            collection.load()
        self.collection = collection
        self.collection_name = collection_name

    def search(self, query: str, top_k: int = 3) -> List[Dict[str, Any]]:
        query_embedding = self.similarity_model.encode([query])
        results = self.collection.search(query_embedding, "embedding", {"metric_type": "COSINE"}, top_k,
                                         output_fields=SEARCH_OUTPUT_FIELDS)
        contexts = []
        for i, result in enumerate(results[0]):
            contexts.append({
                "rank": i + 1,
                "text": result.entity.text,
                "period": result.entity.period,
                "chunk_type": result.entity.chunk_type,
                "statement_type": result.entity.statement_type,
                "primary_value": result.entity.primary_value,
                "score": float(result.score),
            })
        return contexts

    def search_vectors(self, query: str, top_k: int = 3) -> Dict[str, Any]:
        """The MCP tool's envelope: errors never escape (vector_rag_mcp/main.py:137-146)."""
        try:
            contexts = self.search(query, top_k)
            return {"status": "success", "query": query, "results": contexts, "result_count": len(contexts)}
        except Exception as e:  # noqa: BLE001 - the reference swallows everything into the envelope
            return {"status": "error", "message": str(e), "query": query}

    def get_collection_stats(self) -> Dict[str, Any]:
        try:
            return {"status": "success", "collection_name": self.collection_name,
                    "total_chunks": self.collection.num_entities}
        except Exception as e:  # noqa: BLE001
            return {"status": "error", "message": str(e)}

    def retrieve_contexts(self, question: str, top_k: int = 3):
        """retrieve.py:26-47: (text, period, chunk_type, score) tuples in rank order."""
        emb = self.similarity_model.encode([question])
        results = self.collection.search(emb, "embedding", {"metric_type": "COSINE"}, top_k,
                                         output_fields=["text", "period", "chunk_type"])
        return [(r.entity.text, r.entity.period, r.entity.chunk_type, r.score) for r in results[0]]

    def hybrid_vector_chunks(self, question: str, limit: int = 1000) -> List[Dict[str, Any]]:
        """graph_cons.py:272-293: the vector half of hybrid_query_simple (limit=1000)."""
        emb = self.similarity_model.encode([question])
        res = self.collection.search(emb, "embedding", {"metric_type": "COSINE"}, limit=limit,
                                     output_fields=["id", "text", "period", "chunk_type"])
        chunks = []
        if res and len(res) > 0:
            for hit in res[0]:
                chunks.append({"id": hit.entity.get("id"), "text": hit.entity.get("text"),
                               "period": hit.entity.get("period"), "chunk_type": hit.entity.get("chunk_type"),
                               "score": hit.score})
        return chunks


def merge_hybrid(vector_chunks: List[dict], graph_chunks: List[dict]) -> List[dict]:
    """graph_cons.py:326-340: vector hits first, then graph hits, de-duplicated by id."""
    seen, out = set(), []
    for chunk in list(vector_chunks) + list(graph_chunks):
        if chunk["id"] not in seen:
            out.append(chunk)
            seen.add(chunk["id"])
    return out
