"""Timed CPU baseline: a port of the reference's CPU retrieval path.  Test/bench infrastructure only.

The reference searches through Milvus, whose brute-force COSINE path (knowhere -> faiss) is a
per-row SIMD inner product for small query batches and a corpus-blocked sgemm + heap for
larger ones (faiss switches at nq >= 20).  This module restates exactly that
with torch's CPU BLAS (MKL), fp32 corpus, all host threads.  It returns fp32-accumulated
scores (what such a CPU engine returns); tests check it against the canonical oracle within
1e-5 relative.  kind = "port" in bench.py's cpu_baseline.
"""
from __future__ import annotations

import numpy as np
import torch

BLAS_THRESHOLD = 20       # faiss distance_compute_blas_threshold
ROW_BLOCK = 65536         # corpus rows per sgemm block


def fast_topk(stored32: torch.Tensor, queries32: np.ndarray, k: int):
    """stored32: torch fp32 [N, D] normalised rows (CPU). queries32: raw fp32 [nq, D]."""
    q = torch.from_numpy(np.ascontiguousarray(queries32, dtype=np.float32))
    q = q / q.norm(dim=1, keepdim=True).clamp_min(1e-30)
    n = stored32.shape[0]
    kk = min(k, n)
    if q.shape[0] < BLAS_THRESHOLD:
        out_s, out_i = [], []
        for i in range(q.shape[0]):
            s = torch.mv(stored32, q[i])
            v, idx = torch.topk(s, kk)
            out_s.append(v)
            out_i.append(idx)
        return torch.stack(out_i).numpy(), torch.stack(out_s).numpy()
    best_s = torch.full((q.shape[0], kk), -float("inf"))
    best_i = torch.full((q.shape[0], kk), -1, dtype=torch.int64)
    for b0 in range(0, n, ROW_BLOCK):
        blk = stored32[b0:b0 + ROW_BLOCK]
        s = q @ blk.T
        v, idx = torch.topk(s, min(kk, blk.shape[0]), dim=1)
        cs = torch.cat([best_s, v], dim=1)
        ci = torch.cat([best_i, idx + b0], dim=1)
        best_s, sel = torch.topk(cs, kk, dim=1)
        best_i = torch.gather(ci, 1, sel)
    return best_i.numpy(), best_s.numpy()
