"""Writes tests/golden/thirdparty_topk.json: expected top-k id lists and similarities computed by THIRD-PARTY
implementations of the reference's search semantics - scikit-learn's brute-force cosine k-NN on float64 copies of the raw
(un-normalised) rows, cross-checked at generation time against scipy's cosine distance.  The oracle does not take part:
these fixtures pin the oracle (tests/test_oracle_cpu.py) and the engine (tests/test_parity_gpu.py), they do not echo it.

The reference's own engine (Milvus behind pymilvus 2.3.0: reference retrieve.py:28-34, "chunking_storing (1).py":29) cannot
run here, so this is the closest available anchor: COSINE = <q, x> / (|q| |x|), larger is better, k best in descending
order.  Cases are kept only if consecutive similarities differ by more than 1e-5 (no tie can reorder under fp32 rounding),
which is recorded per case as `min_gap`.  Inputs are regenerated in the tests from (seed, n, dim, scale_seed) with the
synthetic generator, whose values are exact in fp32, so no embedding needs to be stored.

    python scripts/make_thirdparty_golden.py
"""
import json
import os
import sys

import numpy as np
import scipy
import sklearn
from scipy.spatial.distance import cdist
from sklearn.neighbors import NearestNeighbors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _thirdparty import raw_inputs  # noqa: E402


if __name__ == "__main__":
    cases = []
    for name, seed, n, dim, nq, k, scale_seed in [
        ("ref_shape_16x384_top3", 900, 16, 384, 6, 3, 1),          # the reference's own shape (16 chunks, MiniLM width, k = 3)
        ("ref_shape_16x384_top5", 902, 16, 384, 8, 5, 2),          # BASELINE config 0: top-5
        ("2000x768_top10", 904, 2000, 768, 6, 10, 3),
        ("30000x768_top10", 906, 30000, 768, 8, 10, 4),
        ("5000x384_top100", 908, 5000, 384, 3, 100, 5),
        ("12000x1024_top10", 910, 12000, 1024, 4, 10, 6),
        ("3000x100_top20", 912, 3000, 100, 5, 20, 7),
        ("hybrid_limit_1000_of_4000x128", 914, 4000, 128, 2, 1000, 8),   # graph_cons.py:279 asks for limit = 1000
    ]:
        x, q = raw_inputs(seed, n, dim, nq, scale_seed)
        x64, q64 = x.astype(np.float64), q.astype(np.float64)
        kk = min(k + 1, n)
        nn = NearestNeighbors(n_neighbors=kk, metric="cosine", algorithm="brute").fit(x64)
        dist, ind = nn.kneighbors(q64)
        d = cdist(q64, x64, metric="cosine")
        order = np.argsort(d, axis=1, kind="stable")[:, :kk]
        assert np.array_equal(order, ind), name                      # two third-party implementations agree
        sims = 1.0 - dist
        gaps = -np.diff(sims, axis=1)                                # gaps[:, k-1] = k-th vs (k+1)-th: the boundary
        # strict cases: every consecutive gap (boundary included) > 1e-5 -> the id LIST must be reproduced;
        # large-k cases: gaps of 1e-7 are natural, so only the boundary must be clear (> 1e-6) and the id SET is compared,
        # plus the order at every position whose neighbours are > 1e-6 apart
        strict = k <= 20
        lim = 1e-5 if strict else 1e-6
        keep = [i for i in range(nq) if (gaps[i].min() > lim if strict else (kk == k or gaps[i, k - 1] > lim))]
        assert len(keep) >= max(1, nq // 2), (name, gaps.min(axis=1))
        ind, sims = ind[:, :k], sims[:, :k]
        cases.append(dict(name=name, seed=seed, n=n, dim=dim, nq=nq, k=k, scale_seed=scale_seed, queries=keep, strict=strict,
                          ids=ind[keep].tolist(), sims=[[float(v) for v in r] for r in sims[keep]],
                          min_gap=float(gaps[keep].min()), boundary_gap=float(gaps[keep][:, k - 1].min()) if kk > k else None))
        print(name, "kept", len(keep), "of", nq)
    out = os.path.join(ROOT, "tests", "golden", "thirdparty_topk.json")
    with open(out, "w") as f:
        json.dump({"generator": "scripts/make_thirdparty_golden.py",
                   "by": f"scikit-learn {sklearn.__version__} NearestNeighbors(metric='cosine', algorithm='brute') on float64; "
                         f"cross-checked with scipy {scipy.__version__} cdist(metric='cosine')",
                   "semantics": "cosine similarity of raw rows, descending, k best", "cases": cases}, f)
    print("wrote", out, os.path.getsize(out), "bytes")
