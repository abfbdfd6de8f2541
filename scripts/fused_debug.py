"""Appended / rescored row counts of the one-kernel search on a few corpora (diagnostics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ragfin_b200
from oracle import c_oracle as C, ragfin_oracle as O

def run(name, x, q, k, dtype):
    idx = ragfin_b200.Index(x.shape[1], dtype, capacity=len(x)); idx.add(x); idx.set_fused(True, 1)
    got = idx.search(q, k); st = idx.stats(); a, r = idx.fused_counts(len(q))
    wi, ws = C.cosine_topk(q, C.normalize_rows(x, dtype), k)
    print(f"{name} {dtype} n={len(x)} nq={len(q)} k={k}: path {st['path']} rescanned {st['queries_rescanned']} appended {a.tolist()} rescored {r.tolist()} "
          f"parity {np.array_equal(got[0], wi) and np.array_equal(got[1].view(np.uint32), ws.view(np.uint32))}", flush=True)
    idx.close()

n, dim, k = 200000, 64, 10
x = O.synth_rows(320, 0, n, dim); q = O.synth_rows(321, 0, 2, dim)
s = C.exact_scores(C.normalize_rows(x, "f32"), C.normalize_rows(q[:1], "f32")[0])
for dtype in ("bf16", "f32"):
    run("random", x, q, k, dtype)
    run("ascending", np.ascontiguousarray(x[np.argsort(s, kind="stable")]), q, k, dtype)
    run("descending", np.ascontiguousarray(x[np.argsort(-s, kind="stable")]), q, k, dtype)
x = O.synth_rows(300, 0, 70000, 128, dup_every=211, zero_every=4099)
for nq, k in ((1, 10), (16, 100), (64, 10), (5, 128)):
    run("random", x, O.synth_rows(301, 0, nq, 128), k, "bf16")
for rows in (1_250_000, 10_000_000):
    idx = ragfin_b200.Index(768, "bf16", capacity=rows)
    for r in range(0, rows, 1_000_000):
        idx.add_synthetic(1234, r, min(1_000_000, rows - r))
    for nq, k in ((1, 10), (16, 10), (64, 10), (1, 100), (16, 100), (4, 128)):
        q = O.synth_rows(1235, 0, nq, 768)
        idx.search(q, k); st = idx.stats(); a, r = idx.fused_counts(nq)
        print(f"synthetic bf16 rows={rows} nq={nq} k={k}: path {st['path']} rescanned {st['queries_rescanned']} appended min/mean/max "
              f"{a.min()}/{a.mean():.0f}/{a.max()} rescored max {r.max()}", flush=True)
    idx.close()
