"""Appended-row counts of the one-kernel search on a templated corpus whose topics lie inside one CTA's slice (diagnostics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ragfin_b200
from oracle import c_oracle as C
from ragfin_b200.synthetic import synth_topic_rows, TOPIC_SEED_OFFSET
rows_fn = lambda s, r0, m, d: C.synth_rows(s, r0, m, d)
for dim, n, topic_rows in ((128, 1_212_416, 4096), (768, 2_424_832, 16384)):
    idx = ragfin_b200.Index(dim, "bf16", capacity=n)
    for r in range(0, n, 500_000):
        idx.add_synthetic_topics(900, r, min(500_000, n - r), topic_rows, 3)
    for t in (3, 17, 40):
        q = (C.synth_rows(900 + TOPIC_SEED_OFFSET, t, 1, dim) + C.synth_rows(5142, t, 1, dim) * np.float32(0.25)).astype(np.float32)
        ids, sc = idx.search(q, 10)
        st = idx.stats(); a, r = idx.fused_counts(1)
        thr, app = idx.fused_ctas()
        top = np.argsort(-app)[:4]
        print("   busiest CTAs:", [(int(c), int(app[c]), float(thr[c])) for c in top], " median thr", float(np.median(thr[:148])), "max thr", float(thr[:148].max()))
        blk = C.normalize_rows(synth_topic_rows(900, t * topic_rows, topic_rows, dim, topic_rows, 3, rows_fn), "bf16")
        s = C.exact_scores(blk, C.normalize_rows(q, "f32")[0])
        ss = np.sort(s)[::-1]
        print(f"dim={dim} n={n} topic={t}: path {st['path']} rescanned {st['queries_rescanned']} appended {a.tolist()} rescored {r.tolist()} "
              f"topic scores mean {s.mean():.5f} std {s.std():.5f} top1 {ss[0]:.5f} top10 {ss[9]:.5f} within 8.6e-4 of 10th: {(s >= ss[9] - 8.6e-4).sum()} "
              f"ids in topic: {bool(((ids[0] // topic_rows) == t).all())}", flush=True)
    idx.close()
