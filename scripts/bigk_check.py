"""k = 1000 (the hybrid call's limit, graph_cons.py:279) on 10M x 768 bf16: batched pipeline against the one-query exact path."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragfin_b200
from ragfin_b200.synthetic import synth_rows
rows = 10_000_000
for mode in ("batched", "one-query"):
    if mode == "one-query":
        os.environ["RAGFIN_NO_BIGK_BATCHED"] = "1"
    idx = ragfin_b200.Index(768, "bf16", capacity=rows)
    for r in range(0, rows, 1_000_000):
        idx.add_synthetic(1234, r, 1_000_000)
    for k in (1000, 16384):
        for nq in (1, 16) if mode == "batched" else (1,):
            q = torch.from_numpy(synth_rows(1235, 0, nq, 768)).cuda()
            for _ in range(2):
                out = idx.search_device(q, k)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); out = idx.search_device(q, k); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            st = idx.stats()
            print(f"{mode} k={k} nq={nq}: {statistics.median(ts):.3f} ms per call = {statistics.median(ts) / nq:.3f} ms per query, launches {st['launches']}, "
                  f"redone {st['queries_rescanned']}, top id {int(out[0][0][0])} score {float(out[1][0][0]):.6f} last {float(out[1][0][k - 1]):.6f}", flush=True)
    idx.close()
