"""Sustained (power-capped) A/B of the large-batch kernels: blocks of 12 back-to-back searches, the configurations alternating
four times on the same GPU.  10M x 768 bf16, k = 10."""
import os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragfin_b200
from ragfin_b200.synthetic import synth_rows
rows = 10_000_000
idx = ragfin_b200.Index(768, "bf16", capacity=rows)
for r in range(0, rows, 1_000_000):
    idx.add_synthetic(1234, r, 1_000_000)
configs = [("single-CTA MMA, pairs", 1, 2), ("2-SM MMA pairs", 4, 2)]
for b in (4096, 1024):
    q = torch.from_numpy(synth_rows(1235, 0, b, 768)).cuda()
    ts = {n: [] for n, _, _ in configs}
    for rep in range(4):
        for name, variant, cluster in configs:
            idx.set_gemm_variant(variant); idx.set_gemm_cluster(cluster)
            for _ in range(3):
                idx.search_device(q, 10)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(12):
                idx.search_device(q, 10)
            e1.record(); torch.cuda.synchronize()
            ts[name].append(e0.elapsed_time(e1) / 12)
    print(f"batch {b} sustained: " + " | ".join(f"{n}: {[round(v, 2) for v in ts[n]]} median {statistics.median(ts[n]):.2f} ms" for n in ts), flush=True)
