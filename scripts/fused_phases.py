"""Phase times of the one-kernel search (globaltimer stamps inside the kernel), diagnostics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragfin_b200
from ragfin_b200.synthetic import synth_rows
for rows in (1_250_000, 10_000_000):
    idx = ragfin_b200.Index(768, "bf16", capacity=rows)
    for r in range(0, rows, 1_000_000):
        idx.add_synthetic(1234, r, min(1_000_000, rows - r))
    for nq, k in ((1, 10), (8, 10), (16, 10), (32, 10), (64, 10), (1, 100), (16, 100)):
        q = torch.from_numpy(synth_rows(1235, 0, nq, 768)).cuda()
        for _ in range(3):
            idx.search_device(q, k)
        torch.cuda.synchronize()
        print(f"rows={rows} nq={nq} k={k}: phases us {idx.fused_times()}", flush=True)
    idx.close()
