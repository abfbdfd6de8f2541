"""Timing of the one-kernel search (csrc/sweep_fused.cuh) against the multi-kernel paths: whole-call latency (CUDA events,
median) on the 10M-row corpus and on the 1.25M-row shard of the 8-GPU split, batches 1..64, k = 10 and 100.
    gpurun --timeout 900 -- 'python scripts/fused_check.py > gpurun_out/fused_check.log 2>&1'"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ragfin_b200
from ragfin_b200.synthetic import synth_rows

dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
for rows in (1_250_000, 10_000_000):
    idx = ragfin_b200.Index(768, dtype, capacity=rows)
    for r in range(0, rows, 1_000_000):
        idx.add_synthetic(1234, r, min(1_000_000, rows - r))
    for k in (10, 100):
        for nq in (1, 2, 8, 16, 32, 64):
            q = torch.from_numpy(synth_rows(1235, 0, nq, 768)).cuda()
            res = {}
            for fused in (True, False):
                idx.set_fused(fused)
                for _ in range(3):
                    out = idx.search_device(q, k)
                torch.cuda.synchronize()
                st = idx.stats()
                ts = []
                for _ in range(15):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    out = idx.search_device(q, k)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    out = idx.search_device(q, k)
                e1.record()
                torch.cuda.synchronize()
                res[fused] = (statistics.median(ts), e0.elapsed_time(e1) / 20, st, out)
            same = torch.equal(res[True][3][0], res[False][3][0]) and torch.equal(res[True][3][1], res[False][3][1])
            f, o = res[True], res[False]
            print(f"rows={rows} {dtype} k={k} nq={nq}: fused {f[0]:.4f} ms (back-to-back {f[1]:.4f}, path {f[2]['path']}, launches {f[2]['launches']}, "
                  f"rescanned {f[2]['queries_rescanned']}) | multi-kernel {o[0]:.4f} ms (back-to-back {o[1]:.4f}, path {o[2]['path']}, launches {o[2]['launches']}) "
                  f"| equal: {same}", flush=True)
    idx.close()
