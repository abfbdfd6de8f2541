"""One search per dispatch mode of the tcgen05 path (append / lists + bound pass / plain lists), for an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch, ragfin_b200
from ragfin_b200.synthetic import synth_rows

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
idx = ragfin_b200.Index(a.dim, a.dtype, capacity=a.rows)
for r in range(0, a.rows, 1_000_000):
    idx.add_synthetic(1234, r, min(1_000_000, a.rows - r))
q = torch.from_numpy(synth_rows(1235, 0, a.batch, a.dim)).cuda()
for name, app, bnd in (("append", True, True), ("lists+bound", False, True), ("lists", False, False)):
    idx.set_append_mode(app)
    idx.set_bound_pass(bnd)
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); idx.search_device(q, a.k); e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1), 3), "ms", idx.stats(), flush=True)
