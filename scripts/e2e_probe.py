"""Where does the host path (ragfin_search_host) spend its time at batch 4096?  H2D / search / D2H measured separately."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, ragfin_b200
from ragfin_b200.synthetic import synth_rows

rows, dim, k, b = 10_000_000, 768, 10, 4096
idx = ragfin_b200.Index(dim, "bf16", capacity=rows)
for r in range(0, rows, 1_000_000):
    idx.add_synthetic(1234, r, 1_000_000)
qh = torch.from_numpy(synth_rows(1235, 0, b, dim)).pin_memory()
qd = qh.cuda()
oi = torch.empty((b, k), dtype=torch.int64).pin_memory()
os_ = torch.empty((b, k), dtype=torch.float32).pin_memory()
qpage = qh.numpy().copy()          # pageable
def wall(f, n=5):
    f(); torch.cuda.synchronize()
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); t.append((time.perf_counter() - t0) * 1e3)
    return round(min(t), 3), round(sorted(t)[len(t) // 2], 3)
print("H2D 12.6 MB pinned (torch copy_)      ms min/median:", wall(lambda: qd.copy_(qh, non_blocking=True)))
print("search_device (queries resident)       ms min/median:", wall(lambda: idx.search_device(qd, k)))
print("search host, pinned in/out             ms min/median:", wall(lambda: idx.search(qh.numpy(), k, out_ids=oi.numpy(), out_scores=os_.numpy())))
print("search host, pageable in, fresh out    ms min/median:", wall(lambda: idx.search(qpage, k)))
ids, sc = idx.search_device(qd, k)
print("D2H ids+scores pinned                  ms min/median:", wall(lambda: (oi.copy_(ids, non_blocking=True), os_.copy_(sc, non_blocking=True))))
# back-to-back host calls: does the sync between calls let the clock drop / rise?
t0 = time.perf_counter()
for _ in range(10):
    idx.search(qh.numpy(), k, out_ids=oi.numpy(), out_scores=os_.numpy())
print("10 host calls back to back, ms per call:", round((time.perf_counter() - t0) * 100, 3))
t0 = time.perf_counter()
for _ in range(10):
    idx.search_device(qd, k)
torch.cuda.synchronize()
print("10 device calls back to back, ms per call:", round((time.perf_counter() - t0) * 100, 3))
