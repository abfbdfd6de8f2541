"""k = 100, one query, 10M rows: event-timed single calls next to the kernel's own phase stamps (diagnostics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragfin_b200
from ragfin_b200.synthetic import synth_rows
rows = 10_000_000
idx = ragfin_b200.Index(768, "bf16", capacity=rows)
for r in range(0, rows, 1_000_000):
    idx.add_synthetic(1234, r, 1_000_000)
for k in (10, 100):
    q = torch.from_numpy(synth_rows(1235, 0, 1, 768)).cuda()
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = idx.search_device(q, k); e1.record(); torch.cuda.synchronize()
        t = idx.fused_times(); a, r = idx.fused_counts(1)
        print(f"k={k} call {i}: events {e0.elapsed_time(e1):.3f} ms | kernel stamps us: prologue {t[1]} first tile {t[2]} sweep {t[3]} arrived {t[4]} staged {t[11]} T {t[12]} selected {t[5]} rescored {t[6]} emitted {t[7]} | appended {a.tolist()} rescored {r.tolist()}", flush=True)
