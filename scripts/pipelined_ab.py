"""Same-GPU A/B of Index.set_pipelined on the headline workload (10M x 768 bf16, batch 1, k = 10): 200 back-to-back searches on
one stream, CUDA events around the block, the two settings alternating four times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragfin_b200
from ragfin_b200.synthetic import synth_rows

rows = int(os.environ.get("AB_ROWS", 10_000_000))
idx = ragfin_b200.Index(768, "bf16", capacity=rows)
for r in range(0, rows, 1_000_000):
    idx.add_synthetic(1234, r, min(1_000_000, rows - r))
q = torch.from_numpy(synth_rows(1235, 0, 4, 768)).cuda()
qs = [q[i:i + 1].contiguous() for i in range(4)]
out_i = torch.empty((1, 10), dtype=torch.int64, device="cuda")
out_s = torch.empty((1, 10), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for rnd in range(4):
    for on in (True, False):
        idx.set_pipelined(on)
        for i in range(5):
            idx.search_device(qs[i % 4], 10, out_ids=out_i, out_scores=out_s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200):
            idx.search_device(qs[i % 4], 10, out_ids=out_i, out_scores=out_s)
        e1.record()
        torch.cuda.synchronize()
        print(f"round {rnd} pipelined={on}: {e0.elapsed_time(e1) / 200:.4f} ms per search", flush=True)
