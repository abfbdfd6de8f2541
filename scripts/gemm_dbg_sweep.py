"""Timing experiments on the tcgen05 path: which stage of the pipeline bounds the mid-batch regime?
RAGFIN_GEMM_DEBUG bits: 1 skip the epilogue filter, 2 skip the MMAs, 4 skip the A (query tile) loads, 8 skip the B loads.
Results with any bit set are INVALID (timing only).
The library honours the variable only when built with -DRAGFIN_TIMING_EXPERIMENTS (make NVFLAGS+=...); the shipped
build ignores it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, statistics, json
import torch, ragfin_b200
from ragfin_b200.synthetic import synth_rows

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--batches", default="16,128")
ap.add_argument("--modes", default="0,1,2,3,4,5,7,8,11")
a = ap.parse_args()
idx = ragfin_b200.Index(a.dim, a.dtype, capacity=a.rows)
for r in range(0, a.rows, 1_000_000):
    idx.add_synthetic(1234, r, min(1_000_000, a.rows - r))
idx.set_gemm_min_batch(2)
out = []
for b in [int(x) for x in a.batches.split(",")]:
    q = torch.from_numpy(synth_rows(1235, 0, b, a.dim)).cuda()
    for mode in [int(x) for x in a.modes.split(",")]:
        os.environ["RAGFIN_GEMM_DEBUG"] = str(mode)
        for _ in range(2):
            idx.search_device(q, a.k)
        idx.profile(True)
        for _ in range(5):
            idx.search_device(q, a.k)
        ms, n = idx.profile_read()
        idx.profile(False)
        out.append({"batch": b, "dbg": mode, "gemm_kernel_ms": round(ms / n, 3)})
        print(out[-1], flush=True)
os.environ["RAGFIN_GEMM_DEBUG"] = "0"
json.dump(out, open("gpurun_out/gemm_dbg_sweep.json", "w"), indent=1)
