"""Latency of one search call vs query-batch size on a synthetic corpus, both dispatch paths (CUDA events, median)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, statistics, json
import torch, ragfin_b200
from ragfin_b200.synthetic import synth_rows

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--k", default="10", help="comma-separated k values")
ap.add_argument("--cluster", default="0", help="comma-separated tcgen05 cluster sizes (0 = automatic)")
ap.add_argument("--out", default="gpurun_out/batch_sweep.json")
ap.add_argument("--variant", default="0", help="comma-separated tcgen05 kernel variants (0 automatic, 1 streaming, 2 A-stationary, 3 swapped roles for <= 16 queries, 4 2-SM MMA pairs)")
ap.add_argument("--batches", default="1,2,3,4,5,6,8,9,12,16,32,64,128,256,512,1024,2048,4096")
ap.add_argument("--paths", default="auto")
a = ap.parse_args()
idx = ragfin_b200.Index(a.dim, a.dtype, capacity=a.rows)
for r in range(0, a.rows, 1_000_000):
    idx.add_synthetic(1234, r, min(1_000_000, a.rows - r))
out = []
for path in a.paths.split(","):
    idx.set_gemm_min_batch({"auto": 0, "scan": 1 << 30, "gemm": 2}.get(path, 9) if not path.isdigit() else int(path))
    for variant in [int(x) for x in a.variant.split(",")]:
        idx.set_gemm_variant(variant)
        for cluster in [int(x) for x in a.cluster.split(",")]:
            idx.set_gemm_cluster(cluster)
            for k in [int(x) for x in a.k.split(",")]:
                for b in [int(x) for x in a.batches.split(",")]:
                    if path == "scan" and b > 16:
                        continue
                    q = torch.from_numpy(synth_rows(1235, 0, b, a.dim)).cuda()
                    for _ in range(3):
                        idx.search_device(q, k)
                    ts = []
                    for _ in range(7):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); idx.search_device(q, k); e1.record(); torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    ms = statistics.median(ts)
                    st = idx.stats()
                    out.append({"path": path, "used": st["path"], "variant": variant, "cluster": cluster, "k": k, "batch": b,
                                "ms": round(ms, 3), "qps": round(b / ms * 1e3, 1), "rescanned": st["queries_rescanned"]})
                    print(out[-1], flush=True)
json.dump(out, open(a.out, "w"), indent=1)
