"""Cluster / kernel policy of the large-batch tensor-core sweep, measured interleaved (thermal and power state drift over a
run, so every configuration of a batch is timed in rotation): 10M x 768 bf16, k = 10, whole search call, CUDA events.
    gpurun --timeout 900 -- 'python scripts/policy_sweep.py > gpurun_out/policy_sweep.log 2>&1'"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragfin_b200
from ragfin_b200.synthetic import synth_rows
rows = 10_000_000
idx = ragfin_b200.Index(768, "bf16", capacity=rows)
for r in range(0, rows, 1_000_000):
    idx.add_synthetic(1234, r, 1_000_000)
configs = [("single-CTA MMA, auto cluster", 1, 0), ("single-CTA MMA, pairs", 1, 2), ("single-CTA MMA, quads", 1, 4), ("2-SM MMA pairs", 4, 2)]
out = {}
for b in (129, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096):
    q = torch.from_numpy(synth_rows(1235, 0, b, 768)).cuda()
    ts = {name: [] for name, _, _ in configs}
    ref = None
    for rep in range(4):
        for name, variant, cluster in configs:
            idx.set_gemm_variant(variant); idx.set_gemm_cluster(cluster)
            if rep == 0:
                ids, _ = idx.search_device(q, 10); torch.cuda.synchronize()
                if ref is None: ref = ids.clone()
                assert torch.equal(ids, ref), (b, name)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                idx.search_device(q, 10)
            e1.record(); torch.cuda.synchronize()
            ts[name].append(e0.elapsed_time(e1) / 2)
    res = {name: round(statistics.median(v), 3) for name, v in ts.items()}
    best = min(res, key=res.get)
    out[b] = res
    print(f"batch {b}: " + " | ".join(f"{n}: {t:.3f} ms ({2 * b * 1e7 * 768 / t / 1e9:.0f} TF/s)" for n, t in res.items()) + f"  -> best: {best}", flush=True)
print(json.dumps(out))
