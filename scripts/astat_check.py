"""A-stationary tcgen05 variant: raw scores vs matmul, search parity vs the oracle, timing vs the streaming variant."""
import sys, os, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ragfin_b200
from oracle import ragfin_oracle as O, c_oracle as C
from ragfin_b200.synthetic import synth_rows

ok = True
for dtype, dim, n, nq in [("bf16", 768, 5000, 130), ("f16", 384, 777, 9), ("bf16", 100, 3000, 128), ("f16", 768, 20000, 300)]:
    x = O.synth_rows(5, 0, n, dim); q = O.synth_rows(6, 0, nq, dim)
    idx = ragfin_b200.Index(dim, dtype, capacity=n); idx.add(x); idx.set_gemm_variant(2)
    got = idx.debug_gemm_scores(torch.from_numpy(q).cuda()); torch.cuda.synchronize()
    stored = idx.read_rows(0, n); qr = O.round_to_storage(O.normalize_rows(q, "f32"), dtype)
    ref = (torch.from_numpy(qr).double() @ torch.from_numpy(stored).double().T).float()
    err = (got.cpu() - ref).abs().max().item()
    print(f"astat raw scores {dtype} dim={dim} n={n} nq={nq}: max abs err {err:.3e}", flush=True)
    ok &= err < 2e-6
    for cl in (1, 2, 4):
        idx.set_gemm_cluster(cl)
        ids, sc = idx.search(q, 10)
        wi, ws = C.cosine_topk(q, C.normalize_rows(x, dtype), 10)
        same = np.array_equal(ids, wi) and np.array_equal(sc.view(np.uint32), ws.view(np.uint32))
        print(f"   cluster {cl}: parity={same} path={idx.stats()['path']}", flush=True)
        ok &= same
    idx.close()
print("ASTAT CHECK", "OK" if ok else "FAILED", flush=True)
if not ok:
    sys.exit(1)
idx = ragfin_b200.Index(768, "bf16", capacity=10_000_000)
for r in range(0, 10_000_000, 1_000_000): idx.add_synthetic(1234, r, 1_000_000)
for variant in (1, 2):
    idx.set_gemm_variant(variant)
    for b in (5, 16, 64, 128, 256, 1024, 4096):
        q = torch.from_numpy(synth_rows(1235, 0, b, 768)).cuda()
        for _ in range(2): idx.search_device(q, 10)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); idx.search_device(q, 10); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(f"variant {variant} batch {b}: {statistics.median(ts):.3f} ms", flush=True)
