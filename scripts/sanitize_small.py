"""Small-shape run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck): ingest, scan
(1/2/4 queries), GEMM (cluster 1/2/4, f16 and tf32 kinds), finalize, exact tier, large-k, merge, save/load."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ragfin_b200
from oracle import ragfin_oracle as O, c_oracle as C

ok = True
def check(tag, got, want):
    global ok
    same = np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))
    ok &= same
    print(tag, "ok" if same else "MISMATCH", flush=True)

for dtype, dim in (("bf16", 768), ("f32", 100)):
    x = O.synth_rows(1, 0, 3000, dim, dup_every=13, zero_every=97)
    x[100:160] = x[7]                      # forces the exact tier for the query below
    st = C.normalize_rows(x, dtype)
    idx = ragfin_b200.Index(dim, dtype, capacity=4000)
    idx.add(x[:1000]); idx.add(torch.from_numpy(x[1000:]).cuda())
    q = O.synth_rows(2, 0, 300, dim); q[0] = x[7]
    for nq in (1, 2, 3, 4):
        check(f"{dtype} scan nq={nq}", idx.search(q[:nq], 10), C.cosine_topk(q[:nq], st, 10))
    for c in (1, 2, 4):
        idx.set_gemm_cluster(c)
        check(f"{dtype} gemm cluster={c}", idx.search(q, 10), C.cosine_topk(q, st, 10))
    check(f"{dtype} gemm k=100", idx.search(q[:40], 100), C.cosine_topk(q[:40], st, 100))
    check(f"{dtype} large-k", idx.search(q[:2], 500), C.cosine_topk(q[:2], st, 500))
    qd = torch.from_numpy(q[:6]).cuda()
    a = idx.search_device(qd, 5)
    mi, ms = ragfin_b200.merge_topk(torch.stack([a[0], a[0] + 5000]), torch.stack([a[1], a[1]]), 2, 5)
    torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as d:
        idx.save(os.path.join(d, "m.ragfin"))
        back = ragfin_b200.Index.load(os.path.join(d, "m.ragfin"))
        check(f"{dtype} reload", back.search(q[:3], 10), C.cosine_topk(q[:3], st, 10))
        back.close()
    idx.close()
print("SANITIZE RUN", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
