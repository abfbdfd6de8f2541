"""First-light check of the tcgen05 path: raw tensor-core scores vs a torch matmul of the same
(stored, rounded-query) operands, then end-to-end parity of the GEMM-routed search vs the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ragfin_b200
from oracle import ragfin_oracle as O, c_oracle as C

def rounded_queries(qhat, dtype):
    if dtype == "f32":
        return (qhat.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)   # tf32 truncation (model)
    return O.round_to_storage(qhat, dtype)

ok = True
for dtype, dim, n, nq in [("bf16", 768, 5000, 130), ("f16", 384, 777, 9), ("bf16", 100, 3000, 128), ("f32", 768, 2100, 40), ("bf16", 1024, 70000, 300)]:
    x = O.synth_rows(5, 0, n, dim); q = O.synth_rows(6, 0, nq, dim)
    idx = ragfin_b200.Index(dim, dtype, capacity=n); idx.add(x)
    got = idx.debug_gemm_scores(torch.from_numpy(q).cuda()); torch.cuda.synchronize()
    stored = idx.read_rows(0, n); qhat = O.normalize_rows(q, "f32")
    srows = stored if dtype != "f32" else (stored.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    ref = (torch.from_numpy(rounded_queries(qhat, dtype)).double() @ torch.from_numpy(srows).double().T).float()
    err = (got.cpu() - ref).abs().max().item()
    tol = 2e-5 if dtype != "f32" else 2e-3
    print(f"raw scores {dtype} dim={dim} n={n} nq={nq}: max abs err {err:.3e} (tol {tol})", flush=True)
    ok &= err < tol
    t0 = time.time(); ids, sc = idx.search(q, 10); dt = time.time() - t0
    st = idx.stats()
    wi, ws = C.cosine_topk(q, C.normalize_rows(x, dtype), 10)
    same = np.array_equal(ids, wi) and np.array_equal(sc.view(np.uint32), ws.view(np.uint32))
    print(f"   search path={st['path']} launches={st['launches']} rescanned={st['queries_rescanned']} parity={same} ({dt*1e3:.1f} ms)", flush=True)
    ok &= same and st["path"] == 1
print("GEMM CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
