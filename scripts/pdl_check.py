"""Programmatic-dependent-launch flavour (make -C ragfin_b200/csrc pdl -> libragfin_pdl.so) against the default build.

Not yet run on a GPU (written after the round's GPU budget was spent).  For each library, in its own child process
under a timeout: (1) parity of a few searches against the C oracle on both dispatch paths, (2) per-call latency of
batch-1 / batch-2 / batch-16 / batch-128 searches on a 1.25M-row shard (the 8-GPU split, where the ~60 us of fixed
per-call cost is 15-20 % of the step) and on the full 10M rows.

    make -C ragfin_b200/csrc pdl            # here (the .so travels with the snapshot)
    gpurun --timeout 900 -- 'python scripts/pdl_check.py > gpurun_out/pdl_check.log 2>&1'
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PDL = os.path.join(ROOT, "ragfin_b200", "csrc", "libragfin_pdl.so")


def child():
    import statistics
    import numpy as np
    import torch
    import ragfin_b200
    from oracle import ragfin_oracle as O, c_oracle as C
    from ragfin_b200.synthetic import synth_rows
    ok = True
    for dtype, dim, n, nq, k in [("bf16", 768, 30000, 1, 10), ("bf16", 768, 30000, 2, 10), ("f16", 384, 20011, 16, 5),
                                 ("bf16", 128, 50000, 300, 10), ("f32", 768, 20000, 130, 10), ("bf16", 768, 3000, 4, 10)]:
        x = O.synth_rows(7, 0, n, dim)
        q = O.synth_rows(8, 0, nq, dim)
        idx = ragfin_b200.Index(dim, dtype, capacity=n)
        idx.add(x)
        wi, ws = C.cosine_topk(q, C.normalize_rows(x, dtype), k)
        for min_batch in (1, 1 << 30):          # tensor-core path / scan path
            idx.set_gemm_min_batch(min_batch)
            for _rep in range(3):               # back-to-back calls: the next call's prep overlaps this call's tail
                ids, sc = idx.search(q, k)
            same = np.array_equal(ids, wi) and np.array_equal(sc.view(np.uint32), ws.view(np.uint32))
            print(f"  parity {dtype} dim={dim} n={n} nq={nq} k={k} path={idx.stats()['path']}: {same}", flush=True)
            ok &= same
        idx.close()
    if not ok:
        return False
    for rows in (1_250_000, 10_000_000):
        idx = ragfin_b200.Index(768, "bf16", capacity=rows)
        for r in range(0, rows, 1_000_000):
            idx.add_synthetic(1234, r, min(1_000_000, rows - r))
        for b in (1, 2, 16, 128):
            q = torch.from_numpy(synth_rows(1235, 0, b, 768)).cuda()
            for _ in range(5):
                idx.search_device(q, 10)
            ts = []
            for _ in range(50):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); idx.search_device(q, 10); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            # back-to-back calls (what bench.py times): 50 calls, one pair of events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                idx.search_device(q, 10)
            e1.record(); torch.cuda.synchronize()
            print(f"  rows={rows} batch={b}: single call median {statistics.median(ts):.4f} ms, "
                  f"back-to-back {e0.elapsed_time(e1) / 50:.4f} ms", flush=True)
        idx.close()
    return True


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        sys.exit(0 if child() else 1)
    rc_all = 0
    for name, lib in (("default", None), ("pdl", PDL)):
        if lib and not os.path.exists(lib):
            print(f"{lib} missing: run `make -C ragfin_b200/csrc pdl` first", flush=True)
            sys.exit(2)
        env = dict(os.environ)
        env.pop("RAGFIN_LIB", None)
        if lib:
            env["RAGFIN_LIB"] = lib
        print(f"== {name} build", flush=True)
        try:
            rc = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, timeout=400).returncode
        except subprocess.TimeoutExpired:
            rc = -9
        print(f"PDL CHECK {name}: {'OK' if rc == 0 else 'FAILED rc=%d' % rc}", flush=True)
        rc_all |= int(rc != 0)
    sys.exit(rc_all)
