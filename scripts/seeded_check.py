"""First GPU run of the self-seeded <= 16-query sweep (gemm variant 5, csrc/gemm_rows_seeded.cuh).

The kernel was written without a GPU at hand and holds a grid-wide barrier, so each stage runs in its own child
process under a timeout.  Stages:
  search  exact top-k through variant 5 against the C oracle (bf16 / f16 / f32, 1..16 queries, k 1..16, corpora whose
          slices are shorter than the sample, duplicates) - and that the variant really ran (3 launches fewer)
  time    1.25M x 768 and 10M x 768 bf16, batch 1 / 2 / 16, variant 3 against variant 5: single-call median and
          back-to-back average, CUDA events

    gpurun --timeout 900 -- 'python scripts/seeded_check.py > gpurun_out/seeded_check.log 2>&1'
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stage_search():
    import numpy as np
    import ragfin_b200
    from oracle import ragfin_oracle as O, c_oracle as C
    ok = True
    cases = [("bf16", 768, 70000, 1, 10), ("bf16", 768, 70000, 2, 1), ("f16", 128, 70000, 16, 10), ("f32", 768, 40000, 7, 16),
             ("bf16", 1024, 40000, 3, 5), ("bf16", 100, 66000, 16, 16), ("bf16", 768, 300000, 1, 10), ("f16", 384, 1250000, 4, 10)]
    for dtype, dim, n, nq, k in cases:
        x = O.synth_rows(196, 0, n, dim, dup_every=61, zero_every=1999)
        q = O.synth_rows(197, 0, nq, dim)
        if nq >= 3:
            x[4000:4030] = q[2] * 2.0        # 30 exact duplicates of one query: ties resolved by row id
        want = C.cosine_topk(q, C.normalize_rows(x, dtype), k)
        idx = ragfin_b200.Index(dim, dtype, capacity=n)
        idx.add(x)
        idx.set_gemm_min_batch(1)
        idx.set_gemm_variant(3)
        idx.search(q, k)
        base_launches = idx.stats()["launches"]
        idx.set_gemm_variant(5)
        for _rep in range(3):                # the arrival counter is monotonic across launches
            ids, sc = idx.search(q, k)
        st = idx.stats()
        same = np.array_equal(ids, want[0]) and np.array_equal(sc.view(np.uint32), want[1].view(np.uint32))
        print(f"seeded search {dtype} dim={dim} n={n} nq={nq} k={k}: parity={same} path={st['path']} launches={st['launches']} "
              f"(variant 3: {base_launches}) rescanned={st['queries_rescanned']}", flush=True)
        ok &= same and st["path"] == 1 and st["launches"] == base_launches - 2 and st["queries_rescanned"] == 0
        idx.close()
    return ok


def stage_time():
    import statistics
    import torch
    import ragfin_b200
    from ragfin_b200.synthetic import synth_rows
    for rows in (1_250_000, 10_000_000):
        idx = ragfin_b200.Index(768, "bf16", capacity=rows)
        for r in range(0, rows, 1_000_000):
            idx.add_synthetic(1234, r, min(1_000_000, rows - r))
        for variant in (3, 5):
            idx.set_gemm_variant(variant)
            for b in (1, 2, 16):
                q = torch.from_numpy(synth_rows(1235, 0, b, 768)).cuda()
                for _ in range(5):
                    idx.search_device(q, 10)
                ts = []
                for _ in range(50):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); idx.search_device(q, 10); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    idx.search_device(q, 10)
                e1.record(); torch.cuda.synchronize()
                print(f"rows={rows} variant {variant} batch {b}: single call median {statistics.median(ts):.4f} ms, "
                      f"back-to-back {e0.elapsed_time(e1) / 50:.4f} ms, launches {idx.stats()['launches']}", flush=True)
        idx.close()
    return True


STAGES = {"search": (stage_search, 150), "time": (stage_time, 300)}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(0 if STAGES[sys.argv[1]][0]() else 1)
    for name, (_fn, limit) in STAGES.items():
        try:
            rc = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=limit).returncode
        except subprocess.TimeoutExpired:
            rc = -9
        print(f"SEEDED CHECK stage {name}: {'OK' if rc == 0 else 'FAILED rc=%d' % rc}", flush=True)
        if rc != 0:
            sys.exit(1)
    print("SEEDED CHECK OK", flush=True)
