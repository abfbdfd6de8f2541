"""First GPU run of the 2-SM MMA sweep (gemm variant 4, csrc/gemm_pair.cuh) - passed at the start of round 2
(`profiles/r02/experiment_pair_kernel_first_run.log`); the kernel is the default from 129 queries since.

The kernel was written without a GPU at hand, so every stage runs in its own child process under a timeout: a hang
(barrier protocol wrong) ends that stage, not the gpurun call.  Stages:
  raw     raw tensor-core scores of the pair kernel (dump mode) against the fp64 product of the stored operands and
          against the single-CTA kernel's dump (the accumulation order per output is the same, so expect 0 difference)
  search  exact top-k through variant 4 (append mode) against the C oracle, bf16 / f16 / f32, ragged batches
  time    10M x 768 bf16, batches 256..4096, variant 0 (automatic) against variant 4, CUDA events, median of 5

    gpurun --timeout 900 -- 'python scripts/pair_check.py > gpurun_out/pair_check.log 2>&1'
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stage_raw():
    import torch
    import ragfin_b200
    from oracle import ragfin_oracle as O
    ok = True
    for dtype, dim, n, nq in [("bf16", 768, 5000, 256), ("f16", 384, 777, 130), ("bf16", 100, 3000, 300), ("f32", 768, 4099, 257)]:
        x = O.synth_rows(5, 0, n, dim)
        q = O.synth_rows(6, 0, nq, dim)
        idx = ragfin_b200.Index(dim, dtype, capacity=n)
        idx.add(x)
        qd = torch.from_numpy(q).cuda()
        idx.set_gemm_variant(1)
        idx.set_gemm_cluster(2)
        base = idx.debug_gemm_scores(qd).cpu()
        idx.set_gemm_variant(4)
        got = idx.debug_gemm_scores(qd).cpu()
        torch.cuda.synchronize()
        stored = idx.read_rows(0, n)
        qr = O.normalize_rows(q, "f32") if dtype == "f32" else O.round_to_storage(O.normalize_rows(q, "f32"), dtype)
        ref = (torch.from_numpy(qr).double() @ torch.from_numpy(stored).double().T).float()
        err = (got - ref).abs().max().item()
        dif = (got - base).abs().max().item()
        tol = 2e-3 if dtype == "f32" else 2e-6      # tf32 truncates both operands to 10 mantissa bits
        print(f"pair raw scores {dtype} dim={dim} n={n} nq={nq}: max |pair - fp64| {err:.3e}, max |pair - single| {dif:.3e}", flush=True)
        ok &= err < tol and dif < tol
        idx.close()
    return ok


def stage_search():
    import numpy as np
    import ragfin_b200
    from oracle import ragfin_oracle as O, c_oracle as C
    ok = True
    for dtype, dim, n, nq, k in [("bf16", 768, 30000, 256, 10), ("f16", 384, 20011, 130, 5), ("bf16", 128, 50000, 1000, 10),
                                 ("f32", 768, 20000, 300, 10), ("bf16", 768, 120000, 513, 100)]:
        x = O.synth_rows(7, 0, n, dim)
        q = O.synth_rows(8, 0, nq, dim)
        idx = ragfin_b200.Index(dim, dtype, capacity=n)
        idx.add(x)
        idx.set_gemm_min_batch(2)
        idx.set_gemm_variant(4)
        ids, sc = idx.search(q, k)
        st = idx.stats()
        wi, ws = C.cosine_topk(q, C.normalize_rows(x, dtype), k)
        same = np.array_equal(ids, wi) and np.array_equal(sc.view(np.uint32), ws.view(np.uint32))
        print(f"pair search {dtype} dim={dim} n={n} nq={nq} k={k}: parity={same} path={st['path']} rescanned={st['queries_rescanned']}", flush=True)
        ok &= same and st["path"] == 1
        idx.close()
    return ok


def stage_time():
    import statistics
    import torch
    import ragfin_b200
    from ragfin_b200.synthetic import synth_rows
    idx = ragfin_b200.Index(768, "bf16", capacity=10_000_000)
    for r in range(0, 10_000_000, 1_000_000):
        idx.add_synthetic(1234, r, 1_000_000)
    ref = {}
    for variant in (0, 4):
        idx.set_gemm_variant(variant)
        for b in (256, 512, 1024, 2048, 4096):
            q = torch.from_numpy(synth_rows(1235, 0, b, 768)).cuda()
            for _ in range(2):
                ids, _sc = idx.search_device(q, 10)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ids, _sc = idx.search_device(q, 10); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            same = ""
            if variant == 0:
                ref[b] = ids.clone()
            else:
                same = f" ids equal to variant 0: {bool(torch.equal(ids, ref[b]))}"
            print(f"variant {variant} batch {b}: {ms:.3f} ms = {2 * b * 1e7 * 768 / ms / 1e9:.0f} TFLOP/s{same}", flush=True)
    return True


STAGES = {"raw": (stage_raw, 120), "search": (stage_search, 150), "time": (stage_time, 300)}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(0 if STAGES[sys.argv[1]][0]() else 1)
    for name, (_fn, limit) in STAGES.items():
        try:
            rc = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=limit).returncode
        except subprocess.TimeoutExpired:
            rc = -9
        print(f"PAIR CHECK stage {name}: {'OK' if rc == 0 else 'FAILED rc=%d' % rc}", flush=True)
        if rc != 0:
            sys.exit(1)
    print("PAIR CHECK OK", flush=True)
