"""Views (ragfin_create_view) and pipelined batch-1 searches.  Written after round 1's GPU budget; first run at the start of
round 2 (`profiles/r02/experiment_views_pipeline_first_run.log`: equal results, 0.284 ms per query on three streams vs 0.327).

  parity   searches through a view return exactly what the parent returns (and the oracle), a view rejects add()
  overlap  10M x 768 and 1.25M x 768 bf16, batch 1: K calls back to back on one stream against the same K calls
           round-robined over 2 and 3 handles (parent + views), each on its own stream - the latency-bound head and
           tail of one call (prep, bound pass, finalize: ~60-85 us) can overlap the neighbour's sweep.  Reported as
           PIPELINED throughput next to the single-stream figure; results of every call are compared.

    gpurun --timeout 600 -- 'python scripts/pipeline_check.py > gpurun_out/pipeline_check.log 2>&1'
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stage_parity():
    import numpy as np
    import ragfin_b200
    from oracle import ragfin_oracle as O, c_oracle as C
    x = O.synth_rows(7, 0, 70000, 128, dup_every=61)
    q = O.synth_rows(8, 0, 5, 128)
    want = C.cosine_topk(q, C.normalize_rows(x, "bf16"), 10)
    idx = ragfin_b200.Index(128, "bf16", capacity=80000)
    idx.add(x)
    v = idx.view()
    ok = len(v) == len(idx) == 70000
    for h in (idx, v):
        for min_batch in (1, 1 << 30):
            h.set_gemm_min_batch(min_batch)
            ids, sc = h.search(q, 10)
            ok &= bool(np.array_equal(ids, want[0]) and np.array_equal(sc.view(np.uint32), want[1].view(np.uint32)))
    try:
        v.add(x[:1])
        ok = False
    except ragfin_b200.RagfinError as e:
        ok &= e.code == -4
    idx.add(x[:10])                      # the parent keeps growing; the view keeps its snapshot
    ok &= len(idx) == 70010 and len(v) == 70000
    v.close()
    idx.close()
    print("view parity:", ok, flush=True)
    return ok


def stage_overlap():
    import torch
    import ragfin_b200
    from ragfin_b200.synthetic import synth_rows
    K = 200
    for rows in (1_250_000, 10_000_000):
        idx = ragfin_b200.Index(768, "bf16", capacity=rows)
        for r in range(0, rows, 1_000_000):
            idx.add_synthetic(1234, r, min(1_000_000, rows - r))
        q = torch.from_numpy(synth_rows(1235, 0, 8, 768)).cuda()
        ref = [idx.search_device(q[i % 8:i % 8 + 1], 10) for i in range(8)]
        torch.cuda.synchronize()
        for depth in (1, 2, 3):
            handles = [idx] + [idx.view() for _ in range(depth - 1)]
            streams = [torch.cuda.Stream() for _ in range(depth)]
            outs = [(torch.empty((1, 10), dtype=torch.int64, device="cuda"), torch.empty((1, 10), dtype=torch.float32, device="cuda"))
                    for _ in range(8 * depth)]
            def run(n):
                for i in range(n):
                    s = i % depth
                    with torch.cuda.stream(streams[s]):
                        o = outs[i % (8 * depth)]
                        handles[s].search_device(q[i % 8:i % 8 + 1], 10, out_ids=o[0], out_scores=o[1], stream=streams[s])
            run(4 * depth); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in streams:
                s.wait_event(e0)
            run(K)
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            same = all(torch.equal(outs[i % (8 * depth)][0], ref[i % 8][0]) and torch.equal(outs[i % (8 * depth)][1], ref[i % 8][1])
                       for i in range(K - 8 * depth, K))
            print(f"rows={rows} handles/streams={depth}: {ms:.4f} ms per query = {1e3 / ms:.1f} queries/s, results equal: {same}", flush=True)
            for h in handles[1:]:
                h.close()
        idx.close()
    return True


STAGES = {"parity": (stage_parity, 180), "overlap": (stage_overlap, 360)}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(0 if STAGES[sys.argv[1]][0]() else 1)
    for name, (_fn, limit) in STAGES.items():
        try:
            rc = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=limit).returncode
        except subprocess.TimeoutExpired:
            rc = -9
        print(f"PIPELINE CHECK stage {name}: {'OK' if rc == 0 else 'FAILED rc=%d' % rc}", flush=True)
        if rc != 0:
            sys.exit(1)
    print("PIPELINE CHECK OK", flush=True)
