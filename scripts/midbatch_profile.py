"""Batches of 17-128 queries on 10M x 768 bf16, k = 10: whole-call time of the multi-kernel path (query-major tcgen05 sweep,
K3) against the one-kernel search's wide configurations, and - with `list` - two searches at one batch size for an ncu launch
list (which kernels of the chain hold the time above the HBM floor of the sweep).
    python scripts/midbatch_profile.py time
    ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/mid_launches.csv \
        python scripts/midbatch_profile.py list 128"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RAGFIN_FUSED_MAX_NQ", "64")
os.environ.setdefault("RAGFIN_FUSED_MAX_NQK", "1000000")
import torch
import ragfin_b200
from ragfin_b200.synthetic import synth_rows

mode = sys.argv[1] if len(sys.argv) > 1 else "time"
rows = int(os.environ.get("MID_ROWS", 10_000_000))
idx = ragfin_b200.Index(768, "bf16", capacity=rows)
for r in range(0, rows, 1_000_000):
    idx.add_synthetic(1234, r, min(1_000_000, rows - r))
qall = torch.from_numpy(synth_rows(1235, 0, 4 * 128, 768)).cuda()


def timed(nq, steps=12, idx=idx):
    qs = [qall[i * nq:(i + 1) * nq].contiguous() for i in range(4)]
    for i in range(3):
        idx.search_device(qs[i % 4], 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per = []
    for i in range(steps):              # one search at a time (events between them: no pipelining), 15 GB each: L2 is flushed
        e0.record()
        idx.search_device(qs[i % 4], 10)
        e1.record()
        e1.synchronize()
        per.append(e0.elapsed_time(e1))
    per = sorted(per[len(per) // 4:])      # the first quarter settles clocks
    return per[len(per) // 2], idx.stats()


if mode == "exp":      # one-kernel search, wide configurations, under experiment knobs (read when a view is created)
    configs = [dict(), dict(RAGFIN_FUSED_REFRESH_EVERY="1"), dict(RAGFIN_FUSED_REFRESH_EVERY="16"), dict(RAGFIN_FUSED_STAGES="3")]
    for cfg in configs:
        os.environ.update(cfg)
        v = idx.view()
        v.set_fused(True, 8192)
        out = []
        for nq in (1, 16, 17, 32, 64):
            ms, st = timed(nq, idx=v)
            tm = v.fused_times()
            out.append(f"nq={nq}: {ms:.3f} ms (sweep {tm[3] - tm[2]:.0f} us: wait {tm[13]:.0f} append {tm[14]:.0f} book {tm[15]:.0f})")
        print(cfg or "default", " | ".join(out), flush=True)
        print("   stamps of the last search (us):", tm, flush=True)
        v.close()
        for key in cfg:
            del os.environ[key]
elif mode == "list":
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    idx.set_fused(False, 8192)
    q = qall[:nq].contiguous()
    for _ in range(2):
        ids, sc = idx.search_device(q, 10)
    torch.cuda.synchronize()
    print("list", nq, idx.stats(), ids[0, :3].tolist())
else:
    # the two paths in separate blocks (the multi-kernel path's padded 128-row MMAs leave the GPU power-capped for the searches
    # that follow: alternating per batch size measures the neighbour's heat), twice, 40 searches per point
    import time
    res = {}
    for rnd in range(2):
        for fused in (False, True):
            idx.set_fused(fused, 8192)
            time.sleep(1.5)
            for nq in (16, 17, 32, 48, 64, 96, 128):
                if fused and nq > 64:
                    continue
                ms, st = timed(nq, steps=40)
                res.setdefault((nq, fused), []).append((ms, st["path"]))
    for nq in (16, 17, 32, 48, 64, 96, 128):
        line = f"nq={nq:4d}  multi-kernel " + " / ".join(f"{m:.3f}" for m, _ in res[(nq, False)]) + f" ms (path {res[(nq, False)][0][1]})"
        if (nq, True) in res:
            line += " | one-kernel " + " / ".join(f"{m:.3f}" for m, _ in res[(nq, True)]) + f" ms (path {res[(nq, True)][0][1]})"
        print(line, flush=True)
