"""BASELINE.json config 5: query-batch sweep at k in {1, 10, 100} over a row-sharded synthetic corpus (default
50M x 1024 bf16 on 8 GPUs).  Run under torchrun, one rank per GPU.  Latency = CUDA events around the whole sharded
search (local search + NCCL all-gather + reduce), max over ranks, median of `--reps`."""
import argparse, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ragfin_b200
from ragfin_b200.sharded import ShardedSearcher, shard_bounds
from ragfin_b200.synthetic import synth_rows

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=50_000_000)
ap.add_argument("--dim", type=int, default=1024)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--k", default="1,10,100")
ap.add_argument("--batches", default="1,4,16,64,256,1024,4096,16384")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--out", default="gpurun_out/sharded_sweep.json")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
row0, cnt = shard_bounds(a.rows, world, rank)
idx = ragfin_b200.Index(a.dim, a.dtype, capacity=max(cnt, 1), device=local)
for r in range(0, cnt, 1_000_000):
    idx.add_synthetic(1234, row0 + r, min(1_000_000, cnt - r))
idx.set_id_base(row0)
s = ShardedSearcher.for_index(idx)
esize = 4 if a.dtype == "f32" else 2
out = []
for k in [int(x) for x in a.k.split(",")]:
    for b in [int(x) for x in a.batches.split(",")]:
        q = torch.from_numpy(synth_rows(1235, 0, b, a.dim)).cuda()
        for _ in range(2):
            s.search(q, k)
        ts = []
        for _ in range(a.reps):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.search(q, k); e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        ms = statistics.median(ts)
        st = idx.stats()
        rec = {"world": world, "rows": a.rows, "dim": a.dim, "dtype": a.dtype, "k": k, "batch": b, "ms": round(ms, 3),
               "qps": round(b / ms * 1e3, 1), "path": st["path"], "rescanned": st["queries_rescanned"],
               "shard_GBps": round(cnt * a.dim * esize / ms / 1e6, 1), "TFLOPs_per_gpu": round(2.0 * b * cnt * a.dim / ms / 1e9, 1)}
        out.append(rec)
        if rank == 0:
            print(rec, flush=True)
if rank == 0:
    json.dump(out, open(a.out, "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
