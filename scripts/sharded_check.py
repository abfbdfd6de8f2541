"""Multi-GPU parity: run under torchrun (one rank per GPU).  Every rank ingests its row shard of the same
synthetic matrix, the sharded search (local exact top-k -> NCCL all-gather -> reduce) must equal the oracle's
unsharded answer bit for bit on every rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import ragfin_b200
from oracle import c_oracle as C, ragfin_oracle as O
from ragfin_b200.sharded import ShardedSearcher, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
CASES = [("bf16", 768, 50001, 1, 10), ("f16", 768, 30000, 4, 100), ("f32", 384, 9000, 40, 5),
         ("bf16", 768, 7, 3, 10), ("bf16", 1024, 40000, 300, 10),
         ("bf16", 128, 600000, 40, 10), ("f16", 64, 900000, 5, 100),    # shards large enough for append mode
         ("bf16", 768, 200000, 1, 10), ("bf16", 768, 200000, 16, 10), ("f32", 384, 160000, 3, 10),
         ("f16", 256, 300000, 64, 5), ("bf16", 768, 100000, 2, 128)]   # one-kernel sharded search on every rank
for dtype, dim, n, nq, k in CASES * int(os.environ.get("SHARDED_CHECK_REPEAT", "1")):   # repeats: hunting intermittent failures
    row0, cnt = shard_bounds(n, world, rank)
    idx = ragfin_b200.Index(dim, dtype, capacity=max(cnt, 1), device=local)
    if cnt:
        idx.add_synthetic(500, row0, cnt, dup_every=53)
    idx.set_id_base(row0)
    q = O.synth_rows(501, 0, nq, dim)
    qd = torch.from_numpy(q).cuda()
    wi, ws = C.cosine_topk(q, C.normalize_rows(O.synth_rows(500, 0, n, dim, dup_every=53), dtype), k)
    for p2p in (True, False):      # peer-memory exchange (CUDA IPC + NVLink stores), then NCCL all-gather + reduce
        s = ShardedSearcher.for_index(idx, p2p=p2p)
        same = True

        def check(gi, gs, what):
            good = np.array_equal(gi, wi) and np.array_equal(gs.view(np.uint32), ws.view(np.uint32))
            if not good:           # say where: which call, which rank, the first query that differs
                bad = [j for j in range(nq) if not (np.array_equal(gi[j], wi[j]) and np.array_equal(gs[j].view(np.uint32), ws[j].view(np.uint32)))]
                j = bad[0]
                col = [c for c in range(k) if gi[j, c] != wi[j, c] or gs[j, c].view(np.uint32) != ws[j, c].view(np.uint32)]
                print(f"  MISMATCH rank {rank} {what} p2p={p2p}: {len(bad)}/{nq} queries differ; query {j} columns {col[:8]}: got ids {gi[j, col[:6]].tolist()} "
                      f"scores {gs[j, col[:6]].tolist()} want ids {wi[j, col[:6]].tolist()} scores {ws[j, col[:6]].tolist()}", flush=True)
            return good

        for step in range(3):      # several steps: the exchange double-buffers by step parity
            ids, sc = s.search(qd, k)
            torch.cuda.synchronize()
            same &= check(ids.cpu().numpy(), sc.cpu().numpy(), f"device call {step}")
        hi, hs = s.search_host(q, k)   # host buffers in and out (the serving call)
        same &= check(hi, hs, "host call")
        hit = torch.empty((nq, k), dtype=torch.int64).pin_memory()   # ... and with the caller's pinned torch buffers
        hst = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        s.search_host(torch.from_numpy(q).pin_memory(), k, out_ids=hit, out_scores=hst)
        same &= check(hit.numpy(), hst.numpy(), "host call, pinned torch buffers")
        if s.exchange is not None and s.fused_ok(nq, k):
            # pipelined: a burst of searches with no synchronisation in between - two searches of a rank in flight, ranks drifting
            # apart by up to a step, tiny shards (the sweep is shorter than the finalize): the gather ring's worst case
            idx.set_pipelined(True)
            burst = [s.search(qd, k) for _ in range(24)]
            torch.cuda.synchronize()
            for step, (ids, sc) in enumerate(burst):
                same &= check(ids.cpu().numpy(), sc.cpu().numpy(), f"pipelined burst call {step}")
            idx.set_pipelined(False)
        ok &= same
        if rank == 0:
            how = ("one kernel (sweep + peer exchange + reduce)" if s.exchange is not None and s.fused_ok(nq, k) else
                   "peer-memory push / merge kernels" if s.exchange is not None else "nccl all-gather + reduce")
            print(f"sharded x{world} {dtype} dim={dim} n={n} nq={nq} k={k} [{how}]: parity={same}", flush=True)
        dist.barrier()
        if s.exchange is not None:
            s.exchange.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED CHECK", "OK" if t.item() == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
