"""Row-sharded search over the GPUs of one box: one process per GPU, torch.distributed plumbing.

Partition: rank r owns the contiguous rows [r * ceil(N / W), ...); its local row 0
has global id `row0`, so "lower id wins" survives the merge.  Every rank searches its shard
(each rank's hits are already EXACT and ordered), the per-rank [nq, k] hit lists are
all-gathered (NCCL over NVLink on GPUs; gloo in the CPU tests) and every rank reduces the
[W, nq, k] buffer to the global top-k.  This replaces the querynode -> proxy reduce a Milvus
deployment of the reference would do over gRPC.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of `rank`: (row0, n_local)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} not in [0, {world})")
    per = -(-n_rows // world)
    row0 = min(rank * per, n_rows)
    return row0, min(per, n_rows - row0)


class ShardedSearcher:
    """Search = local top-k on this rank's shard -> all-gather -> global reduce.

    `local_search(queries, k) -> (ids, scores)` and `merge(ids, scores, parts, k)` are injected so
    that the host logic can be exercised on CPU (gloo) with stand-ins; on a GPU box they are
    `Index.search_device` and `ragfin_b200.merge_topk` (see `for_index`).  With `packed_factory`
    (GPU path) the local hits are written straight into one packed record per rank, so the exchange
    is a single NCCL all-gather of nq*k*12 bytes."""

    def __init__(self, local_search: Callable, merge: Callable, group=None, packed_factory: Optional[Callable] = None,
                 exchange=None, index=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local_search = local_search
        self.merge = merge
        self.packed_factory = packed_factory
        self.exchange = exchange          # PeerExchange: stores over NVLink peer memory instead of the NCCL all-gather
        self.index = index                # with an exchange: shapes every rank can serve with the one-kernel search take
        self._fused = {}                  #   ragfin_search_sharded (sweep + exchange + reduce in ONE kernel per GPU)
        self._local = {}
        self._packed = {}
        self._gather_ids = None
        self._gather_scores = None

    @classmethod
    def for_index(cls, index, group=None, p2p: Optional[bool] = None) -> "ShardedSearcher":
        """GPU wiring.  p2p: True = peer-memory exchange (raises if unavailable), False = NCCL all-gather + reduce,
        None = peer memory unless env RAGFIN_P2P=0, falling back to NCCL if it cannot be set up.  With peer memory, shapes
        that every rank serves with the one-kernel search (<= 64 queries, k <= 128) run sweep + exchange + reduce in ONE
        kernel per GPU (`ragfin_search_sharded`); other shapes use the push / merge kernels or the NCCL all-gather."""
        import os
        import torch.distributed as dist
        from .engine import PackedHits, PeerExchange, merge_topk
        exchange = None
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if p2p is None:
            p2p = os.environ.get("RAGFIN_P2P", "1") == "1"
            required = False
        else:
            required = bool(p2p)
        if p2p and world > 1:
            try:
                exchange = PeerExchange(index.device, group=group)
            except RuntimeError:
                if required:
                    raise
        return cls(index.search_device, merge_topk, group, packed_factory=PackedHits, exchange=exchange, index=index)

    def fused_ok(self, nq: int, k: int) -> bool:
        """True when EVERY rank serves (nq, k) with the one-kernel search (decided once per shape by a MIN all-reduce: shard
        sizes differ by a row, and a rank taking another path would leave its peers waiting in the kernel)."""
        key = (nq, k)
        ok = self._fused.get(key)
        if ok is None:
            import torch
            mine = 1 if (self.exchange is not None and self.index is not None and nq * ((k * 12 + 15) // 16 * 16) <= self.exchange.max_record_bytes
                         and self.index.fused_eligible(nq, k)) else 0
            on_gpu = self.index is not None and torch.cuda.is_available()
            t = torch.tensor([mine], dtype=torch.int32, device=torch.device("cuda", self.index.device) if on_gpu else None)
            self._dist.all_reduce(t, op=self._dist.ReduceOp.MIN, group=self.group)
            ok = self._fused[key] = bool(int(t.item()))
        return ok

    def search_host(self, queries, k: int, out_ids=None, out_scores=None):
        """Host buffers in and out (numpy arrays, lists or torch CPU tensors; results come back in `out_ids` / `out_scores`
        when given, else as numpy), synchronous: the call a serving process makes.  One-kernel path when every rank can take
        it, else device staging around `search`."""
        import numpy as np
        import torch
        if not hasattr(queries, "shape"):
            queries = np.asarray(queries, dtype=np.float32)
        nq = queries.shape[0] if len(queries.shape) == 2 else 1
        if self.world > 1 and self.exchange is not None and self.fused_ok(nq, k):
            return self.exchange.search_sharded_host(self.index, queries, k, out_ids, out_scores)   # numpy or torch CPU buffers
        q = (queries.detach().to("cpu", torch.float32).numpy() if hasattr(queries, "data_ptr")
             else np.ascontiguousarray(queries, dtype=np.float32))
        if q.ndim == 1:
            q = q[None, :]
        out_ids = out_ids.numpy() if hasattr(out_ids, "data_ptr") else out_ids
        out_scores = out_scores.numpy() if hasattr(out_scores, "data_ptr") else out_scores
        dev = torch.device("cuda", self.index.device) if (self.index is not None and torch.cuda.is_available()) else None
        ids, sc = self.search(torch.from_numpy(q).to(dev) if dev is not None else torch.from_numpy(q), k)
        ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
        if out_ids is not None:
            out_ids[...] = ids
            ids = out_ids
        if out_scores is not None:
            out_scores[...] = sc
            sc = out_scores
        return ids, sc

    def search(self, queries, k: int):
        import torch
        if self.world == 1:
            return self.local_search(queries, k)
        nq = queries.shape[0]
        if self.exchange is not None and self.index is not None and self.fused_ok(nq, k):
            return self.exchange.search_sharded(self.index, queries, k)
        if self.exchange is not None and nq * k * 12 <= self.exchange.max_record_bytes:
            key = (nq, k)
            buf = self._local.get(key)
            if buf is None:
                buf = self._local[key] = (torch.empty((nq, k), dtype=torch.int64, device=queries.device),
                                          torch.empty((nq, k), dtype=torch.float32, device=queries.device))
            self.local_search(queries, k, out_ids=buf[0], out_scores=buf[1])
            return self.exchange.allgather_merge(buf[0], buf[1])
        if self.packed_factory is not None:
            key = (nq, k)
            ph = self._packed.get(key)
            if ph is None:
                ph = self._packed[key] = self.packed_factory(nq, k, self.world, queries.device)
            self.local_search(queries, k, out_ids=ph.ids, out_scores=ph.scores)
            self._dist.all_gather_into_tensor(ph.gathered, ph.local, group=self.group)
            return ph.merge()
        ids, scores = self.local_search(queries, k)
        shape = (self.world * nq, k)          # concatenation along dim 0 == [world][nq][k] in memory
        if self._gather_ids is None or tuple(self._gather_ids.shape) != shape or self._gather_ids.device != ids.device:
            self._gather_ids = torch.empty(shape, dtype=ids.dtype, device=ids.device)
            self._gather_scores = torch.empty(shape, dtype=scores.dtype, device=scores.device)
        self._dist.all_gather_into_tensor(self._gather_ids, ids.contiguous(), group=self.group)
        self._dist.all_gather_into_tensor(self._gather_scores, scores.contiguous(), group=self.group)
        return self.merge(self._gather_ids.view(self.world, nq, k), self._gather_scores.view(self.world, nq, k),
                          self.world, k)
