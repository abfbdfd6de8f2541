"""CPU suite for the multi-rank host logic: row partition, all-gather layout and global reduce over a
world_size-2 (and 3) gloo group.  Local search and merge are oracle stand-ins (no GPU here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ragfin_oracle as O
from ragfin_b200.sharded import ShardedSearcher, shard_bounds


def test_shard_bounds_cover_rows_exactly():
    for n, w in ((10, 3), (16, 8), (7, 8), (0, 2), (10_000_000, 8), (100_000_001, 4)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert sum(c for _, c in spans) == n
        pos = 0
        for row0, cnt in spans:
            assert row0 == min(pos, n) and cnt >= 0
            pos += cnt
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = O.normalize_rows(O.synth_rows(21, 0, n, 64, dup_every=5), "bf16")
        q = O.synth_rows(22, 0, 3, 64)
        row0, cnt = shard_bounds(n, world, rank)

        def local_search(queries, kk):
            ids, sc = O.cosine_topk(queries.numpy(), x[row0:row0 + cnt], kk, id_base=row0)
            return torch.from_numpy(ids), torch.from_numpy(sc)

        def merge(ids, sc, parts, kk):            # all-gather layout [parts, nq, k]
            assert ids.shape == (parts, 3, kk)
            mi, ms = O.merge_topk([ids[p].numpy() for p in range(parts)], [sc[p].numpy() for p in range(parts)], kk)
            return torch.from_numpy(mi), torch.from_numpy(ms)

        s = ShardedSearcher(local_search, merge)
        assert s.world == world and s.rank == rank
        for _ in range(2):                        # second call reuses the gather buffers
            ids, sc = s.search(torch.from_numpy(q), k)
        want_i, want_s = O.cosine_topk(q, x, k)
        ok = np.array_equal(ids.numpy(), want_i) and np.array_equal(sc.numpy().view(np.uint32), want_s.view(np.uint32))
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,k", [(2, 101, 10), (3, 50, 20), (2, 5, 10)])
def test_sharded_search_over_gloo(world, n, k):
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n, k, out), nprocs=world, join=True)
    assert dict(out) == {r: 1 for r in range(world)}
