"""The N > 1 path on the CPU: world_size-2/4 `gloo` processes check that item-aligned row shards
plus a sum all-reduce of the per-node partials reproduce the single-shard result -- the logic the
engine's NCCL all-reduce implements on GPUs.  The per-shard scores come from the oracle here
(there is no GPU in this container); the sharding / reduction code is the product's."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from fetching_b200 import distributed as pfd, models
    from oracle import oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data, pos = models.mixture_dataset(23000, 8, 2, seed=5)   # 23 items of 1000 rows: uneven over 2 and 4 ranks
    thetas = models.mixture_thetas(pos, 5, seed=2)
    lo, hi = pfd.shard_rows(len(data), 1000, rank, world)
    assert lo % 1000 == 0 and hi % 1000 == 0
    s, q2 = orc.mixture_frontier(data[lo:hi], thetas, 1000)
    s, q2 = pfd.combine_partials_host(s, q2)
    full = orc.mixture_frontier(data, thetas, 1000)
    ok = bool(np.max(np.abs(s - full[0]) / np.abs(full[0])) < 1e-12 and np.max(np.abs(q2 - full[1]) / np.abs(full[1])) < 1e-12)
    # lasso: 10 items of 500 rows
    X, y, beta = models.blasso_dataset(5000, 12, seed=2)
    th = models.blasso_thetas(beta, 3, seed=1)
    lo, hi = pfd.shard_rows(5000, 500, rank, world)
    s, q2 = pfd.combine_partials_host(*orc.blasso_frontier(X[lo:hi], y[lo:hi], th, 500))
    full = orc.blasso_frontier(X, y, th, 500)
    ok = ok and bool(np.max(np.abs(s - full[0]) / np.abs(full[0])) < 1e-12 and
                     np.max(np.abs(q2 - full[1]) / np.abs(full[1])) < 1e-12)
    q.put((rank, ok, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


class _FakeEngine:
    """Stands in for FrontierEngine in the handle exchange of init_engine_peers (no GPU here)."""

    def __init__(self, rank):
        self.rank, self.attached = rank, None

    def peer_export(self):
        return bytes([self.rank]) * 64

    def peer_attach(self, handles, rank, nranks):
        self.attached = (list(handles), rank, nranks)


def _peer_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from fetching_b200 import distributed as pfd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine(rank)
    pfd.init_engine_peers(eng)
    handles, r, n = eng.attached
    ok = r == rank and n == world and len(handles) == world and all(h == bytes([k]) * 64 for k, h in enumerate(handles))
    # the node-split bookkeeping the in-kernel gather relies on: output node b = record b // per, column b % per
    for B in (3, 13, 64):
        per = -(-B // world)
        cuts = pfd.node_slices(B, world)
        ok = ok and all(cuts[b // per][0] + b % per == b for b in range(B))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_handles_reach_every_rank_in_rank_order(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world * 11 + (os.getpid() % 200)
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r for r, _ in res] == list(range(world)) and all(ok for _, ok in res)


@pytest.mark.parametrize("world", [2, 4])
def test_row_shards_plus_allreduce_equal_single_shard(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world * 7 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    spans = [span for _, _, span in res]
    assert spans[0][0] == 0 and spans[-1][1] == 5000 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_shard_items_partition():
    from fetching_b200.distributed import shard_items, shard_rows
    for n in (1, 7, 100, 100000):
        for w in (1, 2, 3, 8):
            cuts = [shard_items(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    assert shard_rows(100_000_000, 1000, 3, 8) == (37_500_000, 50_000_000)


def test_node_slices_cover_every_node_once():
    from fetching_b200.distributed import node_slices
    for B in (1, 3, 13, 64, 65):
        for w in (1, 2, 4, 8):
            cuts = node_slices(B, w)
            assert len(cuts) == w and cuts[0][0] == 0 and cuts[-1][1] == B
            covered = [b for lo, hi in cuts for b in range(lo, hi)]
            assert covered == list(range(B))
            assert max(hi - lo for lo, hi in cuts) == -(-B // w)
