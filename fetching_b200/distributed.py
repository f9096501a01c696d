"""Row-sharded multi-GPU evaluation: one process per GPU, `torch.distributed` for the plumbing.

The reference has no data parallelism (every worker holds the whole data set).  Here a
large data set is cut into contiguous, ITEM-ALIGNED row shards (so that per-item sums of squares
stay well defined), every rank scores the whole frontier on its shard, and the 2*B partial sums
are all-reduced by NCCL inside the engine (pf_engine_comm_init) -- the only inter-GPU traffic.
"""
from __future__ import annotations

from typing import Tuple


def shard_items(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) ITEM range of `rank`: contiguous, as even as possible, never splitting an item."""
    base, extra = divmod(n_items, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def shard_rows(n_rows: int, item_rows: int, rank: int, world: int, n_items: int = 0) -> Tuple[int, int]:
    """[first, last) ROW range of `rank` for a data set of whole `item_rows`-row items."""
    items = n_items or n_rows // item_rows
    a, b = shard_items(items, rank, world)
    return a * item_rows, b * item_rows


def node_slices(n_nodes: int, world: int):
    """Node-split mode (PF_OPT_NODE_SPLIT): rank r scores nodes [r*per, min((r+1)*per, B)), per = ceil(B/world);
    trailing ranks may get none.  Mirrors eval_host_node_split() in csrc/pf_engine.cu."""
    per = -(-n_nodes // world)
    return [(min(r * per, n_nodes), min((r + 1) * per, n_nodes)) for r in range(world)]


def init_engine_comm(engine, group=None) -> None:
    """Create the engine's NCCL communicator over the ranks of `group` (default: the world).
    The 128-byte NCCL id is made on rank 0 and broadcast through torch.distributed (any backend)."""
    import torch.distributed as dist
    from .engine import nccl_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ids = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    engine.comm_init(ids[0], rank, world)


def init_engine_peers(engine, group=None) -> None:
    """Attach the in-kernel peer exchange (pf_engine_peer_export / _attach): every rank's 64-byte CUDA IPC
    handle is all-gathered through torch.distributed (any backend), then each engine maps its peers' buffers.
    Afterwards the frontier kernels reduce the per-node partials themselves over NVLink peer memory."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    handles = [None] * world
    dist.all_gather_object(handles, engine.peer_export(), group=group)
    engine.peer_attach(handles, rank, world)


def combine_partials_host(sum_local, sum_sq_local, group=None):
    """Host-side equivalent of the engine's all-reduce (used by the gloo CPU tests of the sharding
    logic): partial (sum, sum_sq) per node are additive across item-aligned shards."""
    import torch
    import torch.distributed as dist
    t = torch.stack([torch.as_tensor(sum_local, dtype=torch.float64), torch.as_tensor(sum_sq_local, dtype=torch.float64)])
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0].numpy(), t[1].numpy()
