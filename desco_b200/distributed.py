"""One-node multi-GPU sharding of the hot path (one process per GPU, ``torch.distributed`` over NCCL / NVLink).

The reference runs inference on a single GPU (``main.py:208``) and raises ``NotImplementedError`` for multi-GPU gossip
(``main.py:353-356``); this module is new:

* canonical partition + SHMP counting are independent per centre: rank r owns a contiguous range of centres chosen so
  that the estimated work (sum of degrees) is balanced; NO collective on the data path, results stay sharded;
* the hand-off to gossip is one all-gather of the per-node counts ``x[N, Q]`` (rows of absent neighborhoods are zero);
* gossip is sharded by node range: layer 0 needs the neighbours' counts (already replicated by the all-gather), layer 1
  needs the neighbours' four layer-0 scalars ``s4[N, Q, 4]`` - the halo.  With random labels on a power-law graph almost
  every node is a halo node of some rank, so the halo exchange is a dense all-gather of ``s4``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous near-equal split of range(n)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_shards(weights: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges of ``len(weights)`` items with near-equal total weight (prefix-sum cut points).  For the
    partition, weight = 1 + degree of the centre is a cheap proxy of the neighborhood work."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if n == 0:
        return [(0, 0)] * world
    cum = np.cumsum(w)
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, cum[-1] * r / world, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.asarray(cuts))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def all_gather_rows(local: torch.Tensor, group=None, sizes: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Concatenate per-rank row blocks of different lengths along dim 0 (pad to the longest, all_gather, trim)."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    if sizes is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_n = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(all_n, n, group=group)
        sizes = [int(x.item()) for x in all_n]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def node_ranges(num_nodes: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(num_nodes, r, world) for r in range(world)]


class ShardedPipeline:
    """Rank-local driver: ``count_neighborhoods`` (no collective) -> ``gather_node_counts`` (all-gather) ->
    ``gossip`` (layer 0 local, all-gather of the halo scalars, layer 1 local, all-gather of the result rows)."""

    def __init__(self, graph, neighborhood_model, gossip_model=None, group=None, depth: int = 4):
        self.graph, self.nm, self.gm, self.group, self.depth = graph, neighborhood_model, gossip_model, group, depth
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if graph.host is not None:
            deg = np.diff(graph.host.rowptr)
        else:
            deg = (graph.rowptr[1:] - graph.rowptr[:-1]).cpu().numpy()
        self.centre_shards = balanced_shards(1.0 + deg, self.world)
        self.node_shards = node_ranges(graph.num_nodes, self.world)

    def count_neighborhoods(self):
        """Canonical partition + SHMP counting of this rank's centres.  Returns (centres [G_r], counts [G_r, Q])."""
        from .data import partition_batch

        lo, hi = self.centre_shards[self.rank]
        centres = torch.arange(lo, hi, dtype=torch.int32, device=self.graph.rowptr.device)
        batch = partition_batch(self.graph, centres, self.depth, "hetero")
        with torch.no_grad():
            counts = self.nm.graph_to_count(batch)
        return batch.centre, counts

    def gather_node_counts(self, centres: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
        """x[N, Q] replicated on every rank (``GossipDataset.apply_neighborhood_count``, workload.py:107-112)."""
        lo, hi = self.centre_shards[self.rank]
        Q = counts.shape[1]
        x_local = torch.zeros((hi - lo, Q), dtype=torch.float32, device=counts.device)
        x_local[(centres.long() - lo)] = counts
        sizes = [b - a for a, b in self.centre_shards]
        return all_gather_rows(x_local, self.group, sizes) if self.world > 1 else x_local

    def gossip(self, x: torch.Tensor, query_emb: torch.Tensor) -> torch.Tensor:
        """out[N, Q] replicated on every rank."""
        lo, hi = self.node_shards[self.rank]
        sizes = [b - a for a, b in self.node_shards]
        exchange = (lambda t: all_gather_rows(t, self.group, sizes)) if self.world > 1 else (lambda t: t)
        return self.gm.emb_model.forward_node_range(self.graph.rowptr, self.graph.col, x, query_emb, lo, hi, exchange)
