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


class GossipShardPlan:
    """Row / query-group geometry of the sharded gossip forward: every rank owns ``n_loc`` consecutive node rows (a
    multiple of the 128-row tile, so the in-place all-gathers have equal block sizes), the buffers hold
    ``n_rows = world * n_loc >= N`` rows, the queries are cut into groups of ``query_group``."""

    def __init__(self, num_nodes: int, num_queries: int, world: int, query_group: int = 4, row_align: int = 128):
        per = -(-max(num_nodes, 1) // world)
        self.n_loc = -(-per // row_align) * row_align
        self.n_rows = self.n_loc * world
        self.world, self.num_nodes, self.num_queries = world, num_nodes, num_queries
        self.query_group = max(1, min(int(query_group), max(num_queries, 1)))
        self.ranges = [(min(r * self.n_loc, num_nodes), min((r + 1) * self.n_loc, num_nodes)) for r in range(world)]
        self.groups = [(q0, min(q0 + self.query_group, num_queries)) for q0 in range(0, num_queries, self.query_group)]

    def halo_bytes(self) -> int:
        """Bytes every rank receives in the halo all-gathers of one forward (s4: 16 B per node and query)."""
        return 16 * self.num_queries * (self.n_rows - self.n_loc)

    def output_bytes(self) -> int:
        return 4 * self.num_queries * (self.n_rows - self.n_loc)


def gossip_shard_plan(num_nodes: int, num_queries: int, world: int, query_group: int = 4) -> GossipShardPlan:
    return GossipShardPlan(num_nodes, num_queries, world, query_group)


class _Done:
    def wait(self):
        return True


class ProcessGroupComm:
    """The exchange step of the sharded forward over ``torch.distributed`` (NCCL on the GPUs of one box; gloo in the CPU
    tests): an IN-PLACE all-gather - rank r's rows ``[r * n_loc, (r + 1) * n_loc)`` of ``full`` are the send buffer -
    issued asynchronously; ``wait()`` of the returned handle orders the current CUDA stream after it."""

    def __init__(self, group=None):
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1

    def all_gather_block(self, full: torch.Tensor, n_loc: int, tag=None):
        if self.world == 1:
            return _Done()
        flat = full.view(-1)
        per = flat.numel() // self.world
        return dist.all_gather_into_tensor(flat, flat[self.rank * per:(self.rank + 1) * per], group=self.group, async_op=True)


class LocalComm:
    """``world`` emulated ranks inside ONE process on one device (tests): every rank registers its buffer at
    ``all_gather_block``; ``wait()`` copies the other ranks' blocks in, so every rank must have issued a gather before any
    rank waits for it (``GossipShardedRun``: all ``start()``, then all ``finish()``, then all ``result()``) - the
    ordering a real collective imposes."""

    def __init__(self, world: int):
        self.world = world
        self.bufs = {}

    def for_rank(self, rank: int):
        parent = self

        class _Rank:
            world = parent.world

            def __init__(self):
                self.rank = rank

            def all_gather_block(self, full, n_loc, tag=None):
                parent.bufs[(tag, rank)] = full

                class _W:
                    def wait(_self):
                        for r in range(parent.world):
                            if r != rank:
                                full[r * n_loc:(r + 1) * n_loc].copy_(parent.bufs[(tag, r)][r * n_loc:(r + 1) * n_loc])
                        return True

                return _W()

        return _Rank()


def centre_work_estimate(graph, depth: int) -> np.ndarray:
    """Per-centre proxy of the partition + counting work, for cutting balanced contiguous centre ranges.  The k-hop BFS is
    unrestricted (``data.py:329-338``), so its cost follows the ball: 1 + deg(c) at depth 1, plus the degrees of the
    neighbours from depth 2 on; what is KEPT (nodes <= centre, ``data.py:385``) and emitted grows with the centre's
    position in the id order, hence the (0.25 + 0.75 c/N) factor."""
    rowptr = graph.rowptr.to(torch.int64)
    n = rowptr.numel() - 1
    deg = rowptr[1:] - rowptr[:-1]
    w = 1 + deg
    if depth >= 2 and graph.col.numel():
        c = torch.cumsum(deg[graph.col.long()], 0)
        c = torch.cat([c.new_zeros(1), c])
        w = w + (c[rowptr[1:]] - c[rowptr[:-1]])
    pos = torch.arange(n, device=rowptr.device, dtype=torch.float64) / max(n, 1)
    if graph.num_graphs > 1:  # the <= filter acts inside a graph: position of the node within its own graph
        gp = graph.graph_ptr.to(torch.int64)
        gid = torch.searchsorted(gp, torch.arange(n, device=gp.device), right=True) - 1
        size = (gp[1:] - gp[:-1]).clamp(min=1)[gid]
        pos = (torch.arange(n, device=gp.device) - gp[gid]).to(torch.float64) / size
    return (w.to(torch.float64) * (0.25 + 0.75 * pos)).cpu().numpy()


class ShardedPipeline:
    """Rank-local driver: ``count_neighborhoods`` (no collective) -> ``gather_node_counts`` (all-gather) ->
    ``gossip`` (layer 0 local, all-gather of the halo scalars, layer 1 local, all-gather of the result rows)."""

    def __init__(self, graph, neighborhood_model, gossip_model=None, group=None, depth: int = 4):
        self.graph, self.nm, self.gm, self.group, self.depth = graph, neighborhood_model, gossip_model, group, depth
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.work_estimate = centre_work_estimate(graph, depth)
        self.centre_shards = balanced_shards(self.work_estimate, self.world)
        self.comm = ProcessGroupComm(group)

    def deal_centres(self, centres: np.ndarray) -> np.ndarray:
        """This rank's share of an explicit centre list, dealt by estimated work (longest-processing-time first: the
        heaviest remaining centre goes to the least loaded rank) instead of by contiguous range - the depth-2 ball of a
        hub centre is 10^3 times an ordinary one, which no contiguous cut of a SAMPLE balances.  Deterministic and
        identical on every rank; results stay rank-local either way (no collective)."""
        centres = np.asarray(centres, dtype=np.int64)
        w = self.work_estimate[centres]
        order = np.argsort(-w, kind="stable")
        load = np.zeros(self.world)
        owner = np.empty(len(centres), dtype=np.int64)
        for i in order:
            r = int(np.argmin(load))
            owner[i] = r
            load[r] += w[i]
        return np.sort(centres[owner == self.rank])

    def count_neighborhoods(self, centres: Optional[torch.Tensor] = None, max_centres: Optional[int] = None):
        """Canonical partition + SHMP counting of this rank's centres (default: its whole balanced range; or the given
        rank-local list), in int32-safe chunks (``data.partition_batches``).  No collective.
        Returns (centres [G_r], counts [G_r, Q])."""
        from .data import partition_batches

        if centres is None:
            lo, hi = self.centre_shards[self.rank]
            centres = torch.arange(lo, hi, dtype=torch.int32, device=self.graph.rowptr.device)
        kept, counts = [], []
        self.last_stats = {"neighborhoods": 0, "rows": 0, "directed_edges": 0, "max_rows": 0, "chunks": 0}
        with torch.no_grad():
            for batch in partition_batches(self.graph, centres, self.depth, "hetero", max_centres):
                kept.append(batch.centre)
                counts.append(self.nm.graph_to_count(batch))
                st = self.last_stats
                st["neighborhoods"] += batch.num_neighborhoods
                st["rows"] += batch.num_rows
                st["directed_edges"] += batch.num_edges
                st["max_rows"] = max(st["max_rows"], batch.max_rows if batch.num_neighborhoods else 0)
                st["chunks"] += 1
        if not kept:
            dev = self.graph.rowptr.device
            return torch.empty(0, dtype=torch.int32, device=dev), torch.empty((0, 0), dtype=torch.float32, device=dev)
        return torch.cat(kept), torch.cat(counts)

    def gather_node_counts(self, centres: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
        """x[N, Q] replicated on every rank (``GossipDataset.apply_neighborhood_count``, workload.py:107-112)."""
        lo, hi = self.centre_shards[self.rank]
        Q = counts.shape[1]
        x_local = torch.zeros((hi - lo, Q), dtype=torch.float32, device=counts.device)
        x_local[(centres.long() - lo)] = counts
        sizes = [b - a for a, b in self.centre_shards]
        return all_gather_rows(x_local, self.group, sizes) if self.world > 1 else x_local

    def gossip(self, x: torch.Tensor, query_emb: torch.Tensor, query_group: int = 4, gather_output: bool = True) -> torch.Tensor:
        """out[N, Q] replicated on every rank (or this rank's rows with ``gather_output=False``): the pipelined
        node-range-sharded forward of ``gnn_model.GossipShardedRun``."""
        return self.gm.emb_model.forward_sharded(self.graph.rowptr, self.graph.col, x, query_emb, self.comm, query_group,
                                                 gather_output)
