"""Oracle helpers for targets too large to hand to networkx whole (config 5).  TEST INFRASTRUCTURE ONLY.

The reference's partition (``subgraph_counting/data.py:329-396``) and gossip (``gnn_model.py:231-260,280-359``) are
local computations: a depth-k canonical neighborhood only reads the k-hop ball of its centre, and the two-layer gossip
output of a node only reads its 2-hop closure.  So a seeded SAMPLE of centres / nodes of a 1M- or 10M-node CSR can be
checked against the unmodified oracle (or the reference's own functions) on the induced sub-structure:

* ``BallView``  - duck-types the ``TargetCSR`` interface ``oracle.partition.partition_dataset`` uses, but
  ``to_networkx`` returns the union of the k-hop balls of the sample (global node ids as labels) - every walk of
  length <= k from a centre stays inside its ball, so ``k_neigh`` / ``get_neigh_*`` return what they would on the
  full graph;
* ``gossip_closure`` - order-preserving relabel of the 2-hop closure of a node sample with every edge incident to the
  1-hop ring, so ``j < i`` comparisons (``gnn_model.py:248``) and the neighbour sums of the ring are those of the full
  graph; only the sample's output rows are meaningful.
"""
from __future__ import annotations

from typing import Tuple

import networkx as nx
import numpy as np


def _neighbours(rowptr: np.ndarray, col: np.ndarray, nodes: np.ndarray) -> np.ndarray:
    lo, hi = rowptr[nodes].astype(np.int64), rowptr[nodes + 1].astype(np.int64)
    n = hi - lo
    if n.sum() == 0:
        return np.zeros(0, dtype=np.int64)
    idx = np.repeat(lo - np.concatenate([[0], np.cumsum(n)[:-1]]), n) + np.arange(int(n.sum()))
    return col[idx].astype(np.int64)


def khop_ball(rowptr: np.ndarray, col: np.ndarray, centre: int, k: int) -> np.ndarray:
    """Sorted node ids within <= k hops of ``centre`` (plain BFS on the CSR)."""
    seen = np.array([centre], dtype=np.int64)
    frontier = seen
    for _ in range(k):
        nb = np.unique(_neighbours(rowptr, col, frontier))
        frontier = np.setdiff1d(nb, seen, assume_unique=True)
        if frontier.size == 0:
            break
        seen = np.union1d(seen, frontier)
    return seen


def ball_sizes(rowptr: np.ndarray, col: np.ndarray, centres, k: int) -> np.ndarray:
    return np.array([khop_ball(rowptr, col, int(c), k).size for c in centres], dtype=np.int64)


class BallView:
    """``TargetCSR`` stand-in over the union of the k-hop balls of ``centres`` (one graph, id 0)."""

    def __init__(self, rowptr: np.ndarray, col: np.ndarray, centres, depth: int):
        self.rowptr, self.col = np.asarray(rowptr), np.asarray(col)
        self.num_nodes = len(self.rowptr) - 1
        self.graph_ptr = np.array([0, self.num_nodes], dtype=np.int64)
        nodes = np.zeros(0, dtype=np.int64)
        for c in centres:
            nodes = np.union1d(nodes, khop_ball(self.rowptr, self.col, int(c), depth))
        self.nodes = nodes
        g = nx.Graph()
        g.add_nodes_from(int(u) for u in nodes)
        member = np.zeros(self.num_nodes, dtype=bool)
        member[nodes] = True
        for u in nodes:
            nb = self.col[self.rowptr[u]:self.rowptr[u + 1]]
            nb = nb[member[nb] & (nb > u)]
            g.add_edges_from((int(u), int(v)) for v in nb)
        self._g = g

    def graph_of(self, centres) -> np.ndarray:
        return np.zeros(len(centres), dtype=np.int64)

    def to_networkx(self, gid: int) -> nx.Graph:
        assert gid == 0
        return self._g


def gossip_closure(rowptr: np.ndarray, col: np.ndarray, sample) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(nodes, edge_index, sample_pos): ``nodes`` = sorted global ids of the 2-hop closure of ``sample``; ``edge_index``
    [2, e] over closure-local ids = every directed edge with at least one endpoint in sample + its neighbours (both
    directions); ``sample_pos`` = local ids of the sample."""
    rowptr, col = np.asarray(rowptr), np.asarray(col)
    s = np.unique(np.asarray(sample, dtype=np.int64))
    ring = np.union1d(s, _neighbours(rowptr, col, s))
    nodes = np.union1d(ring, _neighbours(rowptr, col, ring))
    lo, hi = rowptr[ring].astype(np.int64), rowptr[ring + 1].astype(np.int64)
    src = np.repeat(ring, hi - lo)
    dst = _neighbours(rowptr, col, ring)
    in_ring = np.zeros(len(rowptr) - 1, dtype=bool)
    in_ring[ring] = True
    # ring -> anything (all adjacency of ring nodes) plus the reverse of the edges that leave the ring
    out = ~in_ring[dst]
    u = np.concatenate([src, dst[out]])
    v = np.concatenate([dst, src[out]])
    ei = np.stack([np.searchsorted(nodes, u), np.searchsorted(nodes, v)])
    return nodes, ei, np.searchsorted(nodes, np.asarray(sample, dtype=np.int64))
