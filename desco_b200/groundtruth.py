"""Ground-truth canonical counts on the GPU - host side of csrc/groundtruth.cu.

Mirrors the label generation of ``subgraph_counting/workload.py`` (reference @ 4508f7a): ``MatchSubgraphWorker`` :327-348
(VF2 induced-subgraph isomorphisms, every mapping credited to ``max(vmap.keys())``), ``Workload.compute_groundtruth``
:551-726 (per node and query: mappings / ``SymmetricFactor``) and ``data.SymmetricFactor`` / ``GenVMap`` :61-88.
``mappings / |Aut(query)|`` is the number of node SETS whose induced subgraph is the query and whose largest node is the
canonical node - which the CUDA kernel counts directly (ESU enumeration + pattern table), for connected queries of 3 to 5
nodes (the 29 standard queries, ``data.py:37-58``).  There is no CPU fallback.
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .data import DeviceCSR, _ptr, _require_cuda, _stream

_PAIRS = {k: [(i, j) for j in range(k) for i in range(j)] for k in (3, 4, 5)}  # bit b of a pattern <-> node pair; (0,1) first


def SymmetricFactor(graph, node_feat_key: Optional[str] = None) -> int:
    """``data.py:61-68``: the number of automorphisms of ``graph`` (self-mappings VF2 would enumerate)."""
    if node_feat_key is not None:
        raise NotImplementedError("node-feature-aware matching (--use_node_feature) is not built")
    nodes = list(graph.nodes)
    n = len(nodes)
    adj = np.zeros((n, n), dtype=bool)
    pos = {u: i for i, u in enumerate(nodes)}
    for a, b in graph.edges:
        adj[pos[a], pos[b]] = adj[pos[b], pos[a]] = True
    if n > 8:
        import networkx as nx

        return sum(1 for _ in nx.algorithms.isomorphism.GraphMatcher(graph, graph).isomorphisms_iter())
    return sum(1 for p in itertools.permutations(range(n)) if np.array_equal(adj[np.ix_(p, p)], adj))


def _pattern_of(adj: np.ndarray, order: Sequence[int]) -> int:
    k = len(order)
    return sum(1 << b for b, (i, j) in enumerate(_PAIRS[k]) if adj[order[i], order[j]])


def pattern_tables(queries: List) -> List[np.ndarray]:
    """For k = 3, 4, 5: uint8[2^(k(k-1)/2)] mapping the adjacency bits of a node k-tuple (bit of pair (i, j), i < j, at
    position j(j-1)/2 + i) to the column of the query it is isomorphic to, 255 for none.  Built by brute force over the
    node orders of every query - isomorphism-complete for <= 5 nodes."""
    luts = [np.full(1 << len(_PAIRS[k]), 255, dtype=np.uint8) for k in (3, 4, 5)]
    for qi, q in enumerate(queries):
        k = q.number_of_nodes()
        if k not in (3, 4, 5):
            raise NotImplementedError("ground-truth kernel: queries of 3 to 5 nodes")
        import networkx as nx

        if not nx.is_connected(q):
            raise NotImplementedError("ground-truth kernel: connected queries only (the ESU enumeration visits connected sets)")
        nodes = list(q.nodes)
        pos = {u: i for i, u in enumerate(nodes)}
        adj = np.zeros((k, k), dtype=bool)
        for a, b in q.edges:
            if a != b:
                adj[pos[a], pos[b]] = adj[pos[b], pos[a]] = True
        lut = luts[k - 3]
        for perm in itertools.permutations(range(k)):
            p = _pattern_of(adj, perm)
            if lut[p] not in (255, qi):
                raise ValueError(f"queries {int(lut[p])} and {qi} are isomorphic")
            lut[p] = qi
    return luts


def canonical_count_truth(graph: DeviceCSR, query_ids: Optional[List[int]] = None, queries: Optional[List] = None) -> torch.Tensor:
    """truth[node, query] (float32, dataset-global node order) = ``compute_groundtruth``'s ``count_motif``
    (``workload.py:551-699``): induced occurrences of every query whose largest node is the row's node."""
    import networkx as nx

    if (query_ids is None) == (queries is None):
        raise ValueError("query_ids or queries must be given (and not both)")  # workload.py:557-560
    if queries is None:
        queries = [nx.graph_atlas(i) for i in query_ids]
    lib = _lib.load()
    dev = _require_cuda(graph.rowptr.device)
    luts = [torch.from_numpy(t).to(dev) for t in pattern_tables(queries)]
    N, Q = graph.num_nodes, len(queries)
    out = torch.zeros((N, Q), dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    max_k = max(q.number_of_nodes() for q in queries)
    with torch.cuda.device(dev):
        _lib.check(lib.desco_groundtruth_count(_ptr(graph.rowptr), _ptr(graph.col), _ptr(graph.graph_ptr), graph.num_graphs,
                                               graph.max_graph_nodes, _ptr(luts[0]), _ptr(luts[1]), _ptr(luts[2]), Q, max_k,
                                               _ptr(out), _ptr(status), _stream()), "desco_groundtruth_count")
    code = int(status.item())
    if code:
        _lib.check(code, "desco_groundtruth_count (device status)")
    return out.to(torch.float32)
