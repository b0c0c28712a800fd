"""Target-graph container (CSR) and the seeded synthetic stand-ins for BASELINE.json's configs.

The reference keeps targets as PyG ``Data`` objects and converts them to networkx before the
partition (``subgraph_counting/workload.py:223-231``).  Here a whole dataset of targets is one
block-diagonal int32 CSR that lives in HBM:

    rowptr[N+1], col[M]      sorted adjacency, symmetric, simple (no self loops, no duplicates)
    graph_ptr[B+1]           node range of every target graph (graph b owns [graph_ptr[b], graph_ptr[b+1]))

Node ids are dataset-global; ``global = graph_ptr[b] + local``.  Order inside a graph is preserved, so the
reference's ``n <= start_node`` test (``data.py:385``) is the same on global ids.

Generators follow SURVEY.md §8(d): no datasets are available offline, so every config is a seeded
synthetic graph set of the named shape, randomly relabelled (the partition depends on labels).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Sequence, Tuple

import numpy as np


@dataclass
class TargetCSR:
    rowptr: np.ndarray  # int32 [N+1] (int64 when M >= 2**31)
    col: np.ndarray  # int32 [M]
    graph_ptr: np.ndarray  # int32 [B+1]

    @property
    def num_nodes(self) -> int:
        return int(self.rowptr.shape[0] - 1)

    @property
    def num_directed_edges(self) -> int:
        return int(self.col.shape[0])

    @property
    def num_graphs(self) -> int:
        return int(self.graph_ptr.shape[0] - 1)

    def graph_of(self, nodes: np.ndarray) -> np.ndarray:
        return (np.searchsorted(self.graph_ptr, nodes, side="right") - 1).astype(np.int64)

    def to_networkx(self, gid: int):
        """networkx view of target ``gid`` with LOCAL node ids 0..n-1 (what the reference's partition sees)."""
        import networkx as nx

        lo, hi = int(self.graph_ptr[gid]), int(self.graph_ptr[gid + 1])
        g = nx.Graph()
        g.add_nodes_from(range(hi - lo))
        for u in range(lo, hi):
            for v in self.col[self.rowptr[u] : self.rowptr[u + 1]]:
                if v > u:
                    g.add_edge(u - lo, int(v) - lo)
        return g

    def edge_index(self) -> np.ndarray:
        """[2, M] (src,dst) directed edge list, row-major sorted."""
        deg = np.diff(self.rowptr)
        src = np.repeat(np.arange(self.num_nodes, dtype=np.int64), deg)
        return np.stack([src, self.col.astype(np.int64)])


def csr_from_edges(n: int, edges: np.ndarray, graph_ptr: Sequence[int] | None = None) -> TargetCSR:
    """Build the symmetric, deduplicated, self-loop-free sorted CSR from an undirected edge array [m,2]
    (mirrors ``T.ToUndirected`` + ``to_networkx(to_undirected=True)``: ``main.py:80``, ``workload.py:224``)."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    u = np.concatenate([edges[:, 0], edges[:, 1]])
    v = np.concatenate([edges[:, 1], edges[:, 0]])
    keep = u != v
    u, v = u[keep], v[keep]
    key = np.unique(u * n + v)
    u = key // n
    v = key % n
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(np.bincount(u, minlength=n))
    if graph_ptr is None:
        graph_ptr = [0, n]
    idx_t = np.int32 if rowptr[-1] < 2**31 else np.int64
    return TargetCSR(rowptr.astype(idx_t), v.astype(np.int32), np.asarray(graph_ptr, dtype=np.int32))


def csr_from_graph_list(graphs: Iterable[Tuple[int, np.ndarray]]) -> TargetCSR:
    """graphs: iterable of (n, edges[m,2] local ids) -> one block-diagonal CSR."""
    ptr = [0]
    all_e = []
    for n, e in graphs:
        e = np.asarray(e, dtype=np.int64).reshape(-1, 2)
        all_e.append(e + ptr[-1])
        ptr.append(ptr[-1] + int(n))
    e = np.concatenate(all_e) if all_e else np.zeros((0, 2), dtype=np.int64)
    return csr_from_edges(ptr[-1], e, ptr)


def csr_from_networkx(graphs) -> TargetCSR:
    """List of nx graphs whose nodes are 0..n-1 -> CSR (the ingest side of ``workload.py:223-231``)."""
    out = []
    for g in graphs:
        n = g.number_of_nodes()
        e = np.array([(a, b) for a, b in g.edges()], dtype=np.int64).reshape(-1, 2)
        out.append((n, e))
    return csr_from_graph_list(out)


def load_tu_dataset(root: str, name: str) -> TargetCSR:
    """Raw TU-format dataset on disk -> one block-diagonal CSR (the ingest of ``data.py:91-232``, which goes through
    ``torch_geometric.datasets.TUDataset`` + ``T.ToUndirected``, ``main.py:80``): ``<root>/<name>/raw/<name>_A.txt`` (or
    ``<root>/<name>_A.txt``) holds one ``u, v`` pair of 1-based dataset-wide node ids per line,
    ``<name>_graph_indicator.txt`` the 1-based graph id of every node.  Nodes of a graph are contiguous and graphs are
    numbered in file order, so node order inside a graph - what the canonical partition depends on - is the file's."""
    import os

    for d in (os.path.join(root, name, "raw"), os.path.join(root, name), root):
        if os.path.exists(os.path.join(d, f"{name}_A.txt")):
            break
    else:
        raise FileNotFoundError(f"{name}_A.txt not found under {root}")
    gi = np.loadtxt(os.path.join(d, f"{name}_graph_indicator.txt"), dtype=np.int64, ndmin=1)
    if gi.size == 0 or np.any(np.diff(gi) < 0) or gi[0] != 1:
        raise ValueError("graph indicator must be 1-based and non-decreasing")
    n = gi.shape[0]
    edges = np.loadtxt(os.path.join(d, f"{name}_A.txt"), dtype=np.int64, delimiter=",", ndmin=2) - 1
    if edges.size and (edges.min() < 0 or edges.max() >= n):
        raise ValueError("edge endpoint outside the node range")
    if edges.size and np.any(gi[edges[:, 0]] != gi[edges[:, 1]]):
        raise ValueError("edge between two different graphs")
    graph_ptr = np.concatenate([[0], np.cumsum(np.bincount(gi - 1, minlength=int(gi[-1])))])
    return csr_from_edges(n, edges, graph_ptr)


# --------------------------------------------------------------------------------------------
# seeded synthetic graph families
# --------------------------------------------------------------------------------------------


def _random_connected(n: int, m: int, rng: np.random.Generator) -> np.ndarray:
    """Random spanning tree + extra uniform edges until m distinct undirected edges; then random relabel."""
    n = int(n)
    m = int(min(max(m, n - 1), n * (n - 1) // 2))
    if n == 1:
        return np.zeros((0, 2), dtype=np.int64)
    parent = (rng.random(n - 1) * np.arange(1, n)).astype(np.int64)  # node i attaches to a uniform j<i
    es = set((int(parent[i - 1]), i) for i in range(1, n))
    while len(es) < m:
        k = m - len(es)
        a = rng.integers(0, n, size=2 * k + 4)
        b = rng.integers(0, n, size=2 * k + 4)
        for x, y in zip(a, b):
            if x == y:
                continue
            es.add((int(min(x, y)), int(max(x, y))))
            if len(es) >= m:
                break
    e = np.array(sorted(es), dtype=np.int64)
    perm = rng.permutation(n)
    return perm[e]


def gen_mutag_shaped(seed: int = 0, num_graphs: int = 188) -> TargetCSR:
    """Config 1: 188 molecule-like graphs, n~N(17.93,4.59) in [10,28], m~1.104 n."""
    rng = np.random.default_rng(seed)
    gl = []
    for _ in range(num_graphs):
        n = int(np.clip(round(rng.normal(17.93, 4.59)), 10, 28))
        gl.append((n, _random_connected(n, round(1.104 * n), rng)))
    return csr_from_graph_list(gl)


def gen_cox2_shaped(seed: int = 0, num_graphs: int = 467) -> TargetCSR:
    rng = np.random.default_rng(seed + 101)
    gl = []
    for _ in range(num_graphs):
        n = int(max(8, round(rng.normal(41.2, 4.0))))
        gl.append((n, _random_connected(n, round(1.055 * n), rng)))
    return csr_from_graph_list(gl)


def gen_enzymes_shaped(seed: int = 0, num_graphs: int = 600) -> TargetCSR:
    """Config 2 (the bench workload): ENZYMES-like pool, n~N(32.6,15) >= 4, m~min(1.9 n, n(n-1)/2)."""
    rng = np.random.default_rng(seed + 202)
    gl = []
    for _ in range(num_graphs):
        n = int(max(4, round(rng.normal(32.6, 15.0))))
        gl.append((n, _random_connected(n, round(min(1.9 * n, n * (n - 1) / 2)), rng)))
    return csr_from_graph_list(gl)


def gen_imdb_shaped(seed: int = 0, num_graphs: int = 1000) -> TargetCSR:
    """Config 4: dense ego graphs. Ego joined to all, remainder = union of random cliques until m~4.88 n."""
    rng = np.random.default_rng(seed + 404)
    gl = []
    for _ in range(num_graphs):
        n = int(np.clip(round(rng.normal(19.8, 10.0)), 12, 136))
        es = set((0, i) for i in range(1, n))
        target = round(4.88 * n)
        guard = 0
        while len(es) < target and guard < 64:
            guard += 1
            k = int(rng.integers(3, max(4, min(n - 1, 12))))
            members = rng.choice(np.arange(1, n), size=min(k, n - 1), replace=False)
            for i in range(len(members)):
                for j in range(i + 1, len(members)):
                    a, b = int(members[i]), int(members[j])
                    es.add((min(a, b), max(a, b)))
        e = np.array(sorted(es), dtype=np.int64)
        perm = rng.permutation(n)
        gl.append((n, perm[e]))
    return csr_from_graph_list(gl)


def gen_syn1827_shaped(seed: int = 0, stride: int = 1, max_graphs: int | None = None) -> TargetCSR:
    """Config 3: size/density schedule of ``subgraph_counting/syn_data.py:684-729`` (1827 graph ids);
    generator family replaced by G(n,m)+spanning tree.  ``stride`` keeps every stride-th id."""
    rng = np.random.default_rng(seed + 303)
    gl = []
    ids = list(range(0, 1827, stride))
    if max_graphs is not None:
        ids = ids[:max_graphs]
    for gid in ids:
        if gid < 1380:
            n = gid // 23 + 10
            deg = 0.5 * (gid % 23) + 1 + rng.triangular(-0.5, 0.0, 0.5)
        else:
            n = int(5 * ((gid - 1380) // 3) + 60 + round(rng.triangular(-5, 0, 5)))
            deg = 1 + 2 * rng.random()
        n = max(n, 4)
        m = int(np.clip(rng.normal(1.0, 0.1) * np.floor(n * deg), n - 1, n * (n - 1) // 2))
        gl.append((n, _random_connected(n, m, rng)))
    return csr_from_graph_list(gl)


def gen_powerlaw(n: int, m_undirected: int, seed: int = 0, gamma: float = 2.5, max_deg_frac: float = 0.002):
    """Config 5: Chung-Lu power-law graph (expected degrees w_i ~ i^{-1/(gamma-1)}, capped), connected by a
    random spanning tree on a random permutation, then randomly relabelled.  Vectorised; n=1e7/m=1e8 ~ minutes."""
    rng = np.random.default_rng(seed + 505)
    w = (np.arange(1, n + 1, dtype=np.float64)) ** (-1.0 / (gamma - 1.0))
    w = w * (2.0 * m_undirected / w.sum())  # expected degrees
    w = np.minimum(w, max(max_deg_frac * n, 8.0))  # cap the hubs
    p = w / w.sum()
    cdf = np.cumsum(p)
    k = int(m_undirected)
    a = np.searchsorted(cdf, rng.random(k)).clip(0, n - 1)
    b = np.searchsorted(cdf, rng.random(k)).clip(0, n - 1)
    tree_child = np.arange(1, n, dtype=np.int64)
    tree_parent = (rng.random(n - 1) * tree_child).astype(np.int64)
    e = np.stack([np.concatenate([a, tree_child]), np.concatenate([b, tree_parent])], axis=1)
    perm = rng.permutation(n)
    return csr_from_edges(n, perm[e])


def first_nonempty_centres(csr: TargetCSR, count: int) -> np.ndarray:
    """Cheap sufficient test used to pick bench centres before any partition is run: a node whose smallest
    neighbour is below it always has >=1 induced edge in its canonical neighborhood (the edge to that
    neighbour survives the ``<=`` filter and touches the centre), and a node with no smaller neighbour has
    an edge-free neighborhood (its component in G[<=c] is itself).  So this is exact, not a heuristic."""
    deg = np.diff(csr.rowptr)
    first = np.full(csr.num_nodes, np.iinfo(np.int64).max, dtype=np.int64)
    nz = deg > 0
    first[nz] = csr.col[csr.rowptr[:-1][nz]]
    ok = np.nonzero(first < np.arange(csr.num_nodes))[0]
    return ok[:count].astype(np.int32)
