"""Oracle: canonical neighborhood partition (CPU, networkx + Python sets).  TEST INFRASTRUCTURE ONLY.

Restates ``subgraph_counting/data.py:329-396`` (k-hop set BFS, ``<= centre`` filter, component of the centre) and the
driver loop ``subgraph_counting/workload.py:243-260`` (indicator / index bookkeeping), then packs every kept
neighborhood into the flat layout the CUDA path emits so the two can be compared array-for-array.

Packed layout (all numpy):
    nbh_ptr[G+1]   node offsets of the kept neighborhoods
    node_gid[V]    dataset-global node id, ascending inside a neighborhood => the canonical node (the max,
                   ``workload.py:346``) is the LAST row of its neighborhood, every other row is a "count" node
    edge_ptr[V+1]  CSR over batch rows (directed edges; both directions present, ``transforms.py:331``)
    edge_col[E]    batch-global row index of the other endpoint, ascending inside a row
    edge_tri[E]    1 = "triangle" relation, 0 = "tride" (``transforms.py:221``)
    centre[G]      dataset-global id of the canonical node
    index[G,2]     (graph id, local node id)   == ``nx_neighs_index``      (``workload.py:259,293``)
    indicator[C]   kept / dropped per input centre == ``nx_neighs_indicator`` (``workload.py:256-258,294``)
"""
from __future__ import annotations

import ast
import os
from typing import Callable, Dict, Optional

import networkx as nx
import numpy as np

REFERENCE_DATA_PY = "/root/reference/subgraph_counting/data.py"


# ---------------------------------------------------------------------------------------------
# restatement of the four reference functions
# ---------------------------------------------------------------------------------------------


def k_neigh(G: nx.Graph, start_node, k: int):
    """All nodes within <= k hops of ``start_node``; UNRESTRICTED (``data.py:329-338``)."""
    seen = {start_node}
    frontier = {start_node}
    for _ in range(k):
        reached = set()
        for u in frontier:
            reached.update(G.neighbors(u))
        frontier = reached - seen
        seen |= frontier
    return list(seen)


def k_neigh_canonical(G: nx.Graph, start_node, k: int):
    """k-hop BFS that only walks through nodes ``<= start_node`` (``data.py:341-350``)."""
    seen = {start_node}
    frontier = {start_node}
    for _ in range(k):
        reached = set()
        for u in frontier:
            reached.update(v for v in G.neighbors(u) if v <= start_node)
        frontier = reached - seen
        seen |= frontier
    return list(seen)


def _component_of(G: nx.Graph, nodes, start_node) -> nx.Graph:
    sub = G.subgraph(nodes)
    for comp in nx.connected_components(sub):
        if start_node in comp:
            return sub.subgraph(comp).copy()
    raise AssertionError("centre not in its own candidate set")


def get_neigh_hetero(graph: nx.Graph, node, radius: int) -> nx.Graph:
    """Default partition (``data.py:375-396``): unrestricted k-hop set, keep ``<= centre`` (:385), keep the
    centre's component of the induced subgraph (:387-390), node attr type = count / canonical (:391-394)."""
    keep = [u for u in k_neigh(graph, node, radius) if u <= node]
    neigh = _component_of(graph, keep, node)
    for u in neigh.nodes:
        neigh.nodes[u]["type"] = "count"
    neigh.nodes[node]["type"] = "canonical"
    return neigh


def get_neigh_canonical(graph: nx.Graph, node, radius: int) -> nx.Graph:
    """Homogeneous variant (``data.py:353-372``), used when ``hetero_graph=False`` (``workload.py:241``)."""
    import torch

    neigh = _component_of(graph, k_neigh_canonical(graph, node, radius), node)
    for u in neigh.nodes:
        neigh.nodes[u]["node_feature"] = torch.zeros(1)
    neigh.nodes[node]["node_feature"] = torch.ones(1)
    return neigh


def load_reference_functions(path: str = REFERENCE_DATA_PY) -> Optional[Dict[str, Callable]]:
    """Execute the reference's OWN four partition functions from source (they are pure networkx; the rest of
    data.py needs torch_geometric / deepsnap / ogb, so only these FunctionDefs are compiled).  Returns None when
    the reference tree is not mounted (the GPU box)."""
    if not os.path.exists(path):
        return None
    import torch

    wanted = {"k_neigh", "k_neigh_canonical", "get_neigh_canonical", "get_neigh_hetero"}
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    mod = ast.Module(body=body, type_ignores=[])

    class _NoPyG:  # `type(graph) == pyg.data.data.Data` must simply be False for nx inputs
        class data:
            class data:
                class Data:
                    pass

    ns = {"nx": nx, "torch": torch, "pyg": _NoPyG, "pyg_utils": None}
    exec(compile(mod, path, "exec"), ns)
    return {k: ns[k] for k in wanted}


# ---------------------------------------------------------------------------------------------
# SHMP type of every edge of ONE neighborhood (set formulation; the sparse-matmul literal form is in shmp_types.py)
# ---------------------------------------------------------------------------------------------


def edge_is_triangle(neigh: nx.Graph, u, v) -> bool:
    """(A*A^2 + A)[u,v] > 1  <=>  u and v have a common neighbour inside the neighborhood (``transforms.py:201-221``)."""
    return any(True for _ in nx.common_neighbors(neigh, u, v))


# ---------------------------------------------------------------------------------------------
# dataset driver + packing
# ---------------------------------------------------------------------------------------------


def partition_dataset(csr, depth: int, mode: str = "hetero", centres: Optional[np.ndarray] = None,
                      funcs: Optional[Dict[str, Callable]] = None, with_types: bool = True) -> Dict[str, np.ndarray]:
    """Loop of ``NeighborhoodDataset.process`` (``workload.py:243-260``): every centre (default: every node of every
    graph, graph by graph, node ascending), drop edge-free neighborhoods, record indicator + index, pack.

    ``funcs``: pass ``load_reference_functions()`` to run the reference's own code instead of the restatement."""
    f_het = (funcs or {}).get("get_neigh_hetero", get_neigh_hetero)
    f_can = (funcs or {}).get("get_neigh_canonical", get_neigh_canonical)
    get_neigh = f_het if mode == "hetero" else f_can
    if centres is None:
        centres = np.arange(csr.num_nodes, dtype=np.int64)
    centres = np.asarray(centres, dtype=np.int64)
    gids = csr.graph_of(centres)
    cache: Dict[int, nx.Graph] = {}

    nbh_ptr, node_gid, edge_ptr, edge_col, edge_tri = [0], [], [0], [], []
    centre_out, index, indicator = [], [], []
    for c, gid in zip(centres, gids):
        gid = int(gid)
        if gid not in cache:
            if len(cache) > 64:
                cache.clear()
            cache[gid] = csr.to_networkx(gid)
        g = cache[gid]
        base = int(csr.graph_ptr[gid])
        local = int(c) - base
        neigh = get_neigh(g, local, depth)
        if neigh.number_of_edges() == 0:  # workload.py:253-256
            indicator.append(False)
            continue
        indicator.append(True)
        index.append((gid, local))
        centre_out.append(int(c))
        nodes = sorted(neigh.nodes)
        assert nodes[-1] == local  # canonical node is the max of its neighborhood
        row0 = nbh_ptr[-1]
        pos = {u: row0 + i for i, u in enumerate(nodes)}
        for u in nodes:
            nb = sorted(neigh.neighbors(u))
            for v in nb:
                edge_col.append(pos[v])
                edge_tri.append(1 if (with_types and edge_is_triangle(neigh, u, v)) else 0)
            edge_ptr.append(len(edge_col))
            node_gid.append(base + u)
        nbh_ptr.append(row0 + len(nodes))
    return {
        "nbh_ptr": np.asarray(nbh_ptr, dtype=np.int32),
        "node_gid": np.asarray(node_gid, dtype=np.int32),
        "edge_ptr": np.asarray(edge_ptr, dtype=np.int32),
        "edge_col": np.asarray(edge_col, dtype=np.int32),
        "edge_tri": np.asarray(edge_tri, dtype=np.uint8),
        "centre": np.asarray(centre_out, dtype=np.int32),
        "index": np.asarray(index, dtype=np.int64).reshape(-1, 2),
        "indicator": np.asarray(indicator, dtype=bool),
    }


def neighborhoods_as_networkx(batch: Dict[str, np.ndarray]):
    """Inverse of the packing: one nx.Graph per kept neighborhood (global node ids, 'type' attr) - used to feed the
    literal transforms / the PyG-shim run of the reference model."""
    out = []
    nbh_ptr, gid = batch["nbh_ptr"], batch["node_gid"]
    for g in range(len(nbh_ptr) - 1):
        lo, hi = int(nbh_ptr[g]), int(nbh_ptr[g + 1])
        G = nx.Graph()
        for r in range(lo, hi):
            G.add_node(int(gid[r]), type="canonical" if r == hi - 1 else "count")
        for r in range(lo, hi):
            for e in range(int(batch["edge_ptr"][r]), int(batch["edge_ptr"][r + 1])):
                G.add_edge(int(gid[r]), int(gid[batch["edge_col"][e]]))
        out.append(G)
    return out
