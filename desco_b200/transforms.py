"""Graph -> tensor transforms of the DeSCo hot path.

Mirrors ``subgraph_counting/transforms.py`` (reference @ 4508f7a): ``ZeroNodeFeat`` :18, ``ToTconvHetero`` :168,
``ToTCONV`` :45, ``NetworkxToHetero`` :319, ``to_device`` :292, ``get_truth`` :307.  The transforms take and return
PyG-shaped ``HeteroData`` like the reference's (torch_geometric's class when installed, ``desco_b200.hetero``'s subset
otherwise) and also accept the packed ``NeighborhoodBatch`` the batched partition emits; the SHMP test itself always
runs in the CUDA typing kernel.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .data import NeighborhoodBatch, shmp_edge_types


def batch_from_networkx(neighs: List, type_key: str = "type", device="cuda", typed: bool = True) -> NeighborhoodBatch:
    """``NetworkxToHetero`` (transforms.py:319-412) + collate for a LIST of neighborhood graphs: nodes whose ``type`` is
    "canonical" become the last row of their neighborhood, every other node a count row (ascending label order)."""
    nbh_ptr, edge_ptr, edge_col, gid = [0], [0], [], []
    hetero = None
    for g in neighs:
        canon = [u for u in g.nodes if g.nodes[u].get(type_key) == "canonical"]
        h = len(canon) == 1
        hetero = h if hetero is None else hetero
        if h != hetero:
            raise ValueError("mixed typed / untyped neighborhoods in one batch")
        nodes = sorted(u for u in g.nodes if u not in canon) + canon
        row0 = nbh_ptr[-1]
        pos = {u: row0 + i for i, u in enumerate(nodes)}
        for u in nodes:
            edge_col.extend(sorted(pos[v] for v in g.neighbors(u) if v != u))
            edge_ptr.append(len(edge_col))
            gid.append(int(u) if isinstance(u, (int, np.integer)) else len(gid))
        nbh_ptr.append(row0 + len(nodes))
    dev = torch.device(device)
    t = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
    ep, ec = t(edge_ptr), t(edge_col)
    tri = shmp_edge_types(ep, ec) if typed else torch.zeros(len(edge_col), dtype=torch.uint8, device=dev)
    V = nbh_ptr[-1]
    ng = t(gid)
    return NeighborhoodBatch(t(nbh_ptr), ng, ep, ec, tri, ng[(t(nbh_ptr)[1:] - 1).long()] if V else ng, None, None, None,
                             len(neighs), V, len(edge_col), hetero=bool(hetero),
                             max_rows=max((b - a for a, b in zip(nbh_ptr[:-1], nbh_ptr[1:])), default=0))


def NetworkxToHetero(nx_graph, type_key: str = "type", feat_key: str = "feat"):
    """``transforms.py:319-412``: ONE networkx graph -> ``HeteroData`` (torch_geometric's when installed, else
    ``desco_b200.hetero.HeteroData``).  Same contract as the reference: nodes without ``type_key`` become
    ``"union_node"``; per-type local ids follow the graph's node iteration order (:342-348); every directed edge goes
    to ``(type[u], edge type or "union", type[v])`` (:351-367); ``node_feature`` = the stacked ``feat_key`` tensors
    (``zeros(1)`` per node when absent, :380-384), every other node attribute is stacked under its own name."""
    from .hetero import hetero_data_class

    g = nx_graph.to_directed()
    data = hetero_data_class()()
    ids, edges = {}, {}
    for u in g.nodes:
        t = g.nodes[u].setdefault(type_key, "union_node")
        members = ids.setdefault(t, {})
        members[u] = len(members)
    for a, b in g.edges:
        et = (g.nodes[a][type_key], g.edges[a, b].get(type_key, "union"), g.nodes[b][type_key])
        edges.setdefault(et, []).append((ids[et[0]][a], ids[et[2]][b]))
    attrs = [k for k in next(iter(g.nodes(data=True)))[-1].keys() if k != type_key] if len(g) else []
    if feat_key not in attrs:
        for u in g.nodes:
            g.nodes[u][feat_key] = torch.zeros(1)
    else:
        attrs.remove(feat_key)
    as_t = lambda v: (v if isinstance(v, torch.Tensor) else torch.tensor(v)).view(-1)
    for t, members in ids.items():
        order = sorted(members, key=members.get)
        if feat_key is not None:
            data[t].node_feature = torch.stack([as_t(g.nodes[u][feat_key]) for u in order], 0)
        for k in attrs:
            setattr(data[t], k, torch.stack([as_t(g.nodes[u][k]) for u in order], 0))
    for et, lst in edges.items():
        data[et].edge_index = torch.tensor(lst, dtype=torch.long).T
    return data


def _tri_flags(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """Per directed edge of ``edge_index`` [2, E] (any order, duplicates allowed): does it lie in a triangle of the graph
    the edges span - ``(A * A@A + A) > 1`` of ``transforms.py:201-221`` - computed by the CUDA typing kernel on the
    row-sorted, deduplicated edge set and scattered back to the input order."""
    from .data import _require_cuda

    dev = _require_cuda(None)
    ei = edge_index.to(dev).long()
    if ei.size(1) == 0:
        return torch.zeros(0, dtype=torch.bool, device=edge_index.device)
    key = ei[1] * num_nodes + ei[0]  # packed rows hold incoming edges: sort by (dst, src)
    uniq, inverse = torch.unique(key, return_inverse=True)
    dst, src = uniq // num_nodes, uniq % num_nodes
    loops = dst == src  # self loops lie on no triangle and are not adjacency for the common-neighbour test
    edge_ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=dev)
    edge_ptr[1:] = torch.cumsum(torch.bincount(dst[~loops], minlength=num_nodes), 0)
    tri_uniq = torch.zeros(uniq.numel(), dtype=torch.bool, device=dev)
    tri_uniq[~loops] = shmp_edge_types(edge_ptr.to(torch.int32), src[~loops].to(torch.int32).contiguous()).bool()
    return tri_uniq[inverse].to(edge_index.device)


class ToTconvHetero:
    """``transforms.py:168-255``: split every relation ``(s, r, d)`` into ``(s, r + "_triangle", d)`` and
    ``(s, r + "_tride", d)`` (both always present, possibly ``[2, 0]``) by whether the edge has a common neighbour in
    the WHOLE graph of the sample (all node types together), delete ``(s, r, d)``.  Mutates its argument in place and
    returns it, like the reference (:184-187).  The test itself runs on the GPU (``edge_types_kernel``).  A packed
    ``NeighborhoodBatch`` is accepted too (its ``edge_tri`` is recomputed)."""

    def __init__(self, node_attr: str = "x"):
        self.node_attr = node_attr

    def __call__(self, data):
        from .hetero import homogeneous_edges, is_hetero_data

        if isinstance(data, NeighborhoodBatch):
            data.edge_tri = shmp_edge_types(data.edge_ptr, data.edge_col)
            return data
        if not is_hetero_data(data):
            raise TypeError(f"ToTconvHetero expects HeteroData or a NeighborhoodBatch, got {type(data)}")
        ei, node_slices, edge_slices = homogeneous_edges(data)
        n = max(hi for _, hi in node_slices.values()) if node_slices else 0
        tri = _tri_flags(ei, n)
        for (s, r, d), (e0, e1) in edge_slices.items():
            local = data[s, r, d].edge_index
            flag = tri[e0:e1].to(local.device)
            data[s, r + "_triangle", d].edge_index = local[:, flag]
            data[s, r + "_tride", d].edge_index = local[:, ~flag]
            del data[s, r, d]
        return data


class ToTCONV:
    """``transforms.py:45-115``: the single-node-type variant - only relations from ``node_type`` to ``node_type`` are
    split, each by the triangles of ITS OWN edge set; self loops are removed and the emitted edges are in coalesced
    (row-major sorted) order, as in the reference."""

    def __init__(self, node_type: str = "count", node_attr: str = "x"):
        self.node_type, self.node_attr = node_type, node_attr

    def __call__(self, data):
        for et in [e for e in data.metadata()[1] if e[0] == self.node_type and e[2] == self.node_type]:
            s, r, d = et
            ei = data[et].edge_index
            del data[et]
            if ei.numel() == 0:
                data[s, r + "_triangle", d].edge_index = ei
                data[s, r + "_tride", d].edge_index = ei
                continue
            ei = ei[:, ei[0] != ei[1]]
            n = int(getattr(data[self.node_type], self.node_attr).shape[0])
            key = torch.unique(ei[0] * n + ei[1])  # coalesce(): sorted row-major, duplicates merged
            ei = torch.stack([key // n, key % n])
            flag = _tri_flags(ei, n)
            data[s, r + "_triangle", d].edge_index = ei[:, flag]
            data[s, r + "_tride", d].edge_index = ei[:, ~flag]
        return data


class ZeroNodeFeat:
    """``transforms.py:18-42``: zero node features.  The packed batch carries no feature tensor when features are zero
    (the kernels read feat == NULL as zeros), so this only records the width."""

    def __init__(self, node_feat_name: str = "x", node_feat_len: Optional[int] = None):
        self.node_feat_name, self.node_feat_len = node_feat_name, node_feat_len

    def __call__(self, data):
        n = getattr(data, "num_rows", None) or getattr(data, "num_nodes")
        if self.node_feat_len is None:
            x = getattr(data, self.node_feat_name, None)
            self.node_feat_len = x.shape[1] if x is not None else 1
        dev = data.edge_ptr.device if hasattr(data, "edge_ptr") else "cpu"
        setattr(data, self.node_feat_name, torch.zeros(n, self.node_feat_len, device=dev))
        return data


def as_neighborhood_batch(data) -> NeighborhoodBatch:
    """The packed batch the kernels consume: a ``NeighborhoodBatch`` as is, PyG-shaped ``HeteroData`` / ``Batch``
    (torch_geometric's or ``desco_b200.hetero``'s) through ``hetero.to_packed``."""
    from .hetero import is_hetero_data, to_packed

    if isinstance(data, NeighborhoodBatch):
        return data
    if is_hetero_data(data):
        return to_packed(data)
    raise TypeError(
        f"expected a NeighborhoodBatch or HeteroData, got {type(data)}; build one with partition_batch / NetworkxToHetero"
    )


def to_device(data, device):
    """``transforms.py:292-304``."""
    if isinstance(data, NeighborhoodBatch):
        for k in ("nbh_ptr", "node_gid", "edge_ptr", "edge_col", "edge_tri", "centre"):
            setattr(data, k, getattr(data, k).to(device))
        return data
    if hasattr(data, "to"):
        return data.to(device)
    raise NotImplementedError


def get_truth(data):
    """``transforms.py:307-316``."""
    if hasattr(data, "y"):
        return data.y
    raise NotImplementedError
