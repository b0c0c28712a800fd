"""Graph -> tensor transforms of the DeSCo hot path.

Mirrors ``subgraph_counting/transforms.py`` (reference @ 4508f7a): ``ZeroNodeFeat`` :18, ``ToTconvHetero`` :168,
``NetworkxToHetero`` :319, ``to_device`` :292, ``get_truth`` :307 - on the packed ``NeighborhoodBatch`` layout instead of
PyG ``HeteroData`` (torch_geometric is optional; PyG-style objects are adapted by duck typing).
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


def NetworkxToHetero(nx_graph, type_key: str = "type", feat_key: str = "feat", device="cuda") -> NeighborhoodBatch:
    """Drop-in name for ``transforms.py:319``: ONE neighborhood / query graph -> a packed batch of size 1 (untyped
    edges; apply ``ToTconvHetero`` for the SHMP relation split)."""
    return batch_from_networkx([nx_graph], type_key, device, typed=False)


class ToTconvHetero:
    """``transforms.py:168-255``: split every relation into ``_triangle`` / ``_tride``.  On a packed batch that is the
    per-edge flag ``edge_tri`` computed by the CUDA typing kernel; mutates its argument in place and returns it, like
    the reference (:184-187)."""

    def __init__(self, node_attr: str = "x"):
        self.node_attr = node_attr

    def __call__(self, data: NeighborhoodBatch) -> NeighborhoodBatch:
        data = as_neighborhood_batch(data)
        data.edge_tri = shmp_edge_types(data.edge_ptr, data.edge_col)
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
    if isinstance(data, NeighborhoodBatch):
        return data
    raise TypeError(
        f"expected a desco_b200 NeighborhoodBatch, got {type(data)}; build one with partition_batch / batch_from_networkx"
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
