"""PyG-shaped graph containers for the transform / module surface of the hot path.

The reference's transforms take and return ``torch_geometric.data.HeteroData`` (``subgraph_counting/transforms.py:180-255,
319-412``) and its modules read ``data.node_feature_dict`` / ``data.edge_index_dict`` (``gnn_model.py:58-63``).  When
torch_geometric is installed its own classes are used; otherwise this module supplies the small subset of
``HeteroData`` / ``Batch`` that surface needs (stores in access order, ``metadata()``, ``<attr>_dict`` views,
``from_data_list`` collate in the first element's store order), so ``NetworkxToHetero`` -> ``ToTconvHetero`` -> collate ->
``BaseGNN.forward`` reads like the reference's code path.  ``to_packed`` turns either flavour into the
``NeighborhoodBatch`` the CUDA kernels consume.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch


class _Store(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    @property
    def num_nodes(self) -> int:  # PyG NodeStorage.num_nodes: explicit, else dim 0 of a node-level tensor
        if "num_nodes" in self:
            return int(self["num_nodes"])
        for k, v in self.items():
            if isinstance(v, torch.Tensor) and (k in ("x", "feat", "pos", "batch") or "node" in k):
                return int(v.size(0))
        raise AttributeError("num_nodes")

    @property
    def num_edges(self) -> int:
        return int(self["edge_index"].size(1))


class HeteroData:
    """Subset of ``torch_geometric.data.HeteroData``: ``data[node_type].attr``, ``data[src, rel, dst].edge_index``,
    ``del data[edge_type]``, ``metadata()``, ``node_types`` / ``edge_types``, ``<attr>_dict``, ``to(device)``."""

    def __init__(self):
        object.__setattr__(self, "_node_store_dict", OrderedDict())
        object.__setattr__(self, "_edge_store_dict", OrderedDict())
        object.__setattr__(self, "_global_store", _Store())

    def __getitem__(self, key):
        if isinstance(key, tuple):
            return self._edge_store_dict.setdefault(tuple(key), _Store())
        return self._node_store_dict.setdefault(key, _Store())

    def __delitem__(self, key):
        if isinstance(key, tuple):
            del self._edge_store_dict[tuple(key)]
        else:
            del self._node_store_dict[key]

    def __setattr__(self, k, v):
        self._global_store[k] = v

    def __getattr__(self, k):
        if k.endswith("_dict") and not k.startswith("_"):
            stem = k[:-5]
            out = OrderedDict()
            for key, st in list(self._node_store_dict.items()) + list(self._edge_store_dict.items()):
                if stem in st:
                    out[key] = st[stem]
            return out
        g = object.__getattribute__(self, "_global_store")
        if k in g:
            return g[k]
        raise AttributeError(k)

    def metadata(self) -> Tuple[List[str], List[Tuple[str, str, str]]]:
        return list(self._node_store_dict.keys()), list(self._edge_store_dict.keys())

    @property
    def node_types(self):
        return list(self._node_store_dict.keys())

    @property
    def edge_types(self):
        return list(self._edge_store_dict.keys())

    @property
    def num_nodes(self) -> int:
        return sum(st.num_nodes for st in self._node_store_dict.values())

    def to(self, device):
        for st in list(self._node_store_dict.values()) + list(self._edge_store_dict.values()) + [self._global_store]:
            for k, v in st.items():
                if isinstance(v, torch.Tensor):
                    st[k] = v.to(device)
        return self


class Batch(HeteroData):
    """``Batch.from_data_list`` for the subset above (PyG ``collate``: store order of the first element, node attributes
    concatenated, ``edge_index`` offset by the cumulative node counts of its endpoint types, ``batch`` per node type)."""

    @staticmethod
    def from_data_list(data_list) -> "Batch":
        out = Batch()
        first = data_list[0]
        node_types, edge_types = first.metadata()
        for t in node_types:
            sts = [d[t] for d in data_list]
            for k, v in first[t].items():
                if isinstance(v, torch.Tensor):
                    out[t][k] = torch.cat([st[k] for st in sts], 0)
            out[t]["batch"] = torch.cat([torch.full((st.num_nodes,), i, dtype=torch.long) for i, st in enumerate(sts)])
        for et in edge_types:
            s, _, d = et
            parts, cs, cd = [], 0, 0
            for dd in data_list:
                ei = dd[et]["edge_index"]
                parts.append(ei + torch.tensor([[cs], [cd]], dtype=ei.dtype, device=ei.device))
                cs += dd[s].num_nodes
                cd += dd[d].num_nodes
            out[et]["edge_index"] = torch.cat(parts, 1)
        out.num_graphs = len(data_list)
        return out


def hetero_data_class():
    """torch_geometric's ``HeteroData`` when it is installed, else the subset above."""
    try:
        from torch_geometric.data import HeteroData as PygHeteroData  # type: ignore

        return PygHeteroData
    except Exception:
        return HeteroData


def is_hetero_data(obj) -> bool:
    return hasattr(obj, "metadata") and callable(obj.metadata) and hasattr(obj, "edge_index_dict")


def _num_nodes(data, t) -> int:
    return int(data[t].num_nodes)


def homogeneous_edges(data) -> Tuple[torch.Tensor, Dict[str, Tuple[int, int]], Dict[Tuple[str, str, str], Tuple[int, int]]]:
    """``to_homogeneous_edge_index`` (``transforms.py:258-289``): global ids by node-store offsets, edges of all types
    concatenated in edge-store order.  Returns (edge_index [2, E], node_slices, edge_slices)."""
    node_types, edge_types = data.metadata()
    node_slices, cum = {}, 0
    for t in node_types:
        n = _num_nodes(data, t)
        node_slices[t] = (cum, cum + n)
        cum += n
    parts, edge_slices, ce = [], {}, 0
    for et in edge_types:
        ei = data[et].edge_index
        s, _, d = et
        parts.append(ei + torch.tensor([[node_slices[s][0]], [node_slices[d][0]]], dtype=ei.dtype, device=ei.device))
        edge_slices[et] = (ce, ce + ei.size(1))
        ce += ei.size(1)
    ei = torch.cat(parts, 1) if parts else torch.zeros((2, 0), dtype=torch.long)
    return ei, node_slices, edge_slices


def to_packed(data, device=None):
    """PyG-shaped neighborhoods / query graphs (one ``HeteroData`` or a collated ``Batch``) -> ``NeighborhoodBatch``.

    Node types ``count`` + ``canonical`` (one canonical node per neighborhood) or a single type (``union_node``); relations
    either already split by ``ToTconvHetero`` (``*_triangle`` / ``*_tride``) or untyped (typed here by the CUDA kernel).
    Rows of a neighborhood = its count nodes IN THE DATA'S OWN ORDER, then its canonical node - so the edge that the
    reference's ``remove_self_loops`` quirk drops (``gnn_model.py:389-390``: first count node of the collated batch
    position) is reproduced exactly for PyG-shaped input, whatever the node order networkx produced."""
    from .data import NeighborhoodBatch, _require_cuda, shmp_edge_types

    dev = _require_cuda(device)
    node_types, edge_types = data.metadata()
    hetero = "canonical" in node_types
    if hetero and set(node_types) != {"count", "canonical"}:
        raise ValueError(f"unsupported node types {node_types}")
    if not hetero and len(node_types) != 1:
        raise ValueError(f"unsupported node types {node_types}")

    def batch_vec(t):
        st = data[t]
        if "batch" in st:
            return st["batch"].cpu().numpy().astype(np.int64)
        return np.zeros(_num_nodes(data, t), dtype=np.int64)

    main = "count" if hetero else node_types[0]
    bm = batch_vec(main)
    G = int(max(bm.max(initial=-1), batch_vec("canonical").max(initial=-1) if hetero else -1)) + 1
    ncount = np.bincount(bm, minlength=G)
    if np.any(np.diff(bm) < 0):
        raise ValueError("nodes of a collated batch must be grouped by graph")
    sizes = ncount + (1 if hetero else 0)
    nbh_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    cstart = np.concatenate([[0], np.cumsum(ncount)])[:-1]
    row_of = {main: nbh_ptr[bm] + (np.arange(len(bm)) - cstart[bm])}
    if hetero:
        ba = batch_vec("canonical")
        if not np.array_equal(np.sort(ba), np.arange(G)):
            raise ValueError("every neighborhood needs exactly one canonical node")
        row_of["canonical"] = nbh_ptr[ba + 1] - 1
    V = int(nbh_ptr[-1])
    src, dst, tri, typed = [], [], [], None
    for et in edge_types:
        s, rel, d = et
        ei = data[et].edge_index.cpu().numpy()
        if ei.shape[1] == 0:
            is_typed = rel.endswith("_triangle") or rel.endswith("_tride")
        else:
            is_typed = rel.endswith("_triangle") or rel.endswith("_tride")
        typed = is_typed if typed is None else typed
        if typed != is_typed:
            raise ValueError("mixed typed / untyped relations")
        src.append(row_of[s][ei[0]])
        dst.append(row_of[d][ei[1]])
        tri.append(np.full(ei.shape[1], 1 if rel.endswith("_triangle") else 0, dtype=np.uint8))
    src = np.concatenate(src) if src else np.zeros(0, dtype=np.int64)
    dst = np.concatenate(dst) if dst else np.zeros(0, dtype=np.int64)
    tri = np.concatenate(tri) if tri else np.zeros(0, dtype=np.uint8)
    keep = src != dst
    src, dst, tri = src[keep], dst[keep], tri[keep]
    order = np.lexsort((src, dst))  # a packed row holds its INCOMING edges, sources ascending
    src, dst, tri = src[order], dst[order], tri[order]
    edge_ptr = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=V))])
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)
    ep, ec = t(edge_ptr, np.int32), t(src, np.int32)
    et_dev = t(tri, np.uint8) if typed else shmp_edge_types(ep, ec)
    node_gid = torch.arange(V, dtype=torch.int32, device=dev)
    nbh = t(nbh_ptr, np.int32)
    batch = NeighborhoodBatch(nbh, node_gid, ep, ec, et_dev, node_gid[(nbh[1:] - 1).long()] if V else node_gid, None, None,
                              None, G, V, int(len(src)), hetero=hetero, max_rows=int(sizes.max()) if G else 0)
    # node features in packed row order (None when they are all zero: ZeroNodeFeat, the kernels' fast path)
    feats = data.node_feature_dict if hasattr(data, "node_feature_dict") else {}
    if feats and all(tp in feats for tp in node_types):
        width = next(iter(feats.values())).shape[1]
        x = torch.zeros((V, width), dtype=torch.float32)
        for tp in node_types:
            x[torch.from_numpy(row_of[tp])] = feats[tp].detach().to("cpu", torch.float32)
        if bool((x != 0).any()):
            batch._cache["feat"] = x.to(dev)
    batch._cache["pyg_collated"] = True  # the whole input IS one collated PyG batch (quirk evaluated over it)
    return batch
