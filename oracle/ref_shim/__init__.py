"""Run the reference's OWN ``subgraph_counting/gnn_model.py`` without torch_geometric.  TEST INFRASTRUCTURE ONLY.

The reference model classes (``SAGEConv``, ``GossipConv``, ``BaseGNNCore``, ``BaseGNN``) need five PyG primitives:
``nn.MessagePassing`` (``propagate`` with ``aggr="add"``), ``nn.global_add_pool``, ``utils.remove_self_loops``,
``utils.to_undirected`` and (import only) ``torch_sparse``.  This module installs ~100 lines of stand-ins for those
primitives into ``sys.modules`` (PyG 2.2.0 semantics) and imports the reference file from
``/root/reference`` unmodified.  It is used ONLY by ``tests/golden/make_golden.py`` and by CPU tests in the builder
container to pin ``oracle/model.py``; it never runs on the GPU box (``/root/reference`` is absent there).

What this does NOT cover: ``pyg.nn.to_hetero`` (torch.fx rewrite).  The hetero expansion stays a restatement.
"""
from __future__ import annotations

import importlib
import inspect
import os
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


class MessagePassing(torch.nn.Module):
    """PyG ``MessagePassing`` subset: flow source_to_target, ``edge_index[0]`` = j (source), ``[1]`` = i (target);
    ``*_j`` / ``*_i`` arguments of ``message`` are gathered from the propagate kwarg of the same stem; aggregation is a
    scatter-sum over the targets; ``update`` receives the aggregate plus matching kwargs."""

    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        assert aggr in ("add", "sum")
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        x = kwargs.get("x")
        x_pair = x if isinstance(x, (tuple, list)) else (x, x)
        n_dst = size[1] if size is not None else x_pair[1].shape[0]
        margs = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_j") or name.endswith("_i"):
                val = kwargs[name[:-2]]
                pair = val if isinstance(val, (tuple, list)) else (val, val)
                margs[name] = pair[0][src] if name.endswith("_j") else pair[1][dst]
            else:
                margs[name] = kwargs.get(name)
        msg = self.message(**margs)
        out = torch.zeros((n_dst,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        out.index_add_(0, dst, msg)
        uargs = {}
        for name in list(inspect.signature(self.update).parameters)[1:]:
            val = kwargs.get(name)
            uargs[name] = val[1] if (name == "x" and isinstance(val, (tuple, list))) else val
        return self.update(out, **uargs)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out


def global_add_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    out = torch.zeros((size,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    return out.index_add_(0, batch, x)


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], (None if edge_attr is None else edge_attr[mask])


def to_undirected(edge_index, num_nodes=None):
    row = torch.cat([edge_index[0], edge_index[1]])
    col = torch.cat([edge_index[1], edge_index[0]])
    n = int(max(row.max(), col.max())) + 1 if num_nodes is None else num_nodes
    key = torch.unique(row * n + col)  # coalesce: sorted row-major, duplicates dropped
    return torch.stack([key // n, key % n])


def install():
    """Register the stand-ins and return the reference's ``gnn_model`` module (None if the tree is absent)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return None
    if "torch_geometric" not in sys.modules:
        pyg = types.ModuleType("torch_geometric")
        pyg_nn = types.ModuleType("torch_geometric.nn")
        pyg_utils = types.ModuleType("torch_geometric.utils")
        pyg_nn.MessagePassing = MessagePassing
        pyg_nn.global_add_pool = global_add_pool
        for missing in ("GCNConv", "GATConv", "PNAConv", "global_mean_pool"):
            setattr(pyg_nn, missing, None)
        pyg_utils.remove_self_loops = remove_self_loops
        pyg_utils.to_undirected = to_undirected
        pyg.nn, pyg.utils = pyg_nn, pyg_utils
        sys.modules["torch_geometric"] = pyg
        sys.modules["torch_geometric.nn"] = pyg_nn
        sys.modules["torch_geometric.utils"] = pyg_utils
        sys.modules.setdefault("torch_sparse", types.ModuleType("torch_sparse"))
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    return importlib.import_module("subgraph_counting.gnn_model")
