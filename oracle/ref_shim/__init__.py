"""Run the reference's OWN ``subgraph_counting/gnn_model.py`` without torch_geometric.  TEST INFRASTRUCTURE ONLY.

The reference model classes (``SAGEConv``, ``GossipConv``, ``BaseGNNCore``, ``BaseGNN``) need five PyG primitives:
``nn.MessagePassing`` (``propagate`` with ``aggr="add"``), ``nn.global_add_pool``, ``utils.remove_self_loops``,
``utils.to_undirected`` and (import only) ``torch_sparse``.  This module installs ~100 lines of stand-ins for those
primitives into ``sys.modules`` (PyG 2.2.0 semantics) and imports the reference file from
``/root/reference`` unmodified.  It is used ONLY by ``tests/golden/make_golden.py`` and by CPU tests in the builder
container to pin ``oracle/model.py``; it never runs on the GPU box (``/root/reference`` is absent there).

Second tier (``install_full``): the transform / entry-point surface.  ``HeteroData`` / ``Data`` / ``Batch`` (collate in
PyG's store order), ``BaseTransform``, ``utils.sort_edge_index`` and ``nn.to_hetero`` - the latter
as a real ``torch.fx`` trace of the reference's ``BaseGNNCore.forward`` that is re-executed per node / edge type
(``to_hetero.py`` of PyG 2.2.0: MessagePassing modules are leaves duplicated per edge type and called with ``x_src`` or
``(x_src, x_dst)``; every other module / function is duplicated per node type; outputs with the same destination type
are added pairwise in metadata order; duplicated modules are re-initialised with ``reset_parameters``).  With it the
reference's own ``NetworkxToHetero``, ``ToTconvHetero`` (``transforms.py``) and - executed from source -
``NeighborhoodCountingModel.to_hetero_old`` / ``graph_to_count`` / ``embed_to_count`` (``lightning_model.py``) run
unmodified, so the hetero SHMP fixture comes from the reference's forward, not from a restated loop.
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


# ---------------------------------------------------------------------------------------------
# second tier: HeteroData / Batch / transforms / to_hetero
# ---------------------------------------------------------------------------------------------
import ast
import copy
from collections import OrderedDict


class _Store(dict):
    """PyG ``BaseStorage`` subset: attribute access over a dict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    @property
    def num_nodes(self):  # NodeStorage.num_nodes: explicit, else dim 0 of a node-level tensor (N_KEYS or "node" in key)
        if "num_nodes" in self:
            return self["num_nodes"]
        for k, v in self.items():
            if isinstance(v, torch.Tensor) and (k in ("x", "feat", "pos", "batch") or "node" in k):
                return v.size(0)
        raise AttributeError("num_nodes")

    @property
    def num_edges(self):
        return self["edge_index"].size(1)


class HeteroData:
    """PyG ``HeteroData`` subset: node stores keyed by type (str), edge stores by (src, rel, dst), created on first access
    IN ACCESS ORDER (the order ``to_homogeneous`` / ``collate`` follow - the node-type-order trap)."""

    def __init__(self):
        object.__setattr__(self, "_node_store_dict", OrderedDict())
        object.__setattr__(self, "_edge_store_dict", OrderedDict())
        object.__setattr__(self, "_global", _Store())

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
        self._global[k] = v

    def __getattr__(self, k):
        if k.endswith("_dict") and not k.startswith("_"):  # x_dict / node_feature_dict / edge_index_dict
            stem = k[:-5]
            out = OrderedDict()
            for key, st in list(self._node_store_dict.items()) + list(self._edge_store_dict.items()):
                if stem in st:
                    out[key] = st[stem]
            return out
        g = object.__getattribute__(self, "_global")
        if k in g:
            return g[k]
        raise AttributeError(k)

    def metadata(self):
        return list(self._node_store_dict.keys()), list(self._edge_store_dict.keys())

    @property
    def node_types(self):
        return list(self._node_store_dict.keys())

    @property
    def edge_types(self):
        return list(self._edge_store_dict.keys())

    @property
    def num_nodes(self):
        return sum(st.num_nodes for st in self._node_store_dict.values())

    def to(self, device):
        for st in list(self._node_store_dict.values()) + list(self._edge_store_dict.values()) + [self._global]:
            for k, v in st.items():
                if isinstance(v, torch.Tensor):
                    st[k] = v.to(device)
        return self

    def clone(self):
        return copy.deepcopy(self)

    def to_homogeneous(self):
        """Node attributes (incl. ``batch``) concatenated in node-store order; edges offset by the store offsets."""
        off, cum = {}, 0
        for t, st in self._node_store_dict.items():
            off[t] = cum
            cum += st.num_nodes
        out = Data()
        keys = set.intersection(*[set(k for k, v in st.items() if isinstance(v, torch.Tensor)) for st in self._node_store_dict.values()])
        for k in keys:
            setattr(out, k, torch.cat([st[k] for st in self._node_store_dict.values()], 0))
        eis = [st["edge_index"] + torch.tensor([[off[s]], [off[d]]]) for (s, _, d), st in self._edge_store_dict.items()]
        out.edge_index = torch.cat(eis, 1) if eis else torch.zeros(2, 0, dtype=torch.long)
        return out


class Data(_Store):
    def to(self, device):
        for k, v in self.items():
            if isinstance(v, torch.Tensor):
                self[k] = v.to(device)
        return self


class Batch(HeteroData):
    @staticmethod
    def from_data_list(data_list):
        """PyG ``collate`` for HeteroData: store order of the FIRST element; node attributes concatenated; ``edge_index``
        offset by the cumulative (n_src, n_dst) of the earlier elements; ``batch`` vector per node type; graph-level
        tensors concatenated along dim 0."""
        out = Batch()
        first = data_list[0]
        for t in first._node_store_dict:
            sts = [d[t] for d in data_list]
            for k in first[t]:
                if isinstance(first[t][k], torch.Tensor):
                    out[t][k] = torch.cat([st[k] for st in sts], 0)
            out[t]["batch"] = torch.cat([torch.full((st.num_nodes,), i, dtype=torch.long) for i, st in enumerate(sts)])
        for et in first._edge_store_dict:
            s, _, d = et
            parts, cs, cd = [], 0, 0
            for dd in data_list:
                parts.append(dd[et]["edge_index"] + torch.tensor([[cs], [cd]]))
                cs += dd[s].num_nodes
                cd += dd[d].num_nodes
            out[et]["edge_index"] = torch.cat(parts, 1)
        for k, v in first._global.items():
            if isinstance(v, torch.Tensor):
                setattr(out, k, torch.cat([dd._global[k] for dd in data_list], 0))
        out.num_graphs = len(data_list)
        return out


class BaseTransform:
    def __call__(self, data):
        raise NotImplementedError

    def __repr__(self):
        return f"{self.__class__.__name__}()"


def sort_edge_index(edge_index, edge_attr=None, num_nodes=None, sort_by_row=True):
    """PyG 2.2 ``utils.sort_edge_index``: stable row-major sort (``row * n + col``)."""
    n = int(edge_index.max()) + 1 if num_nodes is None and edge_index.numel() else (num_nodes or 0)
    idx = edge_index[1 - int(sort_by_row)] * n + edge_index[int(sort_by_row)]
    perm = idx.argsort(stable=True)
    ei = edge_index[:, perm]
    return ei if edge_attr is None else (ei, edge_attr[perm])


def _key2str(key):
    return "__".join(key) if isinstance(key, tuple) else key


class _HeteroGraphModule(torch.nn.Module):
    """The result of ``to_hetero``: the fx graph of the original ``forward`` re-executed per node / edge type."""

    def __init__(self, module, metadata, aggr):
        super().__init__()
        import torch.fx as fx

        assert aggr == "sum"

        class _Tracer(fx.Tracer):  # PyG traces through containers and keeps MessagePassing (and torch.nn) leaves whole
            def is_leaf_module(self, m, qualname):
                return isinstance(m, MessagePassing) or (super().is_leaf_module(m, qualname))

        graph = _Tracer().trace(module)
        object.__setattr__(self, "_graph", graph)
        object.__setattr__(self, "_metadata", metadata)
        self.training = module.training
        node_types, edge_types = metadata
        done = set()
        for n in graph.nodes:
            if n.op != "call_module" or n.target in done:
                continue
            done.add(n.target)
            sub = module.get_submodule(n.target)
            keys = edge_types if isinstance(sub, MessagePassing) else node_types
            md = torch.nn.ModuleDict()
            for key in keys:
                md[_key2str(key)] = copy.deepcopy(sub)
                if len(keys) > 1 and hasattr(sub, "reset_parameters"):
                    md[_key2str(key)].reset_parameters()
            # register under the original dotted path so the state-dict keys read <path>.<type>.<param> like PyG's
            parent, parts = self, n.target.split(".")
            for p in parts[:-1]:
                if not hasattr(parent, p):
                    parent.add_module(p, torch.nn.Module())
                parent = getattr(parent, p)
            parent.add_module(parts[-1], md)

    def forward(self, x, edge_index, query_emb=None):
        import torch.fx as fx

        node_types, edge_types = self._metadata
        env = {}

        def per_type(arg, t):  # value of an fx argument for node type t
            return fx.node.map_arg(arg, lambda n: env[n][t] if isinstance(env[n], dict) else env[n])

        for n in self._graph.nodes:
            if n.op == "placeholder":
                env[n] = {"x": x, "edge_index": edge_index}.get(n.target, query_emb if n.target == "query_emb" else None)
            elif n.op == "get_attr":
                env[n] = None
            elif n.op == "call_module":
                md = self.get_submodule(n.target)
                first = next(iter(md.values()))
                if isinstance(first, MessagePassing):
                    xin, ein = n.args[0], n.args[1]
                    outs = {t: [] for t in node_types}
                    for et in edge_types:
                        s, _, d = et
                        xs = env[xin][s] if s == d else (env[xin][s], env[xin][d])
                        outs[d].append(md[_key2str(et)](xs, env[ein][et]))
                    red = {}
                    for t, v in outs.items():  # pairwise tree reduction in metadata order (aggr="sum")
                        while len(v) > 1:
                            v = [v[i] + v[i + 1] if i + 1 < len(v) else v[i] for i in range(0, len(v), 2)]
                        red[t] = v[0]
                    env[n] = red
                else:
                    env[n] = {t: md[t](*per_type(n.args, t), **per_type(n.kwargs, t)) for t in node_types}
            elif n.op == "call_function":
                env[n] = {t: n.target(*per_type(n.args, t), **per_type(n.kwargs, t)) for t in node_types}
            elif n.op == "call_method":
                def call(t):
                    a = per_type(n.args, t)
                    return getattr(a[0], n.target)(*a[1:], **per_type(n.kwargs, t))
                env[n] = {t: call(t) for t in node_types}
            elif n.op == "output":
                return env[n.args[0]]
        raise AssertionError("fx graph without output")


def to_hetero(module, metadata, aggr="sum", input_map=None, debug=False):
    return _HeteroGraphModule(module, (list(metadata[0]), [tuple(e) for e in metadata[1]]), aggr)


def install_full():
    """``install()`` plus the data / transform / to_hetero stand-ins.  Returns (gnn_model, transforms) of the reference."""
    ref = install()
    if ref is None:
        return None, None
    pyg = sys.modules["torch_geometric"]
    if not hasattr(pyg, "data"):
        pyg_data = types.ModuleType("torch_geometric.data")
        pyg_data.HeteroData, pyg_data.Data, pyg_data.Batch = HeteroData, Data, Batch
        pyg_data.data = types.SimpleNamespace(Data=Data)
        pyg_tr = types.ModuleType("torch_geometric.transforms")
        pyg_tr.BaseTransform = BaseTransform
        pyg_ty = types.ModuleType("torch_geometric.typing")
        pyg_ty.EdgeType, pyg_ty.NodeType, pyg_ty.QueryType = tuple, str, object
        pyg.data, pyg.transforms, pyg.typing = pyg_data, pyg_tr, pyg_ty
        pyg.nn.to_hetero = to_hetero
        pyg.utils.sort_edge_index = sort_edge_index
        sys.modules["torch_geometric.data"] = pyg_data
        sys.modules["torch_geometric.transforms"] = pyg_tr
        sys.modules["torch_geometric.typing"] = pyg_ty
    return ref, importlib.import_module("subgraph_counting.transforms")


def load_lightning_methods(names=("to_hetero_old", "to_hetero", "graph_to_count", "embed_to_count", "graph_to_embed")):
    """The named methods of the reference's ``NeighborhoodCountingModel`` executed FROM SOURCE
    (``lightning_model.py`` itself needs pytorch_lightning to import): plain functions taking ``self``."""
    path = os.path.join(REFERENCE_ROOT, "subgraph_counting", "lightning_model.py")
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "NeighborhoodCountingModel")
    body = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]
    for fn in body:
        fn.returns = None
        for a in fn.args.args:
            a.annotation = None
    ns = {"torch": torch, "pyg": sys.modules["torch_geometric"], "Tuple": tuple}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return {n: ns[n] for n in names if n in ns}
