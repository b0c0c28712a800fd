"""Model entry points of the DeSCo hot path.

Mirrors ``subgraph_counting/lightning_model.py`` (reference @ 4508f7a): ``gen_queries`` :37,
``NeighborhoodCountingModel`` :90 (``graph_to_count`` :198, ``graph_to_embed`` :224, ``embed_to_count`` :176,
``set_queries`` :291, ``get_query_emb`` :311, ``to_hetero_old`` :371, ``predict_step`` :195) and ``GossipCountingModel``
:535 (``graph_to_count`` :613, ``set_query_emb`` :637, ``_gate_value`` :640).  pytorch_lightning is optional: the
classes are ``torch.nn.Module`` and expose the Lightning hook names the reference's ``main.py`` calls.
"""
from __future__ import annotations

import warnings
from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .data import NeighborhoodBatch, _ptr, _stream, shmp_edge_types
from .gnn_model import PRECISION, QUERY_META, TARGET_META, BaseGNN, GossipBaseGNN, _PackedWeightsMixin

STANDARD_QUERY_IDS = [6, 7, 13, 14, 15, 16, 17, 18, 29, 30, 31, 34, 35, 36, 37, 38, 40, 41, 42, 43, 44, 45, 46, 47,
                      48, 49, 50, 51, 52]  # gen_query_ids([3,4,5]), data.py:37-58


def default_neighborhood_args(**kw) -> SimpleNamespace:
    """``config.py:247-264``."""
    d = dict(conv_type="SAGE", layer_num=8, hidden_dim=64, input_dim=1, dropout=0.0, use_hetero=True, use_tconv=True,
             depth=4, lr=1e-4, weight_decay=0.0, batch_size=512, use_canonical=True)
    d.update(kw)
    return SimpleNamespace(**d)


def gen_queries(query_ids: Optional[List[int]], queries=None, transform=None, node_feat_len: int = 1, hetero=True,
                device="cuda"):
    """``lightning_model.py:37-87``: atlas ids / nx graphs -> (packed query batch, nx graphs).  The SHMP transform
    (``ToTconvHetero`` on every query, :84-85) is applied on the GPU by the same typing kernel as the targets."""
    import networkx as nx

    queries_nx = [nx.graph_atlas(i) for i in query_ids] if queries is None else list(queries)
    nbh_ptr, edge_ptr, edge_col = [0], [0], []
    for g in queries_nx:
        nodes = sorted(g.nodes)
        row0 = nbh_ptr[-1]
        pos = {u: row0 + i for i, u in enumerate(nodes)}
        for u in nodes:
            edge_col.extend(sorted(pos[v] for v in g.neighbors(u) if v != u))
            edge_ptr.append(len(edge_col))
        nbh_ptr.append(row0 + len(nodes))
    dev = torch.device(device)
    t = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
    ep, ec = t(edge_ptr), t(edge_col)
    tri = shmp_edge_types(ep, ec) if transform is not False else torch.zeros(len(edge_col), dtype=torch.uint8, device=dev)
    V = nbh_ptr[-1]
    batch = NeighborhoodBatch(t(nbh_ptr), torch.arange(V, dtype=torch.int32, device=dev), ep, ec, tri,
                              t(nbh_ptr[1:]) - 1, None, None, None, len(queries_nx), V, len(edge_col), hetero=False)
    return batch, queries_nx


def pack_head_weights(count_model: nn.Sequential, hidden: int) -> torch.Tensor:
    W1 = count_model[0].weight.detach().to("cpu", torch.float64)  # [4h, 2h]
    parts = [W1[:, :hidden].t().contiguous().flatten(), W1[:, hidden:].t().contiguous().flatten(),
             count_model[0].bias.detach().to("cpu", torch.float64), count_model[2].weight.detach().to("cpu", torch.float64).flatten(),
             count_model[2].bias.detach().to("cpu", torch.float64)]
    return torch.cat(parts).to(torch.float32).to(count_model[0].weight.device).contiguous()


def pack_head_weights_tc(count_model: nn.Sequential, hidden: int) -> torch.Tensor:
    """Tensor-core operand images of the two halves of count_model[0] (csrc/dense_tc.cu)."""
    from .tcpack import pack_dense_tc

    W1 = count_model[0].weight.detach().to("cpu", torch.float64)  # [4h, 2h]
    return torch.cat([pack_dense_tc(W1[:, :hidden], 128), pack_dense_tc(W1[:, hidden:], 128)]).to(count_model[0].weight.device)



def _load_lightning_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **override):
    """``pl.LightningModule.load_from_checkpoint`` for the reference's ``.ckpt`` files (``main.py:216-233,319-334``): a
    pickled dict with ``state_dict`` (keys as produced by ``to_hetero_old``, SURVEY App. B.3 - the modules here use the
    same names) and ``hyper_parameters`` (the constructor arguments saved by ``save_hyperparameters()``)."""
    try:  # tensors + plain containers + the argparse.Namespace of `args` only: no arbitrary pickle execution
        import argparse
        from types import SimpleNamespace as _SN

        with torch.serialization.safe_globals([argparse.Namespace, _SN]):
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=True)
    except Exception as safe_err:
        # real Lightning checkpoints also pickle pytorch_lightning's AttributeDict (and need Lightning importable):
        # opt in explicitly, the file is then trusted like any pickle
        if not override.pop("trust_checkpoint", False):
            raise RuntimeError(
                f"{checkpoint_path}: not loadable with weights_only=True ({safe_err}); pass trust_checkpoint=True to unpickle "
                "it fully (executes arbitrary code from the file; Lightning-written checkpoints also need pytorch_lightning)")
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
    override.pop("trust_checkpoint", None)
    hp = dict(ckpt.get("hyper_parameters", {}))
    hp.update(override)
    model = cls(hp.pop("input_dim", 1), hp.pop("hidden_dim", 64), hp.pop("args", None), **hp)
    if hasattr(model, "on_load_checkpoint"):
        model.on_load_checkpoint(ckpt)
    model.load_state_dict(ckpt["state_dict"], strict=strict)
    return model


class NeighborhoodCountingModel(_PackedWeightsMixin, nn.Module):
    def __init__(self, input_dim=1, hidden_dim=64, args=None, **kwargs):
        super().__init__()
        args = args or default_neighborhood_args(hidden_dim=hidden_dim, input_dim=input_dim)
        self.hidden_dim, self.input_dim = hidden_dim, input_dim
        self.args, self.kwargs = args, kwargs
        for k, v in vars(args).items():  # lightning_model.py:113-114
            if not hasattr(self, k):
                setattr(self, k, v)
        self.query_loader = None  # the reference holds a DataLoader of query graphs; here: one packed batch
        if getattr(args, "use_hetero", True):
            self.emb_model = BaseGNN(input_dim, hidden_dim, hidden_dim, args, TARGET_META, emb_channels=hidden_dim, **kwargs)
            self.emb_model_query = BaseGNN(input_dim, hidden_dim, hidden_dim, args, QUERY_META, emb_channels=hidden_dim, **kwargs)
        else:  # hetero_graph = False (ablation_gnns.py): the un-converted homogeneous SAGE model on both sides
            from .gnn_model import HOMOG_META

            self.emb_model = BaseGNN(input_dim, hidden_dim, hidden_dim, args, HOMOG_META, emb_channels=hidden_dim, **kwargs)
            self.emb_model_query = BaseGNN(input_dim, hidden_dim, hidden_dim, args, HOMOG_META, emb_channels=hidden_dim, **kwargs)
        self.count_model = nn.Sequential(nn.Linear(2 * hidden_dim, 4 * hidden_dim), nn.LeakyReLU(), nn.Linear(4 * hidden_dim, 1))
        # the reference collates DataLoader batches of args.batch_size neighborhoods (config.py:255): the default unit over
        # which SAGEConv's remove_self_loops quirk is evaluated (set_pyg_batch_size changes it)
        self.emb_model.pyg_batch_size = int(getattr(args, "batch_size", 0) or 0)
        self._init_cache()

    # ---- the reference converts with pyg.nn.to_hetero at run time; these modules are built hetero ----
    def to_hetero_old(self, tconv_target=False, tconv_query=False):
        if not getattr(self.args, "use_hetero", True):
            raise RuntimeError("a use_hetero=False model is not converted (main.py calls to_hetero only when args.use_hetero)")
        if not (tconv_target and tconv_query):
            raise NotImplementedError("only the default SHMP (use_tconv=True) metadata is a CUDA path")
        return self

    def to_hetero(self, order: int = 3, SHMP_target=False, SHMP_query=False):
        if order != 3:
            raise NotImplementedError("order-4 (11 relation) SHMP is not built")
        return self.to_hetero_old(SHMP_target, SHMP_query)

    def on_load_checkpoint(self, checkpoint) -> None:
        a = checkpoint["hyper_parameters"]["args"]
        if a.use_hetero and not (a.use_tconv and getattr(a, "use_canonical", True)):
            raise NotImplementedError("checkpoint was trained with hetero but without tconv/canonical: not a CUDA path")

    load_from_checkpoint = classmethod(_load_lightning_checkpoint)

    # ---- precision / batching knobs ----
    def set_precision(self, precision: str):
        self.emb_model.precision = precision
        self.emb_model_query.precision = "fp32"
        return self

    def set_pyg_batch_size(self, n: int):
        """Size of the collated batches the reference would form (``config.py:255``); only affects which bipartite
        edges SAGEConv's remove_self_loops drops (see gnn_model / DESIGN.md).  0: the whole input is one batch;
        negative: do not reproduce the quirk at all (no edge is dropped)."""
        self.emb_model.pyg_batch_size = int(n)
        return self

    # ---- queries ----
    def set_queries(self, query_ids, queries=None, transform=None, hetero=True, device=None):
        import networkx as nx

        dev = device if device is not None else self.count_model[0].weight.device
        batch, queries_nx = gen_queries(query_ids, queries, transform, self.input_dim, hetero, dev)
        min_len = max(nx.diameter(q) for q in queries_nx)
        if getattr(self, "depth", 4) < min_len:  # lightning_model.py:302-308
            warnings.warn("neighborhood diameter {:d} is too small for the queries, the minimum is {:d}".format(self.depth, min_len))
        self.query_loader = batch
        self._invalidate_caches()

    def get_query_emb(self) -> torch.Tensor:
        if self.query_loader is None:
            raise RuntimeError("call set_queries first")
        return self._cached("query_emb", self.emb_model_query, lambda: self.emb_model_query(self.query_loader))

    # ---- forward ----
    def graph_to_embed(self, batch) -> torch.Tensor:
        return self.emb_model(batch)

    def _head_weights(self):
        return self._cached("head", self.count_model, lambda: pack_head_weights(self.count_model, self.hidden_dim))

    def embed_to_count(self, embs, want_pred=False):
        """``lightning_model.py:176-193`` for ALL queries at once: embs = (emb_targets [G,h], emb_queries [Q,h])."""
        lib = _lib.load()
        emb_t, emb_q = embs
        emb_t, emb_q = emb_t.contiguous(), emb_q.contiguous()
        G, Q = emb_t.shape[0], emb_q.shape[0]
        dev = emb_t.device
        out = torch.empty((2 if want_pred else 1, G, Q), dtype=torch.float32, device=dev)
        wb = int(lib.desco_count_head_workspace_bytes(G, Q))
        work = self._scratch("head", max(wb, 1), dev)
        w = self._head_weights()
        precision = PRECISION[self.emb_model.precision]
        w_tc = self._cached("head_tc", self.count_model, lambda: pack_head_weights_tc(self.count_model, self.hidden_dim)) if precision else None
        status = getattr(self.emb_model, "last_status", None) if precision else None
        if precision and status is None:
            status = self.emb_model.last_status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.desco_count_head(_ptr(emb_t), G, _ptr(emb_q), Q, _ptr(w), _ptr(w_tc), self.hidden_dim,
                                            _ptr(out[1]) if want_pred else 0, _ptr(out[0]), _ptr(work), wb, precision,
                                            _ptr(status), _stream()), "desco_count_head")
        return (out[0], out[1]) if want_pred else out[0]

    def graph_to_count(self, batch) -> torch.Tensor:
        """``lightning_model.py:198-222``: [num_neighborhoods, num_queries] = 2**pred - 1."""
        return self.embed_to_count((self.emb_model(batch), self.get_query_emb()))

    def graph_to_pred(self, batch) -> torch.Tensor:
        """Pre-exponent of graph_to_count (the quantity the training loss sees, :249)."""
        return self.embed_to_count((self.emb_model(batch), self.get_query_emb()), want_pred=True)[1]

    def predict_step(self, batch, batch_idx=0) -> torch.Tensor:
        return self.graph_to_count(batch)

    def forward(self, batch):
        return self.graph_to_count(batch)

    # ---- training (csrc/train.cu through desco_b200/training.py) ----
    def train_forward(self, batch, batch_idx=0) -> torch.Tensor:
        """``lightning_model.py:228-254``: mean over the queries of ``smooth_l1(pred_q, log2(batch.y[:, q] + 1))``.
        The returned scalar carries a grad_fn over the target GNN, the query GNN and the count head."""
        from .training import train_forward

        return train_forward(self, batch)

    def training_step(self, batch, batch_idx=0) -> torch.Tensor:  # :133-136
        return self.train_forward(batch, batch_idx)

    def validation_step(self, batch, batch_idx=0) -> torch.Tensor:  # :147-154
        return self.train_forward(batch, batch_idx).detach()

    def criterion(self, count: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:  # :285-289
        return torch.nn.functional.smooth_l1_loss(count, truth)

    def configure_optimizers(self):
        """``lightning_model.py:160-173``: Adam(lr, weight_decay) + ReduceLROnPlateau(min, 0.5, patience 20, 1e-5)."""
        from .training import FusedAdam

        opt = FusedAdam(self.parameters(), lr=getattr(self, "lr", 1e-4), weight_decay=getattr(self, "weight_decay", 0.0))
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=20, min_lr=1e-5)
        return {"optimizer": opt, "lr_scheduler": sched, "monitor": "neighborhood_counting_val_loss"}


def default_gossip_args(**kw) -> SimpleNamespace:
    """``config.py:312-322``."""
    d = dict(conv_type="GOSSIP", layer_num=2, hidden_dim=64, dropout=0.01, use_hetero=False, lr=1e-3, weight_decay=0.0,
             batch_size=256)
    d.update(kw)
    return SimpleNamespace(**d)


class GossipCountingModel(nn.Module):
    """``lightning_model.py:535-649``."""

    def __init__(self, input_dim=1, hidden_dim=64, args=None, **kwargs):
        super().__init__()
        args = args or default_gossip_args(hidden_dim=hidden_dim)
        self.hidden_dim = hidden_dim
        kwargs["baseline"] = "gossip"
        kwargs.setdefault("emb_channels", 64)
        for k, v in vars(args).items():
            if not hasattr(self, k):
                setattr(self, k, v)
        self.emb_model = GossipBaseGNN(input_dim, hidden_dim, 1, args, **kwargs)
        self.kwargs = kwargs
        self.query_emb = None
        self.eval()

    def train(self, mode: bool = True):
        out = super().train(mode)
        self.emb_model._invalidate_caches()
        return out

    load_from_checkpoint = classmethod(_load_lightning_checkpoint)

    def set_query_emb(self, query_emb: torch.Tensor, query_ids=None, queries=None):
        self.query_emb = query_emb.detach()

    def graph_to_count(self, batch, query_emb=None) -> torch.Tensor:
        """``lightning_model.py:613-628``: batch carries the target CSR (``rowptr``, ``col``) and ``x`` [N, Q]."""
        qe = self.query_emb if query_emb is None else query_emb
        if qe is None:
            raise RuntimeError("call set_query_emb first")
        g = getattr(batch, "graph", batch)
        return self.emb_model.forward_all_queries(g.rowptr, g.col, batch.x, qe)

    def predict_step(self, batch, batch_idx=0) -> torch.Tensor:
        return self.graph_to_count(batch)

    def forward(self, batch):
        return self.graph_to_count(batch)

    # ---- training (csrc/gossip_train.cu through desco_b200/gossip_training.py) ----
    def train_forward(self, batch, batch_idx=0) -> torch.Tensor:
        """``lightning_model.py:585-608``: sum over queries and nodes of ``log2(|x + gossip(x) - y| + 1)``; the returned
        scalar carries a grad_fn over the gossip parameters (``pre_mp`` / ``anchor_mlp`` get none, as in the reference)."""
        from .gossip_training import gossip_train_forward

        return gossip_train_forward(self, batch)

    def training_step(self, batch, batch_idx=0) -> torch.Tensor:  # :553-556
        return self.train_forward(batch, batch_idx)

    def validation_step(self, batch, batch_idx=0) -> torch.Tensor:  # :562-564
        return self.train_forward(batch, batch_idx).detach()

    def test_step(self, batch, batch_idx=0) -> torch.Tensor:  # :558-560
        return self.train_forward(batch, batch_idx).detach()

    def criterion(self, count: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:  # :630-635
        return torch.log2(torch.abs(count - truth) + 1)

    def configure_optimizers(self):
        """``lightning_model.py:570-583``: Adam(lr, weight_decay) + ReduceLROnPlateau(min, 0.5, patience 20, 1e-5)."""
        opt = torch.optim.Adam(self.parameters(), lr=getattr(self, "lr", 1e-3), weight_decay=getattr(self, "weight_decay", 0.0))
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=20, min_lr=1e-5)
        return {"optimizer": opt, "lr_scheduler": sched, "monitor": "gossip_counting_val_loss"}

    def _gate_value(self, query_emb) -> torch.Tensor:
        """``lightning_model.py:640-649``: (#layers, #queries, 1)."""
        lib = _lib.load()
        w = self.emb_model.packed_weights()
        dev = w["wq"].device
        qe = query_emb.to(device=dev, dtype=torch.float32).contiguous()
        Q = qe.shape[0]
        qvec = torch.empty((Q, 256), dtype=torch.float32, device=dev)
        gates = torch.empty((2, Q), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.desco_gossip_prepare_queries(_ptr(qe), Q, _ptr(w["wq"]), _ptr(qvec), _ptr(gates), _stream()),
                       "desco_gossip_prepare_queries")
        return gates.unsqueeze(-1)
