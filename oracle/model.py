"""Oracle: SHMP neighborhood counting and gossip propagation in plain CPU torch.  TEST INFRASTRUCTURE ONLY.

Literal restatement (no fusion, no algebraic shortcuts) of

  * ``SAGEConv``                ``subgraph_counting/gnn_model.py:362-404``  sum-aggregate, then Linear(+bias)
  * ``BaseGNNCore.forward``     ``gnn_model.py:230-277`` as expanded per node / edge type by ``pyg.nn.to_hetero(aggr="sum")``
                                with the metadata of ``to_hetero_old`` (``lightning_model.py:371-421``)
  * ``BaseGNN.forward``         ``gnn_model.py:58-109``  anchor_mlp on canonical rows, global_add_pool, post_mp
  * ``NeighborhoodCountingModel.graph_to_count`` ``lightning_model.py:198-222`` per-query loop, ``2**p - 1``
  * ``GossipConv`` / GOSSIP branch ``gnn_model.py:231-260,280-359`` and ``GossipCountingModel.graph_to_count``
                                ``lightning_model.py:613-628`` per-query loop, per-EDGE ``lin_com``

Module / parameter names reproduce the reference's state-dict keys after ``to_hetero_old`` so the
same ``state_dict`` loads into the product modules in ``desco_b200``.

Inputs are the packed batch of ``oracle/partition.py`` (canonical node = last row of each neighborhood).  PyG
semantics relied on: ``edge_index[0]`` = source j, ``[1]`` = target i, ``out[i] = sum_e msg_e``;
relations with the same destination are combined by pairwise ``torch.add`` in metadata order; the bias of every
incoming relation reaches every destination row, also rows without such neighbours (``gnn_model.py:392-395``).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

TARGET_META = (
    ["count", "canonical"],
    [
        ("count", "union_triangle", "count"),
        ("count", "union_tride", "count"),
        ("count", "union_triangle", "canonical"),
        ("count", "union_tride", "canonical"),
        ("canonical", "union_triangle", "count"),
        ("canonical", "union_tride", "count"),
    ],
)  # lightning_model.py:376-385
QUERY_META = (
    ["union_node"],
    [("union_node", "union_triangle", "union_node"), ("union_node", "union_tride", "union_node")],
)  # lightning_model.py:404-413


def default_args(**kw) -> SimpleNamespace:
    """``config.py:247-264`` (neighborhood) defaults."""
    d = dict(conv_type="SAGE", layer_num=8, hidden_dim=64, input_dim=1, dropout=0.0, use_hetero=True, depth=4)
    d.update(kw)
    return SimpleNamespace(**d)


def default_gossip_args(**kw) -> SimpleNamespace:
    """``config.py:312-322`` (gossip) defaults."""
    d = dict(conv_type="GOSSIP", layer_num=2, hidden_dim=64, dropout=0.01, use_hetero=False)
    d.update(kw)
    return SimpleNamespace(**d)


def _key(et: Tuple[str, str, str]) -> str:
    return "__".join(et)


# ---------------------------------------------------------------------------------------------
# packed batch -> per-type views (what NetworkxToHetero + ToTconvHetero + collate would hold)
# ---------------------------------------------------------------------------------------------


def hetero_views(batch: Dict[str, np.ndarray], hetero: bool = True, pyg_batch_size: Optional[int] = None,
                 self_loop_quirk: bool = True):
    """Per node type: row ids into the packed row space + graph id per row.  Per relation: edge_index in per-type
    local ids (``transforms.py:342-367`` id assignment + ``:227-253`` type split).  ``hetero=False`` = query graphs:
    one node type ``union_node`` (``transforms.py:343-345``).

    REFERENCE QUIRK reproduced when ``self_loop_quirk``: ``SAGEConv.forward`` calls ``remove_self_loops`` on EVERY
    relation (``gnn_model.py:389-390``), including the bipartite count<->canonical ones whose two rows live in
    different id spaces, so an edge whose per-type local source id equals its per-type local destination id is
    silently dropped from message passing.  Ids are local to one collated PyG batch (``pyg_batch_size`` consecutive
    neighborhoods, ``config.py:255`` default 512; None = the whole input is one batch).  In the packed layout this
    hits neighborhood g of a batch exactly when every earlier neighborhood of that batch has 2 rows and the canonical
    node is adjacent to its lowest-id count node.  (Found by running the reference's own SAGEConv, see
    tests/golden/make_golden.py.)"""
    nbh_ptr = torch.as_tensor(batch["nbh_ptr"], dtype=torch.long)
    edge_ptr = torch.as_tensor(batch["edge_ptr"], dtype=torch.long)
    col = torch.as_tensor(batch["edge_col"], dtype=torch.long)
    tri = torch.as_tensor(batch["edge_tri"], dtype=torch.bool)
    V = int(nbh_ptr[-1])
    G = len(nbh_ptr) - 1
    graph_of_row = torch.repeat_interleave(torch.arange(G), nbh_ptr[1:] - nbh_ptr[:-1])
    dst = torch.repeat_interleave(torch.arange(V), edge_ptr[1:] - edge_ptr[:-1])  # row i receives from col j
    src = col
    if hetero:
        is_canon = torch.zeros(V, dtype=torch.bool)
        is_canon[nbh_ptr[1:] - 1] = True
        type_of = {"count": ~is_canon, "canonical": is_canon}
        meta = TARGET_META
    else:
        type_of = {"union_node": torch.ones(V, dtype=torch.bool)}
        meta = QUERY_META
    rows, local, batch_vec = {}, torch.zeros(V, dtype=torch.long), {}
    for t, m in type_of.items():
        rows[t] = torch.nonzero(m).flatten()
        local[rows[t]] = torch.arange(len(rows[t]))
        batch_vec[t] = graph_of_row[rows[t]]
    # per-type local id inside its own collated PyG batch (chunks of pyg_batch_size neighborhoods)
    bs = G if not pyg_batch_size else int(pyg_batch_size)
    chunk_of_row = graph_of_row // max(bs, 1)
    chunk_local = torch.zeros(V, dtype=torch.long)
    for t, m in type_of.items():
        r = rows[t]
        ch = chunk_of_row[r]
        first = torch.zeros(int(ch.max()) + 1 if len(ch) else 1, dtype=torch.long)
        if len(ch):
            first.scatter_reduce_(0, ch, torch.arange(len(r)), reduce="amin", include_self=False)
            chunk_local[r] = torch.arange(len(r)) - first[ch]
    edges = {}
    for (s, r, d) in meta[1]:
        want_tri = r.endswith("triangle")
        m = type_of[s][src] & type_of[d][dst] & (tri == want_tri)
        if self_loop_quirk and s != d:
            m = m & (chunk_local[src] != chunk_local[dst])  # remove_self_loops on a bipartite edge_index
        edges[(s, r, d)] = torch.stack([local[src[m]], local[dst[m]]])
    return SimpleNamespace(rows=rows, batch=batch_vec, edges=edges, num_graphs=G, meta=meta)


# ---------------------------------------------------------------------------------------------
# modules
# ---------------------------------------------------------------------------------------------


class SAGEConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = nn.Linear(in_channels, out_channels)

    def forward(self, x_src, n_dst: int, edge_index):
        agg = torch.zeros(n_dst, x_src.shape[1], dtype=x_src.dtype)
        if edge_index.numel():
            agg.index_add_(0, edge_index[1], x_src[edge_index[0]])  # propagate(aggr="add")  gnn_model.py:392
        return self.lin(agg)  # gnn_model.py:395


class HeteroSAGECore(nn.Module):
    """``BaseGNNCore`` (``gnn_model.py:115-277``) after ``to_hetero``: node-level modules per node type, convs per
    edge type."""

    def __init__(self, input_dim, hidden_dim, args, meta):
        super().__init__()
        self.meta = meta
        self.layer_num = args.layer_num
        self.pre_mp = nn.ModuleList([nn.ModuleDict({t: nn.Linear(input_dim, hidden_dim) for t in meta[0]})])
        self.convs = nn.ModuleList()
        self.updates = nn.ModuleList()
        for _ in range(args.layer_num):  # conv then update per layer: the reference's construction (= RNG) order, :146-190
            self.convs.append(nn.ModuleDict({_key(et): SAGEConv(hidden_dim, hidden_dim) for et in meta[1]}))
            self.updates.append(nn.ModuleDict({t: nn.Linear(2 * hidden_dim, hidden_dim) for t in meta[0]}))
        self.post_input_dim = hidden_dim * args.layer_num + hidden_dim

    def forward(self, x_dict, edge_index_dict):
        x = {t: self.pre_mp[0][t](x_dict[t]) for t in self.meta[0]}  # gnn_model.py:231
        emb = dict(x)
        for l in range(self.layer_num):
            per_dst: Dict[str, List[torch.Tensor]] = {t: [] for t in self.meta[0]}
            for et in self.meta[1]:
                s, _, d = et
                per_dst[d].append(self.convs[l][_key(et)](x[s], x[d].shape[0], edge_index_dict[et]))
            new_x = {}
            for t in self.meta[0]:
                vals = per_dst[t]
                while len(vals) > 1:  # pairwise tree add in metadata order (to_hetero aggr="sum")
                    nxt = [vals[i] + vals[i + 1] for i in range(0, len(vals) - 1, 2)]
                    if len(vals) % 2:
                        nxt.append(vals[-1])
                    vals = nxt
                h = self.updates[l][t](torch.cat((vals[0], x[t]), dim=1))  # gnn_model.py:264
                new_x[t] = F.relu(h)  # :273 ; dropout p=0 (:274, F11)
            x = new_x
            emb = {t: torch.cat((emb[t], x[t]), 1) for t in self.meta[0]}  # :275
        return emb


class BaseGNN(nn.Module):
    """``gnn_model.py:18-109`` hetero SAGE path."""

    def __init__(self, input_dim, hidden_dim, output_dim, args, meta=TARGET_META):
        super().__init__()
        self.gnn_core = HeteroSAGECore(input_dim, hidden_dim, args, meta)
        p = self.gnn_core.post_input_dim
        self.anchor_mlp = nn.Sequential(nn.Linear(p, p), nn.LeakyReLU(0.1))
        self.post_mp = nn.Sequential(
            nn.Linear(p, hidden_dim), nn.Dropout(args.dropout), nn.LeakyReLU(0.1), nn.Linear(hidden_dim, hidden_dim),
            nn.ReLU(), nn.Linear(hidden_dim, 256), nn.ReLU(), nn.Linear(256, output_dim),
        )
        self.input_dim = input_dim

    def forward(self, views, feat: Optional[Dict[str, torch.Tensor]] = None):
        dt = self.post_mp[0].weight.dtype
        if feat is None:  # ZeroNodeFeat  (workload.py:431-440, F9)
            feat = {t: torch.zeros(len(r), self.input_dim, dtype=dt) for t, r in views.rows.items()}
        emb = self.gnn_core(feat, views.edges)
        if "canonical" in emb:
            emb["canonical"] = self.anchor_mlp(emb["canonical"])  # gnn_model.py:69-73
        pooled = torch.zeros(views.num_graphs, self.gnn_core.post_input_dim, dtype=dt)
        for t in emb:  # to_homogeneous + cat + global_add_pool with a CONSISTENT type order (F10)
            pooled.index_add_(0, views.batch[t], emb[t])
        return self.post_mp(pooled)  # gnn_model.py:107-108


class NeighborhoodCountingModel(nn.Module):
    """``lightning_model.py:90-222`` (inference surface)."""

    def __init__(self, input_dim=1, hidden_dim=64, args=None):
        super().__init__()
        args = args or default_args(hidden_dim=hidden_dim, input_dim=input_dim)
        self.emb_model = BaseGNN(input_dim, hidden_dim, hidden_dim, args, TARGET_META)
        self.emb_model_query = BaseGNN(input_dim, hidden_dim, hidden_dim, args, QUERY_META)
        self.count_model = nn.Sequential(
            nn.Linear(2 * hidden_dim, 4 * hidden_dim), nn.LeakyReLU(), nn.Linear(4 * hidden_dim, 1)
        )  # lightning_model.py:127-131

    def get_query_emb(self, query_batch):
        return self.emb_model_query(hetero_views(query_batch, hetero=False))  # :311-316

    def graph_to_embed(self, batch, pyg_batch_size=None, self_loop_quirk=True):
        return self.emb_model(hetero_views(batch, True, pyg_batch_size, self_loop_quirk))

    def pre_exponent(self, batch, query_batch, pyg_batch_size=None, self_loop_quirk=True):
        emb_q = self.get_query_emb(query_batch)
        emb_t = self.graph_to_embed(batch, pyg_batch_size, self_loop_quirk)
        out = []
        for q in emb_q:  # lightning_model.py:212-219
            out.append(self.count_model(torch.cat((emb_t, q.expand_as(emb_t)), dim=-1)))
        return torch.cat(out, dim=-1)

    def graph_to_count(self, batch, query_batch, pyg_batch_size=None, self_loop_quirk=True):
        return 2 ** self.pre_exponent(batch, query_batch, pyg_batch_size, self_loop_quirk) - 1  # :221

    def train_forward(self, batch, query_batch, y, pyg_batch_size=None, self_loop_quirk=True):
        """``lightning_model.py:228-254``: per query ``criterion(results, log2(truth + 1))`` with
        ``criterion = F.smooth_l1_loss`` (:285-289), then the mean of the stacked per-query losses."""
        emb_q = self.get_query_emb(query_batch)
        emb_t = self.graph_to_embed(batch, pyg_batch_size, self_loop_quirk)
        losses = []
        for i, q in enumerate(emb_q):
            results = self.count_model(torch.cat((emb_t, q.expand_as(emb_t)), dim=-1))
            truth = y[:, i].view(-1, 1)
            losses.append(F.smooth_l1_loss(results, torch.log2(truth + 1)))
        return torch.mean(torch.stack(losses))


# ---------------------------------------------------------------------------------------------
# gossip
# ---------------------------------------------------------------------------------------------


class GossipConv(nn.Module):
    """``gnn_model.py:280-359``."""

    def __init__(self, in_channels, out_channels, emb_channels):
        super().__init__()
        self.lin_com = nn.Linear(in_channels, out_channels)
        self.lin_update = nn.Linear(out_channels + in_channels, out_channels)
        self.lin_gate = nn.Sequential(
            nn.Linear(emb_channels, out_channels), nn.Sigmoid(), nn.Linear(out_channels, 1), nn.Sigmoid(), nn.LeakyReLU()
        )

    def forward(self, x, edge_index, edge_weight, query_emb):
        gate = self.lin_gate(query_emb)  # [1,1]
        msg = self.lin_com(x[edge_index[0]])  # per EDGE  (gnn_model.py:341)
        msg = torch.where(edge_weight.view(-1, 1), msg * gate, msg * (1 - gate))  # :342-343
        aggr = torch.zeros(x.shape[0], msg.shape[1], dtype=x.dtype)
        aggr.index_add_(0, edge_index[1], msg)
        return self.lin_update(torch.cat([aggr, x], dim=-1))  # :347-348


class GossipCore(nn.Module):
    def __init__(self, input_dim, hidden_dim, args, emb_channels):
        super().__init__()
        self.pre_mp = nn.Sequential(nn.Linear(input_dim, hidden_dim))
        self.convs = nn.ModuleList()
        for l in range(args.layer_num):
            cin = hidden_dim + emb_channels if l == 0 else hidden_dim  # gnn_model.py:147-153
            self.convs.append(GossipConv(cin, hidden_dim, emb_channels))
        self.post_input_dim = hidden_dim * args.layer_num + hidden_dim + emb_channels  # :207 with :133-134

    def forward(self, x, edge_index, query_emb):
        x = self.pre_mp(x)
        x = torch.cat((query_emb.expand(x.shape[0], -1), x), dim=-1).clone().detach()  # :236-240
        edge_weight = edge_index[0] < edge_index[1]  # :248 (inputs are already simple + symmetric)
        emb = x
        for conv in self.convs:
            x = F.relu(conv(x, edge_index, edge_weight, query_emb))  # :258-260,273 ; dropout off in eval
            emb = torch.cat((emb, x), 1)
        return emb


class GossipBaseGNN(nn.Module):
    def __init__(self, input_dim, hidden_dim, args, emb_channels):
        super().__init__()
        self.gnn_core = GossipCore(input_dim, hidden_dim, args, emb_channels)
        p = self.gnn_core.post_input_dim
        self.anchor_mlp = nn.Sequential(nn.Linear(p, p), nn.LeakyReLU(0.1))  # present in the state dict, unused
        self.post_mp = nn.Sequential(
            nn.Linear(p, hidden_dim), nn.Dropout(args.dropout), nn.LeakyReLU(0.1), nn.Linear(hidden_dim, hidden_dim),
            nn.ReLU(), nn.Linear(hidden_dim, 256), nn.ReLU(), nn.Linear(256, 1),
        )

    def forward(self, node_feature, edge_index, query_emb):
        return self.post_mp(self.gnn_core(node_feature, edge_index, query_emb))  # gnn_model.py:102-103


class GossipCountingModel(nn.Module):
    """``lightning_model.py:535-649`` (inference surface)."""

    def __init__(self, input_dim=1, hidden_dim=64, args=None, emb_channels=64):
        super().__init__()
        args = args or default_gossip_args(hidden_dim=hidden_dim)
        self.emb_model = GossipBaseGNN(input_dim, hidden_dim, args, emb_channels)
        self.query_emb = None
        self.eval()

    def set_query_emb(self, query_emb):
        self.query_emb = query_emb.detach()

    def graph_to_count(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        out = []
        for q in range(self.query_emb.shape[0]):  # :615-625
            c = x[:, q].view(-1, 1)
            out.append(c + self.emb_model(c, edge_index, self.query_emb[q].view(1, -1)))
        return torch.cat(out, dim=-1)

    def train_forward(self, x: torch.Tensor, y: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        """``lightning_model.py:585-608`` with ``criterion`` :630-635: sum over queries and nodes of
        ``log2(|x_q + gossip(x_q) - y_q| + 1)``."""
        losses = []
        for q in range(x.shape[1]):
            c = x[:, q].view(-1, 1)
            pred = c + self.emb_model(c, edge_index, self.query_emb[q].view(1, -1))
            losses.append(torch.log2(torch.abs(pred - y[:, q].view(-1, 1)) + 1))
        return torch.sum(torch.stack(losses))

    def gate_value(self, query_emb):
        return torch.stack([c.lin_gate(query_emb) for c in self.emb_model.gnn_core.convs], dim=0)  # :640-649


# ---------------------------------------------------------------------------------------------
# queries
# ---------------------------------------------------------------------------------------------

STANDARD_QUERY_IDS = [6, 7, 13, 14, 15, 16, 17, 18, 29, 30, 31, 34, 35, 36, 37, 38, 40, 41, 42, 43, 44, 45, 46, 47,
                      48, 49, 50, 51, 52]  # gen_query_ids([3,4,5])  data.py:37-58


def gen_query_ids(query_size=(3, 4, 5)) -> List[int]:
    """``data.py:37-58``: connected atlas graphs of the requested sizes."""
    import networkx as nx

    out = []
    for i in range(6, 209):
        g = nx.graph_atlas(i)
        if len(g) > max(query_size):
            break
        if nx.is_connected(g) and len(g) in query_size:
            out.append(i)
    return out


def query_batch(query_ids=None) -> Dict[str, np.ndarray]:
    """Atlas queries packed like a neighborhood batch (one 'neighborhood' per query, single node type), typed by the
    same SHMP rule (``lightning_model.py:84-85``)."""
    import networkx as nx
    from oracle.partition import edge_is_triangle

    query_ids = STANDARD_QUERY_IDS if query_ids is None else query_ids
    nbh_ptr, edge_ptr, edge_col, edge_tri = [0], [0], [], []
    for qid in query_ids:
        g = nx.graph_atlas(qid)
        nodes = sorted(g.nodes)
        row0 = nbh_ptr[-1]
        pos = {u: row0 + i for i, u in enumerate(nodes)}
        for u in nodes:
            for v in sorted(g.neighbors(u)):
                edge_col.append(pos[v])
                edge_tri.append(1 if edge_is_triangle(g, u, v) else 0)
            edge_ptr.append(len(edge_col))
        nbh_ptr.append(row0 + len(nodes))
    return {
        "nbh_ptr": np.asarray(nbh_ptr, np.int32), "edge_ptr": np.asarray(edge_ptr, np.int32),
        "edge_col": np.asarray(edge_col, np.int32), "edge_tri": np.asarray(edge_tri, np.uint8),
    }
