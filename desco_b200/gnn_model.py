"""GNN modules of the DeSCo hot path - host side of csrc/shmp.cu and csrc/gossip.cu.

Mirrors ``subgraph_counting/gnn_model.py`` (reference @ 4508f7a): ``BaseGNN`` :18, ``BaseGNNCore`` :115, ``SAGEConv`` :362,
``GossipConv`` :280.  The modules own ``torch.nn`` parameters under the reference's state-dict key names (after
``to_hetero_old``: ``gnn_core.pre_mp.0.<type>``, ``gnn_core.convs.<l>.<src>__<rel>__<dst>.lin``,
``gnn_core.updates.<l>.<type>``, ``anchor_mlp.0``, ``post_mp.{0,3,5,7}``) so reference checkpoints
load unchanged; ``forward`` packs them into the fused K-major blobs include/desco_b200.h documents and runs the CUDA
kernels.  There is no eager/CPU forward: without the CUDA library these modules raise.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .data import NeighborhoodBatch, _ptr, _stream
from .tcpack import pack_b_operand, pack_dense_tc, pack_mma_b_frags

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

HOMOG_META = (["node"], [("node", "union", "node")])  # hetero_graph = False: one node type, one relation (ablation_gnns.py)

PRECISION = {"fp32": 0, "bf16x3": 1, "bf16": 2}
TILE_ROWS = 128  # csrc/shmp_internal.h SHMP_TILE_ROWS


def _key(et) -> str:
    return "__".join(et)


_WEIGHTS_EPOCH = [0]  # bumped by anything that rewrites parameter storage behind autograd's back (training.FusedAdam.step)


def bump_weights_epoch() -> None:
    """Invalidate every packed-weight / query-embedding cache in the process.  Call after mutating parameters through a
    path that neither bumps ``Parameter._version`` nor goes through ``load_state_dict`` / ``.to()`` / ``train()``."""
    _WEIGHTS_EPOCH[0] += 1


def _params_version(module: nn.Module) -> Tuple:
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


class _PackedWeightsMixin:
    """Cache of derived (packed / fused) weights.  Validating it exactly - data_ptr and in-place version of every
    parameter - costs ~100 us of Python per call, more than the kernels it guards.  In pure inference (``eval()`` and
    autograd disabled) parameters are therefore treated as frozen between the events that can change them behind our
    back: ``load_state_dict``, ``.to()/.cuda()/.float()`` and ``train()`` bump an epoch that invalidates the cache;
    in training mode or with autograd enabled the exact check runs every call."""

    def _init_cache(self):
        self._cache_epoch = 0
        self._caches = {}
        self._register_load_state_dict_pre_hook(lambda *a, **k: self._invalidate_caches())

    def _invalidate_caches(self):
        self._cache_epoch += 1

    def invalidate_caches(self):
        """Public form: drop the derived weights after weight surgery in ``eval()`` mode (``p.data.copy_``, ``p.normal_``,
        ``vector_to_parameters``, a torch optimizer step taken under ``no_grad`` + ``eval``): in pure inference the cache is
        NOT revalidated against the parameters on every call (see the class docstring)."""
        self._invalidate_caches()
        return self

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._invalidate_caches()
        return out

    def train(self, mode: bool = True):
        self._invalidate_caches()
        return super().train(mode)

    def _scratch(self, name: str, nbytes: int, device) -> torch.Tensor:
        """Grow-only scratch buffer owned by the module (kernel workspaces): consecutive launches on one stream may share
        it, and the step loses a caching-allocator round trip per call.  Not for concurrent use from several streams."""
        buf = self.__dict__.setdefault("_scratch_bufs", {}).get((name, str(device)))
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes * 1.25), 256), dtype=torch.uint8, device=device)
            self.__dict__["_scratch_bufs"][(name, str(device))] = buf
        return buf

    def _cached(self, name: str, module: nn.Module, build):
        exact = self.training or torch.is_grad_enabled()
        epoch = (self._cache_epoch, _WEIGHTS_EPOCH[0])
        hit = self._caches.get(name)
        if hit is not None and hit[0] == epoch and not exact:
            return hit[2]
        key = _params_version(module)
        if hit is None or hit[0] != epoch or hit[1] != key:
            hit = (epoch, key, build())
            self._caches[name] = hit
        return hit[2]


def _csr_by_target(edge_index: torch.Tensor, n_dst: int, extra: Optional[torch.Tensor] = None):
    """edge_index [2, E] (source, target) -> (rowptr[n_dst + 1], col[E] = sources, extra permuted alike), int32, on the
    edges' device.  Index plumbing only (sort + bincount); the arithmetic runs in csrc/conv.cu."""
    src, dst = edge_index[0].long(), edge_index[1].long()
    order = torch.argsort(dst, stable=True)
    rowptr = torch.zeros(n_dst + 1, dtype=torch.int64, device=edge_index.device)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n_dst), 0)
    return rowptr.to(torch.int32), src[order].to(torch.int32).contiguous(), (None if extra is None else extra[order].contiguous())


def _spmm_sum(rowptr, col, edge_w, x: torch.Tensor, n_dst: int) -> torch.Tensor:
    lib = _lib.load()
    x = x.to(torch.float32).contiguous()
    out = torch.empty((n_dst, x.shape[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.desco_spmm_sum(_ptr(rowptr), _ptr(col), _ptr(edge_w), n_dst, _ptr(x), x.stride(0), x.shape[1], _ptr(out),
                                      out.stride(0), _stream()), "desco_spmm_sum")
    return out


def _linear(xs: List[torch.Tensor], ws: List[torch.Tensor], bias: torch.Tensor) -> torch.Tensor:
    """sum_i xs[i] . ws[i]^T + bias through ``desco_train_dense`` (64-wide K blocks, output width a multiple of 64)."""
    from .training import _Ops

    n = ws[0].shape[0]
    if n % 64 or any(x.shape[1] % 64 for x in xs):
        raise NotImplementedError("the CUDA dense primitive needs channel counts that are multiples of 64")
    y = torch.empty((xs[0].shape[0], n), dtype=torch.float32, device=xs[0].device)
    if y.shape[0]:
        with torch.cuda.device(y.device):
            _Ops().dense([x.contiguous() for x in xs], [w.detach() for w in ws], [bias.detach()], y)
    return y


class SAGEConv(nn.Module):
    """``gnn_model.py:362-404``: ``lin(sum_{j->i} x_j)``.  Inside ``BaseGNN.forward`` the arithmetic runs fused for all
    relations and layers (csrc/shmp_fused.cu, shmp_mt.cu); called on its own it is ``forward(x | (x_src, x_dst),
    edge_index)`` of the reference on CUDA tensors: ``remove_self_loops`` (:389-390), sum-aggregation at the targets
    (csrc/conv.cu), then ``lin`` (csrc/train.cu dense).  Inference only (no autograd through the raw launches)."""

    def __init__(self, in_channels, out_channels, aggr="add", **kwargs):
        super().__init__()
        assert aggr == "add"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = nn.Linear(in_channels, out_channels)

    def reset_parameters(self):
        self.lin.reset_parameters()

    def forward(self, x, edge_index, edge_weight=None, size=None, res_n_id=None):
        x_src, x_dst = (x, x) if isinstance(x, torch.Tensor) else x
        if not x_src.is_cuda:
            raise RuntimeError("desco_b200 modules run on CUDA tensors only (no CPU fallback)")
        n_dst = x_dst.shape[0] if size is None else size[1]
        if edge_index is None:
            edge_index = torch.zeros((2, 0), dtype=torch.long, device=x_src.device)
        if edge_index.numel():
            edge_index = edge_index[:, edge_index[0] != edge_index[1]]  # pyg_utils.remove_self_loops
        rowptr, col, _ = _csr_by_target(edge_index, n_dst)
        return _linear([_spmm_sum(rowptr, col, None, x_src, n_dst)], [self.lin.weight], self.lin.bias)

    def __repr__(self):
        return "{}({}, {})".format(self.__class__.__name__, self.in_channels, self.out_channels)


class BaseGNNCore(nn.Module):
    """``gnn_model.py:115-277`` in its ``to_hetero`` form (node-level modules per node type, convs per edge type).
    Only the default SHMP configuration is a CUDA path: conv_type SAGE, hidden 64 (``config.py:247-264``)."""

    def __init__(self, input_dim, hidden_dim, output_dim, args, meta=TARGET_META, **kwargs):
        super().__init__()
        if args.conv_type != "SAGE":
            raise NotImplementedError("only the default SAGE SHMP core is built (GIN/GCN/GAT/PNA are ablations)")
        if hidden_dim != 64:
            raise NotImplementedError("the sm_100a kernels are specialised for hidden_dim = 64 (config.py:250)")
        self.meta = meta
        self.layer_num = args.layer_num
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.dropout = args.dropout
        self.pre_mp = nn.ModuleList([nn.ModuleDict({t: nn.Linear(input_dim, hidden_dim) for t in meta[0]})])
        self.convs = nn.ModuleList()
        self.updates = nn.ModuleList()
        for _ in range(args.layer_num):  # conv then update per layer: the reference's construction order (:146-190)
            self.convs.append(nn.ModuleDict({_key(et): SAGEConv(hidden_dim, hidden_dim) for et in meta[1]}))
            self.updates.append(nn.ModuleDict({t: nn.Linear(2 * hidden_dim, hidden_dim) for t in meta[0]}))
        self.post_input_dim = hidden_dim * args.layer_num + hidden_dim


class HomogGNNCore(nn.Module):
    """``BaseGNNCore`` (``gnn_model.py:115-277``) as the reference builds it BEFORE ``to_hetero``: the homogeneous SAGE model
    of ``hetero_graph=False`` (``workload.py:238-241``, ``ablation_gnns.py``) - one node type, one relation, the centre marked
    by ``node_feature = 1``.  State-dict keys are the reference's own (``pre_mp.0``, ``convs.<l>.lin``, ``updates.<l>``)."""

    def __init__(self, input_dim, hidden_dim, output_dim, args, **kwargs):
        super().__init__()
        if args.conv_type != "SAGE" or hidden_dim != 64:
            raise NotImplementedError("the sm_100a kernels serve the default SAGE core with hidden_dim = 64")
        self.meta = HOMOG_META
        self.layer_num, self.input_dim, self.hidden_dim, self.dropout = args.layer_num, input_dim, hidden_dim, args.dropout
        self.pre_mp = nn.Sequential(nn.Linear(input_dim, hidden_dim))
        self.convs = nn.ModuleList()
        self.updates = nn.ModuleList()
        for _ in range(args.layer_num):  # gnn_model.py:146-190: conv then update, layer by layer
            self.convs.append(SAGEConv(hidden_dim, hidden_dim))
            self.updates.append(nn.Linear(2 * hidden_dim, hidden_dim))
        self.post_input_dim = hidden_dim * args.layer_num + hidden_dim


def pack_shmp_weights(base: "BaseGNN") -> Dict[str, torch.Tensor]:
    """Fuse + transpose the parameters into the blobs of include/desco_b200.h (fp64 on the host, rounded once)."""
    core = base.gnn_core
    hetero = "canonical" in core.meta[0]
    homog = isinstance(core, HomogGNNCore)
    types = core.meta[0]
    F = core.hidden_dim
    dev = base.post_mp[0].weight.device
    d = lambda t: t.detach().to("cpu", torch.float64)
    pre = []
    for t in types:
        lin = core.pre_mp[0] if homog else core.pre_mp[0][t]
        pre += [d(lin.weight).t().contiguous().flatten(), d(lin.bias)]
    layers = []
    tc_layers = []
    mt_layers = []
    for l in range(core.layer_num):
        def fused(dst, rels):
            U, u = d(core.updates[l][dst].weight), d(core.updates[l][dst].bias)
            Um, Uh = U[:, :F], U[:, F:]
            blocks = [(Um @ d(core.convs[l][_key(r)].lin.weight)).t() for r in rels]
            return blocks, Um, Uh, u

        if hetero:
            cc = [("count", "union_triangle", "count"), ("count", "union_tride", "count")]
            ac = [("canonical", "union_triangle", "count"), ("canonical", "union_tride", "count")]
            ca = [("count", "union_triangle", "canonical"), ("count", "union_tride", "canonical")]
            blocks, Um, Uh, u = fused("count", cc)
            Wc = torch.cat(blocks + [Uh.t()], 0)
            bsum = sum(d(core.convs[l][_key(r)].lin.bias) for r in cc + ac)
            bias_c = Um @ bsum + u
            Cw = torch.cat([(Um @ d(core.convs[l][_key(r)].lin.weight)).t() for r in ac], 1)  # [F][2F]
            blocks_a, Uma, Uha, ua = fused("canonical", ca)
            Wa = torch.cat(blocks_a + [Uha.t()], 0)
            bias_a = Uma @ sum(d(core.convs[l][_key(r)].lin.bias) for r in ca) + ua
        elif homog:  # one relation: the triangle and the tride slots of the single-type kernels carry the SAME weights
            U, u = d(core.updates[l].weight), d(core.updates[l].bias)
            Um, Uh = U[:, :F], U[:, F:]
            blk = (Um @ d(core.convs[l].lin.weight)).t()
            Wc = torch.cat([blk, blk, Uh.t()], 0)
            bias_c = Um @ d(core.convs[l].lin.bias) + u
            Cw = torch.zeros(F, 2 * F, dtype=torch.float64)
            Wa = torch.zeros(3 * F, F, dtype=torch.float64)
            bias_a = torch.zeros(F, dtype=torch.float64)
        else:
            uu = [("union_node", "union_triangle", "union_node"), ("union_node", "union_tride", "union_node")]
            blocks, Um, Uh, u = fused("union_node", uu)
            Wc = torch.cat(blocks + [Uh.t()], 0)
            bias_c = Um @ sum(d(core.convs[l][_key(r)].lin.bias) for r in uu) + u
            Cw = torch.zeros(F, 2 * F, dtype=torch.float64)
            Wa = torch.zeros(3 * F, F, dtype=torch.float64)
            bias_a = torch.zeros(F, dtype=torch.float64)
        layers += [Wc.contiguous().flatten(), bias_c, Cw.contiguous().flatten(), Wa.contiguous().flatten(), bias_a]
        WcT = Wc.t()  # multi-tile form (csrc/shmp_mt.cu): per K block (tri | tride | self) the [64 n][64 k] block of Wc^T
        mt_layers += [pack_b_operand(WcT[:, kb * F:(kb + 1) * F]) for kb in range(3)]
        if hetero:  # tensor-core form (csrc/shmp_fused.cu): B operand rows n = output column, k contiguous
            B = torch.cat([Wc[0:F].t(), Wc[F:2 * F].t(), Wc[2 * F:3 * F].t()], 0)  # [192][64]
            f32b = lambda t: t.to(torch.float32).contiguous().view(torch.uint8).reshape(-1)
            tc_layers += [pack_b_operand(B), f32b(bias_c), f32b(bias_a), pack_mma_b_frags(Wa), pack_mma_b_frags(Cw)]
    ro = [d(base.anchor_mlp[0].weight).t().contiguous().flatten(), d(base.anchor_mlp[0].bias)]
    for i in (0, 3, 5, 7):
        ro += [d(base.post_mp[i].weight).t().contiguous().flatten(), d(base.post_mp[i].bias)]
    f32 = lambda parts: torch.cat(parts).to(torch.float32).to(dev).contiguous()
    out = {"pre": f32(pre), "layers": f32(layers), "readout": f32(ro), "layers_mt": torch.cat(mt_layers).to(dev).contiguous()}
    assert out["layers_mt"].numel() == core.layer_num * _lib.load().desco_shmp_mt_layer_bytes()
    if tc_layers or homog:
        out["readout_tc"] = torch.cat([pack_dense_tc(d(base.anchor_mlp[0].weight), 144), pack_dense_tc(d(base.post_mp[0].weight), 64),
                                       pack_dense_tc(d(base.post_mp[3].weight), 64), pack_dense_tc(d(base.post_mp[5].weight), 128),
                                       pack_dense_tc(d(base.post_mp[7].weight), 64)]).to(dev).contiguous()
    if tc_layers:
        out["layers_tc"] = torch.cat(tc_layers).to(dev).contiguous()
        assert out["layers_tc"].numel() == core.layer_num * _lib.load().desco_shmp_tc_layer_bytes()
    return out


class BaseGNN(_PackedWeightsMixin, nn.Module):
    """``gnn_model.py:18-109`` (SHMP path): ``forward(batch) -> [num_neighborhoods, output_dim]``."""

    def __init__(self, input_dim, hidden_dim, output_dim, args, meta=TARGET_META, **kwargs):
        super().__init__()
        if output_dim != hidden_dim:
            raise NotImplementedError("post_mp output width is fixed to hidden_dim on the CUDA path")
        self.args, self.kwargs = args, kwargs
        self.dropout, self.layer_num, self.conv_type = args.dropout, args.layer_num, args.conv_type
        self.use_hetero = getattr(args, "use_hetero", True)
        if meta is HOMOG_META or not self.use_hetero:
            self.gnn_core = HomogGNNCore(input_dim, hidden_dim, output_dim, args, **kwargs)
        else:
            self.gnn_core = BaseGNNCore(input_dim, hidden_dim, output_dim, args, meta, **kwargs)
        p = self.gnn_core.post_input_dim
        self.anchor_mlp = nn.Sequential(nn.Linear(p, p), nn.LeakyReLU(0.1))
        self.post_mp = nn.Sequential(
            nn.Linear(p, hidden_dim), nn.Dropout(args.dropout), nn.LeakyReLU(0.1), nn.Linear(hidden_dim, hidden_dim),
            nn.ReLU(), nn.Linear(hidden_dim, 256), nn.ReLU(), nn.Linear(256, output_dim),
        )
        self.precision = "bf16x3"  # tcgen05 path, ~3e-6 from the fp32 oracle; "fp32" = layer-by-layer FFMA kernels
        self.force_multi_tile = False  # tests: send small neighborhoods through the multi-tile tcgen05 kernels too
        self.pyg_batch_size = 0  # 0: the whole NeighborhoodBatch is one collated PyG batch
        self._init_cache()

    def packed_weights(self) -> Dict[str, torch.Tensor]:
        return self._cached("packed", self, lambda: pack_shmp_weights(self))

    def forward(self, data: NeighborhoodBatch, query_emb=None, feat: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not isinstance(data, NeighborhoodBatch):
            from .transforms import as_neighborhood_batch

            data = as_neighborhood_batch(data)
        if feat is None:
            feat = data._cache.get("feat")  # non-zero node features of a PyG-shaped input
        lib = _lib.load()
        core = self.gnn_core
        hetero = "canonical" in core.meta[0]
        homog = isinstance(core, HomogGNNCore)
        if hetero != data.hetero:
            raise ValueError("batch node-type layout does not match the model metadata (count/canonical vs union_node)")
        # homogeneous model: canonical-mode neighborhoods carry their centre as the LAST row, marked by node_feature = 1
        # (data.py:369-371) - that row goes through anchor_mlp (gnn_model.py:74-83); query graphs have no marked node
        anchored = homog and bool(data._cache.get("centre_last"))
        if anchored and feat is None:
            feat = torch.zeros((data.num_rows, core.input_dim), dtype=torch.float32, device=data.nbh_ptr.device)
            feat[(data.nbh_ptr[1:] - 1).long(), 0] = 1.0
        if self.training and core.dropout > 0:
            raise NotImplementedError("dropout > 0 in training mode is not a CUDA path (config.py:252 default is 0)")
        w = self.packed_weights()
        dev = w["pre"].device
        G, V = data.num_neighborhoods, data.num_rows
        out = torch.empty((G, core.hidden_dim), dtype=torch.float32, device=dev)
        if G == 0:
            return out
        wbytes = int(lib.desco_shmp_workspace_bytes(V, G, core.layer_num))
        work = self._scratch("shmp", wbytes, dev)
        if feat is not None:
            feat = feat.to(device=dev, dtype=torch.float32).contiguous()
            assert feat.shape == (V, core.input_dim)
        # tensor-core precisions: batches whose neighborhoods all fit a 128-row tile take the fused kernel (features never
        # leave the chip); anything larger takes the multi-tile kernels (features in HBM between layers, any size);
        # query graphs (single node type, a few dozen rows, embeddings cached) stay on the fp32 kernels
        precision = PRECISION[self.precision]
        if not hetero and not anchored and not self.force_multi_tile:
            precision = 0
        multi_tile = precision != 0 and (self.force_multi_tile or not hetero or data.max_rows > TILE_ROWS)
        status = torch.zeros(1, dtype=torch.int32, device=dev) if precision else None
        with torch.cuda.device(dev):
            # PyG-shaped input IS one collated batch: the remove_self_loops quirk is evaluated over it as a whole
            pyg_bs = int(self.pyg_batch_size) if not data._cache.get("pyg_collated") else min(int(self.pyg_batch_size), 0)
            common = (_ptr(data.nbh_ptr), _ptr(data.edge_ptr), _ptr(data.edge_col), _ptr(data.edge_tri), G, V,
                      2 if anchored else int(hetero), pyg_bs, _ptr(feat), core.input_dim, _ptr(w["pre"]), _ptr(w["layers"]))
            tail = (_ptr(w["readout"]), _ptr(w.get("readout_tc")), core.layer_num, core.hidden_dim, _ptr(out), _ptr(work),
                    wbytes, precision, _ptr(status), _stream())
            if multi_tile:
                _lib.check(lib.desco_shmp_forward_mt(*common, _ptr(w["layers_mt"]), *tail), "desco_shmp_forward_mt")
            else:
                _lib.check(lib.desco_shmp_forward(*common, _ptr(w.get("layers_tc")), *tail), "desco_shmp_forward")
        if status is not None:
            self.last_status = status  # device int32; 0 = ok.  Checked lazily (check_status) to keep the launch async.
        return out

    def check_status(self) -> None:
        """Raise if the last tensor-core forward reported a device-side error (one host sync)."""
        st = getattr(self, "last_status", None)
        if st is not None:
            _lib.check(int(st.item()), "desco_shmp_forward (device status)")


# ---------------------------------------------------------------------------------------------
# gossip
# ---------------------------------------------------------------------------------------------


class GossipConv(nn.Module):
    """``gnn_model.py:280-359``.  Inside ``GossipBaseGNN`` both layers, all queries and post_mp run fused (csrc/gossip.cu);
    called on its own it is the reference's ``forward(x, edge_index, edge_weight, query_emb)`` on CUDA tensors:
    ``lin_com`` per NODE (the reference applies it per edge, :341 - same sums), messages scaled by the gate or 1 - gate by
    edge direction (:342-343), summed at the targets, ``lin_update(cat(aggr, x))`` (:347-348).  Inference only."""

    def __init__(self, in_channels, out_channels, emb_channels, aggr="add", **kwargs):
        super().__init__()
        assert aggr == "add"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_com = nn.Linear(in_channels, out_channels)
        self.lin_update = nn.Linear(out_channels + in_channels, out_channels)
        self.lin_gate = nn.Sequential(
            nn.Linear(emb_channels, out_channels), nn.Sigmoid(), nn.Linear(out_channels, 1), nn.Sigmoid(), nn.LeakyReLU()
        )

    def __repr__(self):
        return "{}({}, {})".format(self.__class__.__name__, self.in_channels, self.out_channels)

    def _gate_value(self, query_emb: torch.Tensor) -> torch.Tensor:
        """``gnn_model.py:353-356``: ``lin_gate(query_emb)`` -> [Q, 1]."""
        lib = _lib.load()
        qe = query_emb.to(torch.float32).contiguous()
        if not qe.is_cuda:
            raise RuntimeError("desco_b200 modules run on CUDA tensors only (no CPU fallback)")
        l1, l2 = self.lin_gate[0], self.lin_gate[2]
        gate = torch.empty((qe.shape[0], 1), dtype=torch.float32, device=qe.device)
        with torch.cuda.device(qe.device):
            _lib.check(lib.desco_gossip_gate(_ptr(qe), qe.shape[0], qe.shape[1], _ptr(l1.weight.detach().contiguous()),
                                             _ptr(l1.bias.detach()), l1.weight.shape[0], _ptr(l2.weight.detach().contiguous()),
                                             _ptr(l2.bias.detach()), _ptr(gate), _stream()), "desco_gossip_gate")
        return gate

    def forward(self, x, edge_index, edge_weight=None, size=None, res_n_id=None, query_emb=None):
        if not x.is_cuda:
            raise RuntimeError("desco_b200 modules run on CUDA tensors only (no CPU fallback)")
        n = x.shape[0]
        edge_index = edge_index[:, edge_index[0] != edge_index[1]] if edge_weight is None else edge_index
        if edge_weight is None:  # :316-321: symmetrise + coalesce, direction flag = source < target
            key = torch.unique(torch.cat([edge_index[0] * n + edge_index[1], edge_index[1] * n + edge_index[0]]))
            edge_index = torch.stack([key // n, key % n])
            edge_weight = edge_index[0] < edge_index[1]
        else:
            keep = edge_index[0] != edge_index[1]
            edge_index, edge_weight = edge_index[:, keep], edge_weight[keep]
        gate = self._gate_value(query_emb).view(-1)[0] if query_emb is not None else torch.tensor(0.5, device=x.device)
        w = torch.where(edge_weight.bool(), gate, 1.0 - gate).to(torch.float32)
        rowptr, col, w = _csr_by_target(edge_index, n, w)
        y = _linear([x.to(torch.float32)], [self.lin_com.weight], self.lin_com.bias)  # lin_com(x_j), once per node
        aggr = _spmm_sum(rowptr, col, w, y, n)
        oc = self.out_channels
        return _linear([aggr, x.to(torch.float32)], [self.lin_update.weight[:, :oc], self.lin_update.weight[:, oc:]],
                       self.lin_update.bias)


class GossipCore(nn.Module):
    """GOSSIP configuration of ``BaseGNNCore`` (``gnn_model.py:131-134,146-183``): pre_mp + 2 GossipConv."""

    def __init__(self, input_dim, hidden_dim, args, emb_channels):
        super().__init__()
        if args.conv_type != "GOSSIP" or args.layer_num != 2 or hidden_dim != 64 or emb_channels != 64 or input_dim != 1:
            raise NotImplementedError("the gossip kernels are specialised for config.py:312-322 defaults "
                                      "(GOSSIP, 2 layers, hidden 64, query embedding 64, input_dim 1)")
        self.dropout = float(args.dropout)  # F.dropout after every conv in training mode (gnn_model.py:274; config.py:316: 0.01)
        self.pre_mp = nn.Sequential(nn.Linear(input_dim, hidden_dim))
        self.convs = nn.ModuleList()
        for l in range(args.layer_num):
            cin = hidden_dim + emb_channels if l == 0 else hidden_dim
            self.convs.append(GossipConv(cin, hidden_dim, emb_channels))
        self.post_input_dim = hidden_dim * args.layer_num + hidden_dim + emb_channels
        self.input_pattern_emb = True


def pack_gossip_weights(base: "GossipBaseGNN") -> Dict[str, torch.Tensor]:
    """Fold the rank-structured layer-0 / x0 terms (DESIGN.md "Gossip kernels") in fp64 and lay the rest out K-major."""
    core = base.gnn_core
    F = 64
    dev = base.post_mp[0].weight.device
    d = lambda t: t.detach().to("cpu", torch.float64)
    w_pre, b_pre = d(core.pre_mp[0].weight)[:, 0], d(core.pre_mp[0].bias)
    c0, c1 = core.convs[0], core.convs[1]
    Wcom0, bcom0 = d(c0.lin_com.weight), d(c0.lin_com.bias)
    Wcom0q, Wcom0c = Wcom0[:, :F], Wcom0[:, F:]
    Wup0, bup0 = d(c0.lin_update.weight), d(c0.lin_update.bias)
    Wup0a, Wup0q, Wup0c = Wup0[:, :F], Wup0[:, F:2 * F], Wup0[:, 2 * F:]
    Wcom1, bcom1 = d(c1.lin_com.weight), d(c1.lin_com.bias)
    Wup1, bup1 = d(c1.lin_update.weight), d(c1.lin_update.bias)
    Wup1a, Wup1b = Wup1[:, :F], Wup1[:, F:]
    P0, b0 = d(base.post_mp[0].weight), d(base.post_mp[0].bias)
    P0q, P0c, P0x1, P0x2 = P0[:, :F], P0[:, F:2 * F], P0[:, 2 * F:3 * F], P0[:, 3 * F:]
    pad3 = torch.zeros(3, dtype=torch.float64)
    wq = [(Wup0a @ Wcom0q).t().contiguous().flatten(), Wup0a @ (Wcom0c @ b_pre + bcom0),
          Wup0q.t().contiguous().flatten(), Wup0c @ b_pre + bup0,
          P0q.t().contiguous().flatten(), P0c @ b_pre + b0]
    for c in (c0, c1):
        wq += [d(c.lin_gate[0].weight).t().contiguous().flatten(), d(c.lin_gate[0].bias), d(c.lin_gate[2].weight)[0],
               d(c.lin_gate[2].bias), pad3]
    wg = [Wup0a @ (Wcom0c @ w_pre), Wup0c @ w_pre, P0c @ w_pre, Wup1a @ bcom1, bup1,
          torch.cat([(Wup1a @ Wcom1).t(), Wup1b.t()], 0).contiguous().flatten(),
          torch.cat([P0x1.t(), P0x2.t()], 0).contiguous().flatten(),
          d(base.post_mp[3].weight).t().contiguous().flatten(), d(base.post_mp[3].bias),
          d(base.post_mp[5].weight).t().contiguous().flatten(), d(base.post_mp[5].bias),
          d(base.post_mp[7].weight)[0], d(base.post_mp[7].bias), pad3]
    f32 = lambda parts: torch.cat(parts).to(torch.float32).to(dev).contiguous()
    # tensor-core operand images of the four GEMMs of the layer-1 / post_mp chain (csrc/gossip.cu IMG_*): B operand
    # W[n, k] per 64-wide K block, bf16 hi rows then lo rows, appended to the blob as raw bytes
    Wx2 = torch.cat([(Wup1a @ Wcom1).t(), Wup1b.t()], 0).t()  # [64 n][128 k]
    Wy1 = torch.cat([P0x1.t(), P0x2.t()], 0).t()
    images = torch.cat([pack_b_operand(Wx2[:, :F]), pack_b_operand(Wx2[:, F:]), pack_b_operand(Wy1[:, :F]),
                        pack_b_operand(Wy1[:, F:]), pack_b_operand(d(base.post_mp[3].weight)),
                        pack_b_operand(d(base.post_mp[5].weight))])
    out = {"wq": f32(wq), "wg": torch.cat([f32(wg), images.view(torch.float32).to(dev)]).contiguous()}
    lib = _lib.load()
    assert out["wq"].numel() == lib.desco_gossip_query_weight_floats(), (out["wq"].numel(), lib.desco_gossip_query_weight_floats())
    assert out["wg"].numel() == lib.desco_gossip_weight_floats(), (out["wg"].numel(), lib.desco_gossip_weight_floats())
    return out


class GossipBaseGNN(_PackedWeightsMixin, nn.Module):
    """``BaseGNN`` with ``baseline="gossip"`` (``gnn_model.py:18-109``): per-node output, no pooling, no anchor."""

    def __init__(self, input_dim, hidden_dim, output_dim, args, **kwargs):
        super().__init__()
        assert output_dim == 1
        self.args, self.kwargs = args, kwargs
        self.gnn_core = GossipCore(input_dim, hidden_dim, args, kwargs.get("emb_channels", 64))
        p = self.gnn_core.post_input_dim
        self.anchor_mlp = nn.Sequential(nn.Linear(p, p), nn.LeakyReLU(0.1))  # in the reference state dict, unused
        self.post_mp = nn.Sequential(
            nn.Linear(p, hidden_dim), nn.Dropout(args.dropout), nn.LeakyReLU(0.1), nn.Linear(hidden_dim, hidden_dim),
            nn.ReLU(), nn.Linear(hidden_dim, 256), nn.ReLU(), nn.Linear(256, output_dim),
        )
        self.precision = "bf16x3"  # tcgen05 layer-1 / post_mp chain (fp32-grade products); "fp32" = FFMA kernel
        self._init_cache()

    def packed_weights(self):
        return self._cached("packed", self, lambda: pack_gossip_weights(self))

    def forward_all_queries(self, rowptr, col, x, query_emb, want_gates=False):
        """out[N,Q] = x + gossip correction for every query column at once."""
        if self.training and self.gnn_core.dropout > 0:
            raise NotImplementedError("the fused gossip forward is the inference path (eval mode, lightning_model.py:613-628); "
                                      "training goes through GossipCountingModel.train_forward (desco_b200/gossip_training.py)")
        lib = _lib.load()
        w = self.packed_weights()
        dev = w["wg"].device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        query_emb = query_emb.to(device=dev, dtype=torch.float32).contiguous()
        N, Q = x.shape
        out = torch.empty_like(x)
        gates = torch.empty((2, Q), dtype=torch.float32, device=dev) if want_gates else None
        wb = int(lib.desco_gossip_workspace_bytes(N, Q))
        work = torch.empty(max(wb, 1), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.desco_gossip_forward(_ptr(rowptr), _ptr(col), N, _ptr(x), Q, _ptr(query_emb), _ptr(w["wg"]),
                                                _ptr(w["wq"]), _ptr(out), _ptr(gates), _ptr(work), wb,
                                                PRECISION[self.precision], _stream()), "desco_gossip_forward")
        return (out, gates) if want_gates else out


class GossipShardedRun:
    """One node-range-sharded gossip forward of one rank (the reference has no multi-GPU gossip,
    ``main.py:353-356``).  ``x[N, Q]`` is replicated (the hand-off all-gather of the counting stage); rank r owns the node
    rows ``[r * n_loc, (r + 1) * n_loc)``.

    ``start()``: per-query vectors, layer 0 of the own rows for every query (one launch, written in the query-grouped
    layout of ``desco_gossip_layer0_grouped``), then one in-place all-gather PER QUERY GROUP of the layer-0 scalars - the
    halo; with random labels on a power-law graph nearly every node is in some rank's halo, so the exchange is dense -
    issued asynchronously, back to back.
    ``finish()``: for each group, wait for ITS halo only, run layer 1 + post_mp of the own rows (gather + tcgen05 chain
    kernels), and hand the group's output rows to an asynchronous all-gather: the halo of group g+1 and the output of
    group g-1 move over NVLink while group g computes, so only the first halo group and the last output group are exposed.
    ``result()``: wait for the output all-gathers and return ``out[N, Q]`` (replicated) or, with
    ``gather_output=False``, the own rows ``[hi - lo, Q]``."""

    def __init__(self, model: "GossipBaseGNN", rowptr, col, x, query_emb, comm, query_group: int = 4, gather_output: bool = True):
        from .distributed import gossip_shard_plan

        if model.training and model.gnn_core.dropout > 0:
            raise NotImplementedError("the sharded gossip forward is the inference path; see gossip_training.py for training")
        self.m, self.rowptr, self.col, self.comm, self.gather_output = model, rowptr, col, comm, gather_output
        self.lib = _lib.load()
        self.w = model.packed_weights()
        self.dev = dev = self.w["wg"].device
        self.x = x.to(device=dev, dtype=torch.float32).contiguous()
        self.qe = query_emb.to(device=dev, dtype=torch.float32).contiguous()
        self.N, self.Q = self.x.shape
        self.plan = gossip_shard_plan(self.N, self.Q, comm.world, query_group)
        self.lo, self.hi = self.plan.ranges[comm.rank]
        self.halo_work: List = []

    def start(self):
        lib, w, dev, plan = self.lib, self.w, self.dev, self.plan
        N, Q = self.N, self.Q
        self.qvec = torch.empty((Q, 256), dtype=torch.float32, device=dev)
        self.s4 = torch.empty(plan.n_rows * Q * 4, dtype=torch.float32, device=dev)  # grouped [g][n_rows][qc][4]
        with torch.cuda.device(dev):
            st = _stream()
            _lib.check(lib.desco_gossip_prepare_queries(_ptr(self.qe), Q, _ptr(w["wq"]), _ptr(self.qvec), 0, st),
                       "desco_gossip_prepare_queries")
            if self.hi > self.lo:
                _lib.check(lib.desco_gossip_layer0_grouped(_ptr(self.rowptr), _ptr(self.col), self.lo, self.hi, _ptr(self.x), Q,
                                                           _ptr(self.qvec), _ptr(self.s4), plan.query_group, plan.n_rows, st),
                           "desco_gossip_layer0_grouped")
        self.s4_groups = [self.s4[4 * q0 * plan.n_rows: 4 * q1 * plan.n_rows].view(plan.n_rows, q1 - q0, 4) for q0, q1 in plan.groups]
        self.halo_work = [self.comm.all_gather_block(b, plan.n_loc, ("s4", gi)) for gi, b in enumerate(self.s4_groups)]
        return self

    def finish(self):
        lib, w, dev, plan = self.lib, self.w, self.dev, self.plan
        N, Q = self.N, self.Q
        prec = PRECISION[self.m.precision]
        n_own = self.hi - self.lo
        out = torch.empty(plan.n_rows * Q, dtype=torch.float32, device=dev)  # grouped [g][n_rows][qc]
        out_groups = [out[q0 * plan.n_rows: q1 * plan.n_rows].view(plan.n_rows, q1 - q0) for q0, q1 in plan.groups]
        sb = int(lib.desco_gossip_layer1_workspace_bytes(max(n_own, 1), plan.query_group, prec))
        stage = self.m._scratch("gossip_l1", max(sb, 1), dev)
        out_work = []
        with torch.cuda.device(dev):
            st = _stream()
            for (q0, q1), s4g, og, hw in zip(plan.groups, self.s4_groups, out_groups, self.halo_work):
                hw.wait()  # the current stream waits for this group's halo only
                if n_own:
                    _lib.check(lib.desco_gossip_layer1_group(_ptr(self.rowptr), _ptr(self.col), self.lo, self.hi, _ptr(s4g), q0,
                                                             q1 - q0, _ptr(self.qvec), _ptr(w["wg"]), _ptr(og), q1 - q0, prec,
                                                             _ptr(stage), sb, st), "desco_gossip_layer1_group")
                if self.gather_output:
                    out_work.append(self.comm.all_gather_block(og, plan.n_loc, ("out", q0)))
        self.out_groups, self.out_work = out_groups, out_work
        return self

    def result(self) -> torch.Tensor:
        if not self.gather_output:
            return torch.cat([og[self.lo:self.hi] for og in self.out_groups], dim=1)
        for wk in self.out_work:
            wk.wait()
        return torch.cat([og[:self.N] for og in self.out_groups], dim=1)


def _gossip_forward_sharded(self, rowptr, col, x, query_emb, comm, query_group: int = 4, gather_output: bool = True):
    return GossipShardedRun(self, rowptr, col, x, query_emb, comm, query_group, gather_output).start().finish().result()


GossipBaseGNN.forward_sharded = _gossip_forward_sharded
