"""Training step of the neighborhood-counting model - host side of csrc/train.cu.

Mirrors ``NeighborhoodCountingModel.train_forward`` (``subgraph_counting/lightning_model.py:228-254``: per-query
``smooth_l1(pred_q, log2(y_q + 1))``, mean over the queries), ``criterion`` (:285-289) and ``configure_optimizers``
(:160-173, ``torch.optim.Adam`` + ``ReduceLROnPlateau``).  The reference lets autograd record the PyG graph; here the
backward pass is written out: every O(rows) / O(edges) operation is a CUDA kernel of csrc/train.cu called through the C
ABI, PyTorch only owns the buffers.  The result is exposed as ONE ``torch.autograd.Function`` whose inputs are the
model parameters, so ``loss.backward()`` and any optimizer keep working; ``FusedAdam`` is the one-launch Adam.

No CPU fallback: everything raises without the CUDA library.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .data import NeighborhoodBatch, _ptr, _stream
from .gnn_model import BaseGNN, _key

F = 64
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2

_VPA = {}
_I32A = {}


def _vp_array(ptrs: Sequence[int]):
    n = len(ptrs)
    t = _VPA.get(n)
    if t is None:
        t = _VPA[n] = ctypes.c_void_p * n
    return t(*ptrs)


def _i32_array(vals: Sequence[int]):
    n = len(vals)
    t = _I32A.get(n)
    if t is None:
        t = _I32A[n] = ctypes.c_int32 * n
    return t(*vals)


def _check2d(t: torch.Tensor):
    assert t.dim() == 2 and t.dtype == torch.float32 and t.is_cuda and (t.shape[1] == 0 or t.stride(1) == 1), (t.shape, t.stride())


class _Ops:
    """Thin typed wrappers over the desco_train_* entry points (all launches go to the current stream)."""

    def __init__(self):
        self.lib = _lib.load()
        self.st = _stream()

    # Y (+)= act(sum_i X_i . W_i^T + sum bias);  X_i [M, K_i], W_i [N, K_i] (torch Linear weight or a column slice)
    def dense(self, xs: List[torch.Tensor], ws: List[torch.Tensor], biases: List[torch.Tensor], y: torch.Tensor,
              act: int = ACT_NONE, slope: float = 0.0, accumulate: bool = False):
        xp, xl, wp, wl = [], [], [], []
        M, N = y.shape
        for x, w in zip(xs, ws):
            _check2d(x); _check2d(w)
            K = x.shape[1]
            assert x.shape[0] == M and w.shape == (N, K) and K % F == 0, (x.shape, w.shape, y.shape)
            for b in range(K // F):
                xp.append(x.data_ptr() + 4 * F * b); xl.append(x.stride(0))
                wp.append(w.data_ptr() + 4 * F * b); wl.append(w.stride(0))
        _check2d(y)
        bp = [b.data_ptr() for b in biases]
        _lib.check(self.lib.desco_train_dense(_vp_array(xp), _i32_array(xl), _vp_array(wp), _i32_array(wl), len(xp), 1,
                                              _vp_array(bp) if bp else None, len(bp), _ptr(y), y.stride(0), M, N, act,
                                              slope, int(accumulate), self.st), "desco_train_dense")

    # dX (+)= dY . W;  dY [M, N], W [N, K] -> dX [M, K]
    def dgrad(self, dy: torch.Tensor, w: torch.Tensor, dx: torch.Tensor, accumulate: bool = False):
        _check2d(dy); _check2d(w); _check2d(dx)
        M, N = dy.shape
        K = w.shape[1]
        assert w.shape[0] == N and dx.shape == (M, K) and N % F == 0 and K % F == 0, (dy.shape, w.shape, dx.shape)
        nb = N // F
        xp = [dy.data_ptr() + 4 * F * b for b in range(nb)]
        wp = [w.data_ptr() + 4 * F * b * w.stride(0) for b in range(nb)]
        _lib.check(self.lib.desco_train_dense(_vp_array(xp), _i32_array([dy.stride(0)] * nb), _vp_array(wp),
                                              _i32_array([w.stride(0)] * nb), nb, 0, None, 0, _ptr(dx), dx.stride(0), M, K,
                                              ACT_NONE, 0.0, int(accumulate), self.st), "desco_train_dense (dgrad)")

    # dW_i += dY^T X[:, 64 i : 64 i + 64] for the list of [N, 64] gradient views dws;  db_j += colsum(dY)
    def wgrad(self, x: torch.Tensor, dy: torch.Tensor, dws: List[torch.Tensor], dbs: List[torch.Tensor] = ()):
        _check2d(x); _check2d(dy)
        M, N = dy.shape
        assert x.shape[0] == M and x.shape[1] == F * len(dws) and N % F == 0
        for d in dws:
            _check2d(d)
            assert d.shape == (N, F)
        dp = [d.data_ptr() for d in dws]
        bp = [b.data_ptr() for b in dbs]
        _lib.check(self.lib.desco_train_wgrad(_ptr(x), x.stride(0), len(dws), _ptr(dy), dy.stride(0), N, M, _vp_array(dp),
                                              _i32_array([d.stride(0) for d in dws]), _vp_array(bp) if bp else None,
                                              len(bp), self.st), "desco_train_wgrad")

    def act_backward(self, dx: torch.Tensor, fwd: torch.Tensor, act: int, slope: float = 0.0):
        _check2d(dx); _check2d(fwd)
        assert dx.shape == fwd.shape
        _lib.check(self.lib.desco_train_act_backward(_ptr(dx), dx.stride(0), _ptr(fwd), fwd.stride(0), dx.shape[0],
                                                     dx.shape[1], act, slope, self.st), "desco_train_act_backward")

    def fill_rows(self, y: torch.Tensor, bias: torch.Tensor):
        _check2d(y)
        _lib.check(self.lib.desco_train_fill_rows(_ptr(y), y.stride(0), y.shape[0], y.shape[1], _ptr(bias), self.st),
                   "desco_train_fill_rows")

    def colsum(self, dy: torch.Tensor, db: torch.Tensor):
        _check2d(dy)
        _lib.check(self.lib.desco_train_colsum(_ptr(dy), dy.stride(0), dy.shape[0], dy.shape[1], _ptr(db), self.st),
                   "desco_train_colsum")


def _split_cols(w: torch.Tensor) -> List[torch.Tensor]:
    """[N, 64 k] -> k views [N, 64]."""
    return [w[:, F * i:F * (i + 1)] for i in range(w.shape[1] // F)]


class _ShmpTape:
    """Forward of one ``BaseGNN`` (``gnn_model.py:58-109`` over ``:230-277``) with everything the backward needs."""

    def __init__(self, base: BaseGNN, batch: NeighborhoodBatch, ops: _Ops):
        core = base.gnn_core
        if core.dropout > 0 and base.training:
            raise NotImplementedError("dropout > 0 in training is not a CUDA path (config.py:252 default 0.0)")
        self.base, self.batch, self.ops = base, batch, ops
        self.hetero = "canonical" in core.meta[0]
        if self.hetero != batch.hetero:
            raise ValueError("batch node-type layout does not match the model metadata")
        self.L = core.layer_num
        self.G, self.V = batch.num_neighborhoods, batch.num_rows
        self.Vc = self.V - self.G if self.hetero else self.V
        self.ct = "count" if self.hetero else "union_node"
        # relation order = slot order of desco_train_aggregate: (source type, tri/tride)
        if self.hetero:
            self.rel_c = [("count", "union_triangle", "count"), ("count", "union_tride", "count"),
                          ("canonical", "union_triangle", "count"), ("canonical", "union_tride", "count")]
            self.rel_a = [("count", "union_triangle", "canonical"), ("count", "union_tride", "canonical")]
        else:
            self.rel_c = [("union_node", "union_triangle", "union_node"), ("union_node", "union_tride", "union_node")]
            self.rel_a = []

    def _aggregate(self, l: int, transpose: bool, xc, xa, ac, aa):
        b, lib = self.batch, self.ops.lib
        _lib.check(lib.desco_train_aggregate(
            _ptr(b.nbh_ptr), _ptr(b.edge_ptr), _ptr(b.edge_col), _ptr(b.edge_tri), _ptr(self.row_nbh), _ptr(self.quirk),
            self.V, int(self.hetero), int(transpose), _ptr(xc), xc.stride(0), _ptr(xa) if xa is not None else 0,
            xa.stride(0) if xa is not None else 0, _ptr(ac), ac.stride(0), _ptr(aa) if aa is not None else 0,
            aa.stride(0) if aa is not None else 0, self.ops.st), "desco_train_aggregate")

    def forward(self) -> torch.Tensor:
        base, ops, core = self.base, self.ops, self.base.gnn_core
        b, lib = self.batch, ops.lib
        dev = base.post_mp[0].weight.device
        G, V, Vc, L, het = self.G, self.V, self.Vc, self.L, self.hetero
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        E = (L + 1) * F
        self.row_nbh = torch.empty(max(V, 1), **i32)
        self.crow_nbh = torch.empty(max(Vc, 1), **i32)
        self.quirk = torch.empty(max(G, 1), **i32)
        _lib.check(lib.desco_train_plan(_ptr(b.nbh_ptr), G, int(het), int(base.pyg_batch_size), _ptr(self.row_nbh),
                                        _ptr(self.crow_nbh), _ptr(self.quirk), ops.st), "desco_train_plan")
        S = len(self.rel_c)
        self.emb_c = torch.empty((Vc, E), **f32)
        self.Ac = torch.empty((L, Vc, S * F), **f32)
        self.Mc = torch.empty((L, Vc, F), **f32)
        ops.fill_rows(self.emb_c[:, :F], core.pre_mp[0][self.ct].bias)  # ZeroNodeFeat: h0 = bias (gnn_model.py:231)
        if het:
            self.emb_a = torch.empty((G, E), **f32)
            self.Aa = torch.empty((L, G, 2 * F), **f32)
            self.Ma = torch.empty((L, G, F), **f32)
            ops.fill_rows(self.emb_a[:, :F], core.pre_mp[0]["canonical"].bias)
        else:
            self.emb_a = self.Aa = self.Ma = None
        for l in range(L):
            hc = self.emb_c[:, F * l:F * (l + 1)]
            ha = self.emb_a[:, F * l:F * (l + 1)] if het else None
            self._aggregate(l, False, hc, ha, self.Ac[l], self.Aa[l] if het else None)
            convs = [core.convs[l][_key(r)].lin for r in self.rel_c]
            ops.dense(_split_cols(self.Ac[l]), [c.weight for c in convs], [c.bias for c in convs], self.Mc[l])
            U = core.updates[l][self.ct]
            ops.dense([self.Mc[l], hc], [U.weight[:, :F], U.weight[:, F:]], [U.bias], self.emb_c[:, F * (l + 1):F * (l + 2)],
                      ACT_RELU)  # gnn_model.py:264,273
            if het:
                convs = [core.convs[l][_key(r)].lin for r in self.rel_a]
                ops.dense(_split_cols(self.Aa[l]), [c.weight for c in convs], [c.bias for c in convs], self.Ma[l])
                U = core.updates[l]["canonical"]
                ops.dense([self.Ma[l], ha], [U.weight[:, :F], U.weight[:, F:]], [U.bias],
                          self.emb_a[:, F * (l + 1):F * (l + 2)], ACT_RELU)
        self.z_a = None
        if het:  # anchor_mlp on the canonical rows (gnn_model.py:69-73)
            self.z_a = torch.empty((G, E), **f32)
            A = base.anchor_mlp[0]
            ops.dense([self.emb_a], [A.weight], [A.bias], self.z_a, ACT_LEAKY, 0.1)
        self.pooled = torch.empty((G, E), **f32)  # global_add_pool over both node types (gnn_model.py:88-89,107)
        _lib.check(lib.desco_train_pool(_ptr(b.nbh_ptr), _ptr(self.crow_nbh), G, Vc, int(het), E, 0, _ptr(self.emb_c),
                                        self.emb_c.stride(0), _ptr(self.z_a) if het else 0, E if het else 0,
                                        _ptr(self.pooled), E, ops.st), "desco_train_pool")
        P = base.post_mp
        self.t1 = torch.empty((G, F), **f32)
        self.t2 = torch.empty((G, F), **f32)
        self.t3 = torch.empty((G, 4 * F), **f32)
        out = torch.empty((G, F), **f32)
        ops.dense([self.pooled], [P[0].weight], [P[0].bias], self.t1, ACT_LEAKY, 0.1)  # gnn_model.py:44-53
        ops.dense([self.t1], [P[3].weight], [P[3].bias], self.t2, ACT_RELU)
        ops.dense([self.t2], [P[5].weight], [P[5].bias], self.t3, ACT_RELU)
        ops.dense([self.t3], [P[7].weight], [P[7].bias], out)
        return out

    def backward(self, d_out: torch.Tensor, grads: Dict[nn.Parameter, torch.Tensor]):
        base, ops, core = self.base, self.ops, self.base.gnn_core
        b, lib = self.batch, ops.lib
        G, Vc, L, het = self.G, self.Vc, self.L, self.hetero
        dev = d_out.device
        f32 = dict(dtype=torch.float32, device=dev)
        E = (L + 1) * F
        g = lambda p: grads[p]
        P = base.post_mp
        # ---- post_mp ----
        ops.wgrad(self.t3, d_out, _split_cols(g(P[7].weight)), [g(P[7].bias)])
        dt3 = torch.empty((G, 4 * F), **f32)
        ops.dgrad(d_out, P[7].weight, dt3)
        ops.act_backward(dt3, self.t3, ACT_RELU)
        ops.wgrad(self.t2, dt3, _split_cols(g(P[5].weight)), [g(P[5].bias)])
        dt2 = torch.empty((G, F), **f32)
        ops.dgrad(dt3, P[5].weight, dt2)
        ops.act_backward(dt2, self.t2, ACT_RELU)
        ops.wgrad(self.t1, dt2, _split_cols(g(P[3].weight)), [g(P[3].bias)])
        dt1 = torch.empty((G, F), **f32)
        ops.dgrad(dt2, P[3].weight, dt1)
        ops.act_backward(dt1, self.t1, ACT_LEAKY, 0.1)
        ops.wgrad(self.pooled, dt1, _split_cols(g(P[0].weight)), [g(P[0].bias)])
        dpooled = torch.empty((G, E), **f32)
        ops.dgrad(dt1, P[0].weight, dpooled)
        # ---- pooling, anchor ----
        demb_c = torch.empty((Vc, E), **f32)
        _lib.check(lib.desco_train_pool(_ptr(b.nbh_ptr), _ptr(self.crow_nbh), G, Vc, int(het), E, 1, _ptr(demb_c), E, 0, 0,
                                        _ptr(dpooled), E, ops.st), "desco_train_pool (backward)")
        demb_a = None
        if het:
            ops.act_backward(dpooled, self.z_a, ACT_LEAKY, 0.1)  # dpooled is now d(anchor pre-activation)
            A = base.anchor_mlp[0]
            ops.wgrad(self.emb_a, dpooled, _split_cols(g(A.weight)), [g(A.bias)])
            demb_a = torch.empty((G, E), **f32)
            ops.dgrad(dpooled, A.weight, demb_a)
        # ---- message-passing layers, last to first ----
        S = len(self.rel_c)
        dMc = torch.empty((Vc, F), **f32)
        dAc = torch.empty((Vc, S * F), **f32)
        dMa = torch.empty((G, F), **f32) if het else None
        dAa = torch.empty((G, 2 * F), **f32) if het else None
        for l in range(L - 1, -1, -1):
            hc = self.emb_c[:, F * l:F * (l + 1)]
            dh_next = demb_c[:, F * (l + 1):F * (l + 2)]
            ops.act_backward(dh_next, self.emb_c[:, F * (l + 1):F * (l + 2)], ACT_RELU)
            U = core.updates[l][self.ct]
            ops.wgrad(self.Mc[l], dh_next, [g(U.weight)[:, :F]], [g(U.bias)])
            ops.wgrad(hc, dh_next, [g(U.weight)[:, F:]])
            ops.dgrad(dh_next, U.weight[:, :F], dMc)
            ops.dgrad(dh_next, U.weight[:, F:], demb_c[:, F * l:F * (l + 1)], accumulate=True)
            convs = [core.convs[l][_key(r)].lin for r in self.rel_c]
            ops.wgrad(self.Ac[l], dMc, [g(c.weight) for c in convs], [g(c.bias) for c in convs])
            for s, c in enumerate(convs):
                ops.dgrad(dMc, c.weight, dAc[:, F * s:F * (s + 1)])
            if het:
                ha = self.emb_a[:, F * l:F * (l + 1)]
                da_next = demb_a[:, F * (l + 1):F * (l + 2)]
                ops.act_backward(da_next, self.emb_a[:, F * (l + 1):F * (l + 2)], ACT_RELU)
                U = core.updates[l]["canonical"]
                ops.wgrad(self.Ma[l], da_next, [g(U.weight)[:, :F]], [g(U.bias)])
                ops.wgrad(ha, da_next, [g(U.weight)[:, F:]])
                ops.dgrad(da_next, U.weight[:, :F], dMa)
                ops.dgrad(da_next, U.weight[:, F:], demb_a[:, F * l:F * (l + 1)], accumulate=True)
                convs = [core.convs[l][_key(r)].lin for r in self.rel_a]
                ops.wgrad(self.Aa[l], dMa, [g(c.weight) for c in convs], [g(c.bias) for c in convs])
                for s, c in enumerate(convs):
                    ops.dgrad(dMa, c.weight, dAa[:, F * s:F * (s + 1)])
            self._aggregate(l, True, demb_c[:, F * l:F * (l + 1)], demb_a[:, F * l:F * (l + 1)] if het else None, dAc, dAa)
        # ---- pre_mp: zero input features -> only the bias receives gradient ----
        ops.colsum(demb_c[:, :F], g(core.pre_mp[0][self.ct].bias))
        if het:
            ops.colsum(demb_a[:, :F], g(core.pre_mp[0]["canonical"].bias))


class _TrainForward(torch.autograd.Function):
    """loss = train_forward(batch) as ONE autograd node over the model parameters."""

    @staticmethod
    def forward(ctx, model, batch, y, *params):
        ops = _Ops()
        dev = params[0].device
        with torch.cuda.device(dev), torch.no_grad():
            ops.st = _stream()
            # gradients of this step: one flat zeroed buffer.  When the parameters' .grad are the views of a FusedAdam
            # flat buffer with the same layout, backward() adds the whole buffer into it with ONE launch (instead of a
            # fill, a scale and an accumulate per parameter = ~600 launches for the 200 tensors of the model).
            opt = _flat_optimizer(params)
            sizes = [p.numel() for p in params]
            flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            grads, off = {}, 0
            for p, k in zip(params, sizes):
                grads[p] = flat[off:off + k].view(p.shape)
                off += k
            G = batch.num_neighborhoods
            tape_t = _ShmpTape(model.emb_model, batch, ops)
            tape_q = _ShmpTape(model.emb_model_query, model.query_loader, ops)
            emb_q = tape_q.forward()  # the query GNN is re-run every step (lightning_model.py:233-236)
            emb_t = tape_t.forward()
            Q = emb_q.shape[0]
            f32 = dict(dtype=torch.float32, device=dev)
            lin1, lin2 = model.count_model[0], model.count_model[2]
            T = torch.empty((G, 4 * F), **f32)
            Bq = torch.empty((Q, 4 * F), **f32)
            ops.dense([emb_t], [lin1.weight[:, :F]], [], T)  # cat((t, q)) . W1^T split by halves (lightning_model.py:191)
            ops.dense([emb_q], [lin1.weight[:, F:]], [lin1.bias], Bq)
            y = y.to(device=dev, dtype=torch.float32).contiguous()
            assert y.shape == (G, Q), (y.shape, G, Q)
            dpred = torch.empty((G, Q), **f32)
            loss = torch.zeros(1, **f32)
            lib = ops.lib
            _lib.check(lib.desco_train_head_loss(_ptr(T), _ptr(Bq), _ptr(lin2.weight), _ptr(lin2.bias), _ptr(y), G, Q, 0,
                                                 _ptr(dpred), _ptr(loss), ops.st), "desco_train_head_loss")
            dT = torch.empty((G, 4 * F), **f32)
            dBq = torch.zeros((Q, 4 * F), **f32)
            _lib.check(lib.desco_train_head_backward(_ptr(T), _ptr(Bq), _ptr(lin2.weight), _ptr(dpred), G, Q, _ptr(dT),
                                                     _ptr(dBq), _ptr(grads[lin2.weight]), _ptr(grads[lin2.bias]), ops.st),
                       "desco_train_head_backward")
            gW1 = grads[lin1.weight]
            ops.wgrad(emb_t, dT, [gW1[:, :F]])
            ops.wgrad(emb_q, dBq, [gW1[:, F:]], [grads[lin1.bias]])
            d_t = torch.empty((G, F), **f32)
            d_q = torch.empty((Q, F), **f32)
            ops.dgrad(dT, lin1.weight[:, :F], d_t)
            ops.dgrad(dBq, lin1.weight[:, F:], d_q)
            tape_t.backward(d_t, grads)
            tape_q.backward(d_q, grads)
        ctx.flat, ctx.opt, ctx.shapes, ctx.sizes = flat, opt, [p.shape for p in params], sizes
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.opt is not None and ctx.opt.owns_grads():
            ctx.opt.flat_g.addcmul_(ctx.flat, grad_out.to(torch.float32))  # p.grad += dL/dp for every parameter at once
            return (None, None, None) + (None,) * len(ctx.sizes)
        scaled = ctx.flat * grad_out
        out, off = [], 0
        for shape, k in zip(ctx.shapes, ctx.sizes):
            out.append(scaled[off:off + k].view(shape))
            off += k
        return (None, None, None) + tuple(out)


def _flat_optimizer(params):
    """The FusedAdam whose flat gradient buffer lays these parameters out in this order, or None."""
    opt, off = None, 0
    for p in params:
        tag = getattr(p, "_desco_flat", None)
        if tag is None or (opt is not None and tag[0] is not opt) or tag[1] != off:
            return None
        opt = tag[0]
        off += p.numel()
    return opt if opt is not None and off == opt.flat_g.numel() else None


def train_forward(model, batch: NeighborhoodBatch, y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``lightning_model.py:228-254``: scalar loss with a grad_fn over every parameter that the reference trains
    (target GNN, query GNN, count head).  ``y`` [G, Q] = canonical count truth (default ``batch.y``)."""
    if model.query_loader is None:
        raise RuntimeError("call set_queries first")
    if y is None:
        y = getattr(batch, "y", None)
    if y is None:
        raise ValueError("train_forward needs the truth counts y[G, Q] (batch.y)")
    if batch.num_neighborhoods == 0:
        raise ValueError("empty batch")
    params = [p for p in model.parameters() if p.requires_grad]
    return _TrainForward.apply(model, batch, y, *params)


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` (``lightning_model.py:160-163``) as one kernel launch: parameters, gradients and both
    moments live in flat fp32 buffers (parameters and ``.grad`` are re-pointed to views of them)."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if not params or not all(p.is_cuda and p.dtype == torch.float32 for p in params):
            raise RuntimeError("FusedAdam needs fp32 CUDA parameters (there is no CPU fallback)")
        self._params = params
        n = sum(p.numel() for p in params)
        dev = params[0].device
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self._views = []
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                pv = self.flat_p[off:off + k].view_as(p)
                pv.copy_(p)
                p.data = pv
                gv = self.flat_g[off:off + k].view_as(p)
                if p.grad is not None:
                    gv.copy_(p.grad)
                p.grad = gv
                p._desco_flat = (self, off)
                self._views.append(gv)
                off += k
        self.step_count = 0

    def owns_grads(self) -> bool:
        """True while every parameter's .grad still is its view of the flat gradient buffer."""
        return all(p.grad is not None and p.grad.data_ptr() == gv.data_ptr() for p, gv in zip(self._params, self._views))

    def zero_grad(self, set_to_none: bool = False):
        self.flat_g.zero_()
        for p, gv in zip(self._params, self._views):
            p.grad = gv

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for p, gv in zip(self._params, self._views):
            if p.grad is None:
                gv.zero_()
            elif p.grad.data_ptr() != gv.data_ptr():  # autograd replaced the view: bring the values home
                gv.copy_(p.grad)
            p.grad = gv
        grp = self.param_groups[0]
        self.step_count += 1
        lib = _lib.load()
        with torch.cuda.device(self.flat_p.device):
            _lib.check(lib.desco_train_adam(_ptr(self.flat_p), _ptr(self.flat_g), _ptr(self.flat_m), _ptr(self.flat_v),
                                            self.flat_p.numel(), float(grp["lr"]), float(grp["betas"][0]),
                                            float(grp["betas"][1]), float(grp["eps"]), float(grp["weight_decay"]),
                                            self.step_count, _stream()), "desco_train_adam")
        # the update went through a raw pointer: no Parameter._version moved, so the packed-weight / query-embedding
        # caches of the models (gnn_model._PackedWeightsMixin) must be told
        from .gnn_model import bump_weights_epoch

        bump_weights_epoch()
        return loss
