"""Training step of the gossip model on the GPU - host side of csrc/gossip_train.cu.

Mirrors ``GossipCountingModel.train_forward`` / ``criterion`` (``subgraph_counting/lightning_model.py:585-608, 630-635``,
reference @ 4508f7a): per query, ``loss_q = sum_i log2(|x_iq + gossip(x_q)_i - y_iq| + 1)``, summed over the queries, with
autograd through ``GossipConv`` (``gnn_model.py:294-348``) and ``post_mp``.  The input of the first layer is detached in
the reference (``gnn_model.py:236-240``), so ``pre_mp`` and the (unused) ``anchor_mlp`` receive no gradient.

``train_forward`` is one autograd node over the gossip parameters: forward and backward run as raw launches of the dense /
weight-gradient / aggregation primitives (csrc/train.cu, csrc/conv.cu) plus the gate / dropout / loss kernels of
csrc/gossip_train.cu, query by query like the reference.  It is the correctness path of SURVEY.md section 8 row f3, not
a tuned one: gossip TRAINING is outside BASELINE.json's metric.  There is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import _lib
from .data import _ptr, _stream
from .training import ACT_LEAKY, ACT_NONE, ACT_RELU, _Ops

F = 64


class _GossipOps(_Ops):
    def spmm(self, rowptr, col, w, x: torch.Tensor, out: torch.Tensor):
        _lib.check(self.lib.desco_spmm_sum(_ptr(rowptr), _ptr(col), _ptr(w), x.shape[0], _ptr(x), x.stride(0), x.shape[1],
                                           _ptr(out), out.stride(0), self.st), "desco_spmm_sum")

    def mix(self, a, b, gate, out):
        assert a.is_contiguous() and b.is_contiguous() and out.is_contiguous()
        _lib.check(self.lib.desco_gossip_gated_mix(_ptr(a), _ptr(b), _ptr(gate), _ptr(out), a.numel(), self.st), "desco_gossip_gated_mix")

    def gate(self, qe, lin_gate, out):
        l1, l2 = lin_gate[0], lin_gate[2]
        _lib.check(self.lib.desco_gossip_gate(_ptr(qe), 1, qe.shape[1], _ptr(l1.weight), _ptr(l1.bias), l1.weight.shape[0],
                                              _ptr(l2.weight), _ptr(l2.bias), _ptr(out), self.st), "desco_gossip_gate")

    def gate_backward(self, qe, lin_gate, dgate, g: Dict):
        l1, l2 = lin_gate[0], lin_gate[2]
        _lib.check(self.lib.desco_gossip_gate_backward(_ptr(qe), qe.shape[1], _ptr(l1.weight), _ptr(l1.bias), l1.weight.shape[0],
                                                       _ptr(l2.weight), _ptr(l2.bias), _ptr(dgate), _ptr(g[l1.weight]), _ptr(g[l1.bias]),
                                                       _ptr(g[l2.weight]), _ptr(g[l2.bias]), self.st), "desco_gossip_gate_backward")

    def gate_grad(self, d, a, b, dgate):
        assert d.is_contiguous() and a.is_contiguous() and b.is_contiguous()
        _lib.check(self.lib.desco_gossip_gate_grad(_ptr(d), _ptr(a), _ptr(b), d.numel(), _ptr(dgate), self.st), "desco_gossip_gate_grad")

    def dropout(self, x, mask, scale):
        assert x.is_contiguous() and mask.is_contiguous()
        _lib.check(self.lib.desco_train_dropout(_ptr(x), _ptr(mask), float(scale), x.numel(), self.st), "desco_train_dropout")

    def loss(self, c, out, y, dout, loss):
        _lib.check(self.lib.desco_gossip_loss(_ptr(c), c.stride(0), _ptr(out), out.stride(0), _ptr(y), y.stride(0), c.shape[0], 0,
                                              _ptr(dout), dout.stride(0), _ptr(loss), self.st), "desco_gossip_loss")


def _loss_and_grads(model, rowptr, col, x, y, query_emb):
    """Loss (device scalar) and {parameter: gradient} of one training batch."""
    base = model.emb_model
    core = base.gnn_core
    dev = base.post_mp[0].weight.device
    f32 = dict(dtype=torch.float32, device=dev)
    x = x.to(**f32).contiguous()
    y = y.to(**f32).contiguous()
    qe_all = query_emb.detach().to(**f32).contiguous()
    N, Q = x.shape
    p_drop = float(core.dropout) if base.training else 0.0
    keep_scale = 1.0 / (1.0 - p_drop) if p_drop > 0 else 1.0
    rowptr = rowptr.to(torch.int32)
    col = col.to(torch.int32)
    deg = (rowptr[1:] - rowptr[:-1]).long()
    row = torch.repeat_interleave(torch.arange(N, device=dev), deg)
    w_lt = (col.long() < row).to(torch.float32).contiguous()  # source < target: the gated direction (gnn_model.py:248)
    w_gt = (1.0 - w_lt).contiguous()
    params = [p for p in base.parameters()]
    g = {p: torch.zeros_like(p) for p in params}
    c0, c1 = core.convs[0], core.convs[1]
    P0, P1, P2, P3 = base.post_mp[0], base.post_mp[3], base.post_mp[5], base.post_mp[7]
    # Linear(1, 64) and Linear(256, 1) through the 64-wide dense primitive: zero-padded copies (index plumbing)
    Wpre = torch.zeros((F, F), **f32)
    Wpre[:, 0] = core.pre_mp[0].weight.detach()[:, 0]
    W3 = torch.zeros((F, 4 * F), **f32)
    W3[0] = P3.weight.detach()[0]
    b3 = torch.zeros(F, **f32)
    b3[0] = P3.bias.detach()[0]
    dW3 = torch.zeros_like(W3)
    db3 = torch.zeros_like(b3)
    loss = torch.zeros(1, **f32)
    with torch.cuda.device(dev):
        ops = _GossipOps()
        gate = torch.empty(2, **f32)
        dgate = torch.zeros(2, **f32)
        X0 = torch.zeros((N, F), **f32)
        emb = torch.empty((N, 4 * F), **f32)          # [qe | pre_mp(c) | x1 | x2]
        cat0 = torch.empty((N, 3 * F), **f32)         # [a0 | x0]
        cat1 = torch.empty((N, 2 * F), **f32)         # [a1 | x1]
        yv, alt, agt, a, t1, t2 = (torch.empty((N, F), **f32) for _ in range(6))
        h1, h2, out = (torch.empty((N, F), **f32) for _ in range(3))
        h3 = torch.empty((N, 4 * F), **f32)
        dout = torch.zeros((N, F), **f32)
        dh3 = torch.empty((N, 4 * F), **f32)
        dh2, dh1, dx2, dx1, da, dy = (torch.empty((N, F), **f32) for _ in range(6))
        demb = torch.empty((N, 4 * F), **f32)
        for q in range(Q):
            qe = qe_all[q:q + 1]
            c = x[:, q]
            X0[:, 0] = c
            emb[:, :F] = qe
            ops.dense([X0], [Wpre], [core.pre_mp[0].bias], emb[:, F:2 * F])          # pre_mp (gnn_model.py:231); detached below
            ops.gate(qe, c0.lin_gate, gate[0:1])
            ops.gate(qe, c1.lin_gate, gate[1:2])
            masks: List = []
            # ---- forward ----
            x0 = emb[:, :2 * F]
            for l, (conv, xin, cat, xout) in enumerate(((c0, x0, cat0, emb[:, 2 * F:3 * F]), (c1, emb[:, 2 * F:3 * F], cat1, emb[:, 3 * F:]))):
                ops.dense([xin], [conv.lin_com.weight], [conv.lin_com.bias], yv)        # lin_com, once per node
                ops.spmm(rowptr, col, w_lt, yv, alt)
                ops.spmm(rowptr, col, w_gt, yv, agt)
                ops.mix(alt, agt, gate[l:l + 1], a)                                    # gate * (j < i) + (1 - gate) * (j > i)
                cat[:, :F] = a
                cat[:, F:] = xin
                oc = conv.out_channels
                ops.dense([cat], [conv.lin_update.weight], [conv.lin_update.bias], xout, ACT_RELU)  # relu(lin_update(cat(aggr, x)))
                if l == 0:
                    alt0, agt0 = alt.clone(), agt.clone()
                if p_drop > 0:
                    m = (torch.rand((N, F), device=dev) >= p_drop).to(torch.uint8)
                    masks.append(m)
                    tmp = xout.contiguous()
                    ops.dropout(tmp, m, keep_scale)
                    xout.copy_(tmp)
            ops.dense([emb], [P0.weight], [P0.bias], h1, ACT_LEAKY, 0.1)
            h1_act = h1
            if p_drop > 0:  # nn.Dropout sits before the LeakyReLU; both orders agree (positive homogeneity)
                m = (torch.rand((N, F), device=dev) >= p_drop).to(torch.uint8)
                masks.append(m)
                h1_act = h1.clone()
                ops.dropout(h1, m, keep_scale)
            ops.dense([h1], [P1.weight], [P1.bias], h2, ACT_RELU)
            ops.dense([h2], [P2.weight], [P2.bias], h3, ACT_RELU)
            ops.dense([h3], [W3], [b3], out)
            ops.loss(c, out[:, 0], y[:, q], dout[:, 0], loss)
            # ---- backward ----
            ops.dgrad(dout, W3, dh3)
            ops.wgrad(h3, dout, [dW3[:, F * i:F * (i + 1)] for i in range(4)], [db3])
            ops.act_backward(dh3, h3, ACT_RELU)
            ops.dgrad(dh3, P2.weight.detach(), dh2)
            ops.wgrad(h2, dh3, [g[P2.weight]], [g[P2.bias]])
            ops.act_backward(dh2, h2, ACT_RELU)
            ops.dgrad(dh2, P1.weight.detach(), dh1)
            ops.wgrad(h1, dh2, [g[P1.weight]], [g[P1.bias]])
            if p_drop > 0:
                ops.dropout(dh1, masks[2], keep_scale)
            ops.act_backward(dh1, h1_act, ACT_LEAKY, 0.1)
            ops.dgrad(dh1, P0.weight.detach(), demb)
            ops.wgrad(emb, dh1, [g[P0.weight][:, F * i:F * (i + 1)] for i in range(4)], [g[P0.bias]])
            dx1.copy_(demb[:, 2 * F:3 * F])
            dx2.copy_(demb[:, 3 * F:])
            for l, conv, cat, xin, xout, dxo, lt, gt in ((1, c1, cat1, emb[:, 2 * F:3 * F], emb[:, 3 * F:], dx2, alt, agt),
                                                         (0, c0, cat0, x0, emb[:, 2 * F:3 * F], dx1, alt0, agt0)):
                if p_drop > 0:
                    ops.dropout(dxo, masks[l], keep_scale)
                xo = xout.contiguous()
                ops.act_backward(dxo, xo, ACT_RELU)  # (a dropped unit has dxo = 0 already; relu' read off the kept output)
                Wup = conv.lin_update.weight
                ops.wgrad(cat, dxo, [g[Wup][:, F * i:F * (i + 1)] for i in range(Wup.shape[1] // F)], [g[conv.lin_update.bias]])
                ops.dgrad(dxo, Wup.detach()[:, :F], da)
                if l == 1:
                    ops.dgrad(dxo, Wup.detach()[:, F:], dx1, accumulate=True)
                dgate.zero_()
                ops.gate_grad(da, lt, gt, dgate[l:l + 1])
                ops.gate_backward(qe, conv.lin_gate, dgate[l:l + 1], g)
                ops.spmm(rowptr, col, w_gt, da, t1)   # adjoint of the gated aggregation on a symmetric edge set:
                ops.spmm(rowptr, col, w_lt, da, t2)   # the two directions swap roles
                ops.mix(t1, t2, gate[l:l + 1], dy)
                Wc = conv.lin_com.weight
                ops.wgrad(xin.contiguous(), dy, [g[Wc][:, F * i:F * (i + 1)] for i in range(Wc.shape[1] // F)], [g[conv.lin_com.bias]])
                if l == 1:
                    ops.dgrad(dy, Wc.detach(), dx1, accumulate=True)
        g[P3.weight][0] += dW3[0]
        g[P3.bias][0] += db3[0]
    return loss[0], g


class _GossipTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, rowptr, col, x, y, query_emb, *params):
        loss, grads = _loss_and_grads(model, rowptr, col, x, y, query_emb)
        ctx.grads = [grads[p] for p in params]
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        return (None, None, None, None, None, None) + tuple(grad_out * g for g in ctx.grads)


def gossip_train_forward(model, batch, query_emb=None) -> torch.Tensor:
    """``GossipCountingModel.train_forward`` (``lightning_model.py:585-608``): batch carries the target CSR (``graph``), the
    neighborhood counts ``x`` [N, Q] and the ground truth ``y`` [N, Q]."""
    qe = model.query_emb if query_emb is None else query_emb
    if qe is None:
        raise RuntimeError("call set_query_emb first")
    graph = getattr(batch, "graph", batch)
    params = [p for p in model.emb_model.parameters()]
    return _GossipTrainFn.apply(model, graph.rowptr, graph.col, batch.x, batch.y, qe, *params)
