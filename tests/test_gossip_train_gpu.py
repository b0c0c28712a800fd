"""GPU parity: training step of the gossip model (GossipCountingModel.train_forward through csrc/gossip_train.cu + the
dense / aggregation primitives) - loss and every parameter gradient against torch autograd on the oracle
(lightning_model.py:585-608, 630-635; gnn_model.py:294-348)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from desco_b200.graph import gen_enzymes_shaped, gen_mutag_shaped, gen_powerlaw

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _pair(seed, dropout=0.0):
    from desco_b200.lightning_model import GossipCountingModel, default_gossip_args
    from oracle import model as M

    torch.manual_seed(seed)
    om = M.GossipCountingModel(args=M.default_gossip_args(dropout=dropout))
    pm = GossipCountingModel(args=default_gossip_args(dropout=dropout))
    pm.emb_model.load_state_dict(om.emb_model.state_dict())
    return om, pm.cuda()


@pytest.mark.parametrize("gen,kw,Q", [(gen_mutag_shaped, dict(num_graphs=12), 5), (gen_enzymes_shaped, dict(num_graphs=6), 3),
                                       (gen_powerlaw, dict(n=600, m_undirected=2500), 2)])
def test_gossip_train_forward_loss_and_grads_match_autograd(cuda_device, gen, kw, Q):
    from desco_b200.data import DeviceCSR

    om, pm = _pair(3)
    csr = gen(seed=4, **kw)
    g = torch.Generator().manual_seed(5)
    x = torch.floor(torch.exp(torch.randn(csr.num_nodes, Q, generator=g)))
    y = torch.floor(torch.exp(torch.randn(csr.num_nodes, Q, generator=g) * 1.5))
    qe = torch.randn(Q, 64, generator=g)
    om.train()
    om.set_query_emb(qe)
    ref_loss = om.train_forward(x, y, torch.from_numpy(csr.edge_index()))
    ref_loss.backward()
    pm.train()
    pm.set_query_emb(qe.cuda())
    batch = SimpleNamespace(graph=DeviceCSR.from_host(csr), x=x.cuda(), y=y.cuda())
    loss = pm.train_forward(batch)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= TOL * max(1.0, abs(ref_loss.item()))
    ref = dict(om.emb_model.named_parameters())
    checked = 0
    for name, p in pm.emb_model.named_parameters():
        r = ref[name].grad
        if r is None:  # pre_mp (detached input, gnn_model.py:236-240) and the unused anchor_mlp
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        d = (p.grad.cpu() - r).abs().max().item()
        assert d <= TOL * max(1.0, r.abs().max().item()), f"{name}: {d} vs scale {r.abs().max().item()}"
        checked += 1
    assert checked >= 18  # 2 x (lin_com, lin_update, lin_gate.0, lin_gate.2) + post_mp, weights and biases


def test_gossip_training_reduces_the_loss_and_handles_dropout(cuda_device):
    from desco_b200.data import DeviceCSR

    _, pm = _pair(6, dropout=0.01)  # the reference default (config.py:316)
    csr = gen_mutag_shaped(seed=7, num_graphs=20)
    g = torch.Generator().manual_seed(8)
    Q = 4
    x = torch.floor(torch.exp(torch.randn(csr.num_nodes, Q, generator=g))).cuda()
    y = (x + torch.floor(torch.exp(torch.randn(csr.num_nodes, Q, generator=g))).cuda())
    pm.set_query_emb(torch.randn(Q, 64, generator=g).cuda())
    batch = SimpleNamespace(graph=DeviceCSR.from_host(csr), x=x, y=y)
    pm.train()
    opt = pm.configure_optimizers()["optimizer"]
    losses = []
    for step in range(12):
        opt.zero_grad()
        loss = pm.training_step(batch, step)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and np.mean(losses[-3:]) < np.mean(losses[:3])
    pm.eval()
    with torch.no_grad():
        out = pm.graph_to_count(batch)  # the fused inference path sees the trained weights
        val = pm.validation_step(batch)
    assert torch.isfinite(out).all() and torch.isfinite(val)
    ref_val = float(torch.log2((out - y).abs() + 1).sum())
    assert abs(val.item() - ref_val) <= 1e-3 * max(1.0, abs(ref_val))
