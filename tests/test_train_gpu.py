"""GPU parity: training step of the neighborhood-counting model (csrc/train.cu through the C ABI) vs torch autograd on
the oracle (lightning_model.py:228-254).  Tolerance: loss and every parameter gradient within 1e-4 of the fp32 oracle,
checked per tensor as max|d| <= tol * max|ref| (gradients of one tensor share a scale); one Adam step within 1e-6 abs."""
import numpy as np
import pytest
import torch

from desco_b200.graph import gen_enzymes_shaped, gen_imdb_shaped, gen_mutag_shaped, gen_syn1827_shaped

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _setup(csr, seed=0, max_nbh=None, pyg_bs=None):
    from desco_b200.data import NeighborhoodBatch
    from desco_b200.lightning_model import STANDARD_QUERY_IDS, NeighborhoodCountingModel
    from oracle import model as M
    from oracle import partition as P

    b_np = P.partition_dataset(csr, 4, mode="hetero")
    if max_nbh is not None and len(b_np["centre"]) > max_nbh:
        G = max_nbh
        V, E = int(b_np["nbh_ptr"][G]), int(b_np["edge_ptr"][int(b_np["nbh_ptr"][G])])
        b_np = dict(b_np, nbh_ptr=b_np["nbh_ptr"][:G + 1], edge_ptr=b_np["edge_ptr"][:V + 1], edge_col=b_np["edge_col"][:E],
                    edge_tri=b_np["edge_tri"][:E], centre=b_np["centre"][:G], node_gid=b_np["node_gid"][:V])
    G = len(b_np["nbh_ptr"]) - 1
    torch.manual_seed(seed)
    om = M.NeighborhoodCountingModel().train()
    with torch.no_grad():  # spread the pre-exponent so that both smooth-L1 branches are exercised
        om.count_model[2].weight.mul_(30.0)
    pm = NeighborhoodCountingModel()
    pm.load_state_dict(om.state_dict())
    pm = pm.cuda().train()
    pm.set_queries(STANDARD_QUERY_IDS)
    pm.set_pyg_batch_size(pyg_bs or 0)
    rng = np.random.default_rng(seed + 3)
    y = torch.from_numpy(np.floor(np.exp(rng.normal(0.0, 1.5, size=(G, 29)))).astype(np.float32))
    batch = NeighborhoodBatch.from_numpy(b_np)
    batch.y = y.cuda()
    return om, pm, b_np, batch, y


def _compare_grads(om, pm, tol=TOL):
    worst = 0.0
    ref = dict(om.named_parameters())
    for name, p in pm.named_parameters():
        r = ref[name].grad
        g = p.grad
        assert g is not None, name
        if r is None:
            r = torch.zeros_like(ref[name])
        scale = max(r.abs().max().item(), 1e-12)
        d = (g.cpu() - r).abs().max().item() / scale
        if r.abs().max().item() == 0.0:
            assert g.abs().max().item() == 0.0, name
            continue
        worst = max(worst, d)
        assert d <= tol, f"{name}: grad diff {d} (scale {scale})"
    return worst


@pytest.mark.parametrize("gen,kw,pyg_bs", [(gen_mutag_shaped, dict(num_graphs=12), None),
                                           (gen_enzymes_shaped, dict(num_graphs=10), 64),
                                           (gen_imdb_shaped, dict(num_graphs=8), None)])
def test_train_forward_loss_and_grads_match_autograd(cuda_device, gen, kw, pyg_bs):
    from oracle import model as M

    om, pm, b_np, batch, y = _setup(gen(seed=1, **kw), seed=0, pyg_bs=pyg_bs)
    ref_loss = om.train_forward(b_np, M.query_batch(), y, pyg_batch_size=pyg_bs)
    ref_loss.backward()
    loss = pm.training_step(batch, 0)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss.item()) <= TOL * max(1.0, abs(ref_loss.item())), (loss.item(), ref_loss.item())
    _compare_grads(om, pm)


def test_train_large_neighborhoods_and_quirk(cuda_device):
    """Syn_1827-shaped neighborhoods (config 3: up to hundreds of rows) and a PyG batch whose first neighborhoods have two
    rows (the remove_self_loops quirk, gnn_model.py:389-390)."""
    import networkx as nx

    from desco_b200.graph import csr_from_networkx
    from oracle import model as M

    om, pm, b_np, batch, y = _setup(gen_syn1827_shaped(seed=0, stride=300), seed=2, max_nbh=256)
    ref_loss = om.train_forward(b_np, M.query_batch(), y)
    ref_loss.backward()
    loss = pm.train_forward(batch)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= TOL * max(1.0, abs(ref_loss.item()))
    _compare_grads(om, pm)

    g = nx.Graph([(0, 1), (2, 3), (3, 4), (2, 4), (5, 6), (6, 7), (7, 8), (5, 8)])
    om, pm, b_np, batch, y = _setup(csr_from_networkx([g]), seed=4)
    ref_loss = om.train_forward(b_np, M.query_batch(), y)
    ref_loss.backward()
    loss = pm.train_forward(batch)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= TOL * max(1.0, abs(ref_loss.item()))
    _compare_grads(om, pm)


def test_adam_steps_match_torch_adam(cuda_device):
    """Three optimisation steps at the reference's lr (1e-4, lightning_model.py:160-163): FusedAdam on the product
    model vs torch.optim.Adam on the oracle; the loss trajectories coincide, go down, and the parameters stay within 5 %
    of one step's update of each other."""
    from oracle import model as M

    om, pm, b_np, batch, y = _setup(gen_mutag_shaped(seed=5, num_graphs=16), seed=1)
    pm.lr = 1e-4
    cfg = pm.configure_optimizers()
    opt = cfg["optimizer"]
    ref_opt = torch.optim.Adam(om.parameters(), lr=1e-4, weight_decay=0.0)
    qb = M.query_batch()
    losses, ref_losses = [], []
    for step in range(3):
        ref_opt.zero_grad()
        rl = om.train_forward(b_np, qb, y)
        rl.backward()
        ref_opt.step()
        opt.zero_grad()
        loss = pm.training_step(batch, step)
        loss.backward()
        opt.step()
        losses.append(loss.item())
        ref_losses.append(rl.item())
    assert losses[-1] < losses[0] and ref_losses[-1] < ref_losses[0]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (losses, ref_losses)
    ref = dict(om.named_parameters())
    for name, p in pm.named_parameters():
        r = ref[name].detach()
        d = (p.detach().cpu() - r).abs().max().item()
        # Adam normalises the update to ~lr per step whatever the gradient scale, so compare on the lr scale
        assert d <= 3 * 1e-4 * 0.1, f"{name}: parameter drift {d}"


def test_training_then_inference_uses_new_weights(cuda_device):
    """The packed-weight caches of the inference path are invalidated by an optimizer step."""
    om, pm, b_np, batch, y = _setup(gen_mutag_shaped(seed=6, num_graphs=8), seed=3)
    pm.eval()
    with torch.no_grad():
        before = pm.graph_to_count(batch).clone()
    pm.train()
    opt = pm.configure_optimizers()["optimizer"]
    for g in opt.param_groups:
        g["lr"] = 1e-2
    opt.zero_grad()
    pm.training_step(batch, 0).backward()
    opt.step()
    pm.eval()
    with torch.no_grad():
        after = pm.graph_to_count(batch)
    assert (after - before).abs().max().item() > 0


def test_monitoring_in_train_mode_sees_every_optimizer_step(cuda_device):
    """FusedAdam updates the flat parameter buffer through a raw pointer (no Parameter._version moves): calling
    graph_to_count / get_query_emb between steps WITHOUT an eval()/train() toggle - periodic monitoring inside a training
    loop - must still see the new weights, in both grad modes."""
    om, pm, b_np, batch, y = _setup(gen_mutag_shaped(seed=7, num_graphs=8), seed=4)
    pm.train()
    opt = pm.configure_optimizers()["optimizer"]
    for g in opt.param_groups:
        g["lr"] = 1e-2
    seen, seen_q = [], []
    for step in range(3):
        with torch.no_grad():
            seen.append(pm.graph_to_count(batch).clone())
            seen_q.append(pm.get_query_emb().clone())
        opt.zero_grad()
        pm.training_step(batch, step).backward()
        opt.step()
    with torch.no_grad():
        seen.append(pm.graph_to_count(batch).clone())
    assert all((a - b).abs().max().item() > 0 for a, b in zip(seen[:-1], seen[1:]))
    assert all((a - b).abs().max().item() > 0 for a, b in zip(seen_q[:-1], seen_q[1:]))
    # eval + no_grad after an optimizer step taken while already in eval mode (the frozen-cache regime)
    pm.eval()
    with torch.no_grad():
        a = pm.graph_to_count(batch).clone()
    pm.train_forward(batch).backward()
    opt.step()
    with torch.no_grad():
        b = pm.graph_to_count(batch)
    assert (a - b).abs().max().item() > 0
    # weight surgery in eval mode needs the public invalidate_caches()
    with torch.no_grad():
        pm.count_model[2].bias.add_(1.0)
        pm.invalidate_caches()
        c = pm.graph_to_pred(batch)
        b_pred = torch.log2(b + 1)
    assert ((c - b_pred) - 1.0).abs().max().item() < 1e-3
