"""GPU: the PyG-shaped transform / module surface: NetworkxToHetero -> ToTconvHetero (in place,
typing on the GPU) -> collate -> BaseGNN.forward / graph_to_count on HeteroData input, against the fixture produced by the
reference's own pipeline (tests/golden/shmp_pipeline_ref.npz) and the literal A*A@A+A formulation."""
import os

import networkx as nx
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "shmp_pipeline_ref.npz"))
    return z, {k[2:]: z[k] for k in z.files if k.startswith("b_")}


def _literal_flags(ei, n):
    """transforms.py:201-221 on a dense matrix: edge (u, v) is a triangle edge iff (A * A@A + A)[u, v] > 1."""
    A = torch.zeros(n, n)
    A[ei[0], ei[1]] = 1.0
    A.fill_diagonal_(0.0)
    T = A * (A @ A) + A
    return T[ei[0], ei[1]] > 1


def test_to_tconv_hetero_in_place_matches_reference_types(cuda_device, golden_dir):
    from desco_b200.transforms import NetworkxToHetero, ToTconvHetero
    from oracle import partition as P

    z, b = _fixture(golden_dir)
    graphs = P.neighborhoods_as_networkx(b)
    e_seen = 0
    for g_i, g in enumerate(graphs[:60]):
        data = NetworkxToHetero(g, type_key="type", feat_key="feat")
        out = ToTconvHetero()(data)
        assert out is data  # mutates in place and returns its argument (transforms.py:184-187)
        node_types, edge_types = data.metadata()
        assert all(r.endswith("_triangle") or r.endswith("_tride") for _, r, _ in edge_types)
        # compare with the flags the REFERENCE's ToTconvHetero assigned (stored in packed edge order)
        lo, hi = int(b["nbh_ptr"][g_i]), int(b["nbh_ptr"][g_i + 1])
        gid = b["node_gid"][lo:hi]
        order = {"count": [u for u in g.nodes if g.nodes[u]["type"] == "count"],
                 "canonical": [u for u in g.nodes if g.nodes[u]["type"] == "canonical"]}
        row = {int(u): lo + i for i, u in enumerate(gid)}
        for (s, r, d) in edge_types:
            for a, c in data[s, r, d].edge_index.T.tolist():
                ru, rv = row[order[s][a]], row[order[d][c]]
                es = np.arange(b["edge_ptr"][rv], b["edge_ptr"][rv + 1])
                e = es[b["edge_col"][es] == ru]
                assert len(e) == 1 and int(b["edge_tri"][e[0]]) == (1 if r.endswith("_triangle") else 0)
                e_seen += 1
    assert e_seen > 500


def test_to_tconv_single_type_matches_literal_formula(cuda_device):
    from desco_b200.hetero import HeteroData
    from desco_b200.transforms import ToTCONV, ToTconvHetero

    rng = np.random.default_rng(0)
    for n, m in ((12, 30), (40, 200), (5, 4)):
        g = nx.gnm_random_graph(n, m, seed=int(rng.integers(1 << 30)))
        ei = torch.tensor([(a, c) for a, c in g.to_directed().edges], dtype=torch.long).T.reshape(2, -1)
        ei = torch.cat([ei, torch.tensor([[0], [0]])], 1)  # a self loop: removed by ToTCONV (transforms.py:82)
        d = HeteroData()
        d["count"].x = torch.zeros(n, 1)
        d["count", "union", "count"].edge_index = ei.clone()
        d["count", "union", "canonical"].edge_index = torch.zeros((2, 0), dtype=torch.long)  # untouched by ToTCONV
        out = ToTCONV(node_type="count", node_attr="x")(d)
        assert out is d and ("count", "union", "canonical") in d.metadata()[1]
        tri, trd = d["count", "union_triangle", "count"].edge_index, d["count", "union_tride", "count"].edge_index
        clean = ei[:, ei[0] != ei[1]]
        key = torch.unique(clean[0] * n + clean[1])
        co = torch.stack([key // n, key % n])
        flags = _literal_flags(co, n)
        assert torch.equal(tri, co[:, flags]) and torch.equal(trd, co[:, ~flags])
        # the hetero transform on the same single-type graph (the query path, lightning_model.py:84-85)
        q = HeteroData()
        q["union_node"].node_feature = torch.zeros(n, 1)
        q["union_node", "union", "union_node"].edge_index = clean.clone()
        ToTconvHetero()(q)
        f2 = _literal_flags(clean, n)
        assert torch.equal(q["union_node", "union_triangle", "union_node"].edge_index, clean[:, f2])
        assert torch.equal(q["union_node", "union_tride", "union_node"].edge_index, clean[:, ~f2])


@pytest.mark.parametrize("typed_by", ["transform", "kernel_at_packing"])
def test_pyg_shaped_batches_through_the_model_match_the_reference_pipeline(cuda_device, golden_dir, typed_by):
    """The drop-in path with PyG-shaped input: per neighborhood NetworkxToHetero (+ ToTconvHetero), collate 64 per batch
    like the DataLoader, hand the Batch to graph_to_count - against the counts of the reference's own pipeline."""
    from desco_b200.hetero import Batch
    from desco_b200.lightning_model import STANDARD_QUERY_IDS, NeighborhoodCountingModel
    from desco_b200.transforms import NetworkxToHetero, ToTconvHetero
    from oracle import model as M
    from oracle import partition as P

    z, b = _fixture(golden_dir)
    torch.manual_seed(int(z["seed"]))
    om = M.NeighborhoodCountingModel().eval()
    pm = NeighborhoodCountingModel().eval()
    pm.load_state_dict(om.state_dict())
    pm = pm.cuda()
    pm.set_queries(STANDARD_QUERY_IDS)
    pm.set_pyg_batch_size(512)  # ignored for PyG-shaped input: the Batch itself is the collated batch
    datas = []
    for g in P.neighborhoods_as_networkx(b):
        d = NetworkxToHetero(g, type_key="type", feat_key="feat")
        for et in [("count", "union", "count"), ("count", "union", "canonical"), ("canonical", "union", "count")]:
            if et not in d.metadata()[1]:  # workload.py:275-282
                d[et].edge_index = torch.empty((2, 0), dtype=torch.long)
        datas.append(ToTconvHetero()(d) if typed_by == "transform" else d)
    bs = int(z["pyg_batch_size"])
    outs = []
    with torch.no_grad():
        for i in range(0, len(datas), bs):
            outs.append(pm.graph_to_count(Batch.from_data_list(datas[i:i + bs])).cpu())
    got = torch.cat(outs, 0)
    assert (got - torch.from_numpy(z["count"])).abs().max().item() <= TOL


def test_sage_and_gossip_conv_modules_are_callable_like_the_reference(cuda_device):
    """SAGEConv.forward(x | (x_src, x_dst), edge_index) and GossipConv.forward(x, edge_index, edge_weight, query_emb)
    (gnn_model.py:303-359, 372-404) through the CUDA primitives, against their literal definitions."""
    from desco_b200.gnn_model import GossipConv, SAGEConv

    torch.manual_seed(0)
    n, e = 50, 400
    ei = torch.randint(0, n, (2, e), device="cuda")
    x = torch.randn(n, 64, device="cuda")
    conv = SAGEConv(64, 64).cuda()
    with torch.no_grad():
        got = conv(x, ei)
        keep = ei[0] != ei[1]  # gnn_model.py:389-390
        agg = torch.zeros(n, 64, device="cuda").index_add_(0, ei[1][keep], x[ei[0][keep]])
        ref = conv.lin(agg)
        assert (got - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
        xs, xd = torch.randn(30, 64, device="cuda"), torch.randn(20, 64, device="cuda")
        eb = torch.stack([torch.randint(0, 30, (100,), device="cuda"), torch.randint(0, 20, (100,), device="cuda")])
        got = conv((xs, xd), eb)
        keep = eb[0] != eb[1]
        ref = conv.lin(torch.zeros(20, 64, device="cuda").index_add_(0, eb[1][keep], xs[eb[0][keep]]))
        assert got.shape == (20, 64) and (got - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
        empty = conv(x, torch.zeros((2, 0), dtype=torch.long, device="cuda"))  # gnn_model.py:384-388
        assert (empty - conv.lin.bias).abs().max().item() <= 1e-6
    for cin in (128, 64):
        gc = GossipConv(cin, 64, 64).cuda()
        xg = torch.randn(n, cin, device="cuda")
        qe = torch.randn(1, 64, device="cuda")
        und = torch.unique(torch.cat([ei[0] * n + ei[1], ei[1] * n + ei[0]]))
        eu = torch.stack([und // n, und % n])
        eu = eu[:, eu[0] != eu[1]]
        w = eu[0] < eu[1]
        with torch.no_grad():
            got = gc(xg, eu, edge_weight=w, query_emb=qe)  # keyword call, like gnn_model.py:258-260
            gate = gc.lin_gate(qe)
            msg = gc.lin_com(xg[eu[0]])
            msg = torch.where(w.view(-1, 1), msg * gate, msg * (1 - gate))
            aggr = torch.zeros(n, 64, device="cuda").index_add_(0, eu[1], msg)
            ref = gc.lin_update(torch.cat((aggr, xg), -1))
            assert (gc._gate_value(qe) - gate).abs().max().item() <= 1e-6
        assert (got - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
