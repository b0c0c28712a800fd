"""CPU: the oracle against the committed golden fixtures (generated from the reference, tests/golden/make_golden.py)
and, when /root/reference is mounted, against the reference executed live."""
import os

import networkx as nx
import numpy as np
import pytest
import torch

from desco_b200.graph import TargetCSR, csr_from_networkx, gen_mutag_shaped
from oracle import model as M
from oracle import partition as P
from oracle.shmp_types import type_batch

KEYS = ("nbh_ptr", "node_gid", "edge_ptr", "edge_col", "edge_tri", "centre", "index", "indicator")


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name), allow_pickle=False)
    return z, TargetCSR(z["rowptr"], z["col"], z["graph_ptr"])


@pytest.mark.parametrize("name", ["kat", "mutag24", "enzymes12", "imdb6"])
@pytest.mark.parametrize("mode", ["hetero", "canonical"])
def test_partition_oracle_matches_reference_golden(golden_dir, name, mode):
    z, csr = _load(golden_dir, f"partition_{name}.npz")
    for depth in (1, 2, 3, 4):
        if name != "kat" and depth in (1, 3) and mode == "canonical":
            continue  # keep the CPU suite short
        b = P.partition_dataset(csr, depth, mode=mode)
        for k in KEYS:
            assert np.array_equal(b[k], z[f"{mode}_d{depth}_{k}"]), (name, mode, depth, k)


def test_partition_known_answers():
    """SURVEY.md App. C KATs (reference get_neigh_* executed from source gave exactly these)."""
    G4 = nx.Graph([(9, 10), (10, 0), (9, 1), (1, 2), (2, 3), (3, 4), (4, 0)])
    assert sorted(P.get_neigh_hetero(G4, 9, 2).nodes) == [1, 2, 9]
    assert sorted(P.get_neigh_hetero(G4, 9, 3).nodes) == [0, 1, 2, 3, 4, 9]
    assert sorted(P.get_neigh_canonical(G4, 9, 3).nodes) == [1, 2, 3, 9]
    cyc = nx.Graph([(4, 9), (9, 3), (3, 8), (8, 2), (2, 4)])
    for k in (1, 2, 3, 4):
        assert sorted(P.get_neigh_hetero(cyc, 4, k).nodes) == [2, 4]
    G3 = nx.Graph([(5, 7), (7, 1), (5, 4)])
    assert sorted(P.get_neigh_hetero(G3, 5, 4).nodes) == [4, 5]
    wf = nx.Graph([(0, 1), (0, 2), (1, 2), (2, 3), (3, 4), (3, 5), (3, 6), (5, 6)])
    for v in range(7):
        assert sorted(P.get_neigh_hetero(wf, v, 4).nodes) == list(range(v + 1))
    tri = {tuple(sorted(e)) for e in wf.edges if P.edge_is_triangle(wf, *e)}
    assert tri == {(0, 1), (0, 2), (1, 2), (3, 5), (3, 6), (5, 6)}


def test_partition_invariants():
    csr = gen_mutag_shaped(seed=1, num_graphs=12)
    b = P.partition_dataset(csr, 4)
    assert b["indicator"].sum() == len(b["centre"])
    assert len(set(b["centre"].tolist())) == len(b["centre"])  # every node is the max of at most one neighborhood
    last = b["node_gid"][b["nbh_ptr"][1:] - 1]
    assert np.array_equal(last, b["centre"])
    # tri flags symmetric under edge reversal
    V = int(b["nbh_ptr"][-1])
    dst = np.repeat(np.arange(V), np.diff(b["edge_ptr"]))
    fwd = {(int(d), int(s)): int(t) for d, s, t in zip(dst, b["edge_col"], b["edge_tri"])}
    assert all(fwd[(s, d)] == t for (d, s), t in fwd.items())


def test_typing_literal_sparse_equals_set_rule(golden_dir):
    z, csr = _load(golden_dir, "partition_imdb6.npz")
    b = P.partition_dataset(csr, 4)  # set rule
    assert np.array_equal(type_batch(b), b["edge_tri"])  # literal A*A@A+A > 1
    assert 0.3 < b["edge_tri"].mean() <= 1.0


@pytest.mark.skipif(not os.path.exists(P.REFERENCE_DATA_PY), reason="reference tree not mounted")
def test_partition_oracle_matches_reference_live():
    ref = P.load_reference_functions()
    csr = gen_mutag_shaped(seed=2, num_graphs=10)
    for mode in ("hetero", "canonical"):
        for depth in (2, 4):
            a = P.partition_dataset(csr, depth, mode=mode)
            b = P.partition_dataset(csr, depth, mode=mode, funcs=ref)
            for k in KEYS:
                assert np.array_equal(a[k], b[k])


def test_query_ids():
    assert M.gen_query_ids((3, 4, 5)) == M.STANDARD_QUERY_IDS
    qb = M.query_batch()
    assert len(qb["nbh_ptr"]) - 1 == 29


def test_gossip_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "gossip_ref.npz"))
    torch.manual_seed(int(z["seed"]))
    # same construction order as ref.BaseGNN: gnn_core (pre_mp, convs) then anchor_mlp then post_mp
    om = M.GossipCountingModel()
    ck = float(sum(v.double().abs().sum() for v in om.emb_model.state_dict().values()))
    assert abs(ck - float(z["checksum"])) < 1e-6 * abs(ck), "seeded init differs from the fixture's"
    assert list(om.emb_model.state_dict().keys()) == list(z["keys"])
    csr = TargetCSR(z["rowptr"], z["col"], z["graph_ptr"])
    om.set_query_emb(torch.from_numpy(z["query_emb"]))
    with torch.no_grad():
        out = om.graph_to_count(torch.from_numpy(z["x"]), torch.from_numpy(csr.edge_index()))
        gates = om.gate_value(torch.from_numpy(z["query_emb"]))
    assert torch.allclose(out, torch.from_numpy(z["out"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(gates, torch.from_numpy(z["gates"]), rtol=1e-6, atol=1e-7)


def test_shmp_oracle_matches_reference_leaf_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "shmp_hetero_ref.npz"))
    torch.manual_seed(int(z["seed"]))
    om = M.NeighborhoodCountingModel().eval()
    ck = float(sum(v.double().abs().sum() for v in om.state_dict().values()))
    assert abs(ck - float(z["checksum"])) < 1e-6 * abs(ck)
    b = {k[2:]: z[k] for k in z.files if k.startswith("b_")}
    with torch.no_grad():
        pred = om.pre_exponent(b, M.query_batch())
        q = om.get_query_emb(M.query_batch())
    assert torch.allclose(q, torch.from_numpy(z["query_emb"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(pred, torch.from_numpy(z["pred"]), rtol=1e-5, atol=1e-6)


def test_homog_sage_oracle_matches_reference_golden(golden_dir):
    """Reference BaseGNN (homogeneous SAGE, end to end on the PyG stand-in) == oracle core with one node type."""
    z = np.load(os.path.join(golden_dir, "sage_homog_ref.npz"))
    torch.manual_seed(int(z["seed"]))
    meta = (["n"], [("n", "r", "n")])
    oc = M.BaseGNN(1, 64, 64, M.default_args(use_hetero=False), meta).eval()
    ck = float(sum(v.double().abs().sum() for v in oc.state_dict().values()))
    assert abs(ck - float(z["checksum"])) < 1e-6 * abs(ck)
    V, G = int(z["nbh_ptr"][-1]), len(z["centre"])
    dst = torch.repeat_interleave(torch.arange(V), torch.as_tensor(np.diff(z["edge_ptr"]), dtype=torch.long))
    src = torch.as_tensor(z["edge_col"], dtype=torch.long)
    bvec = torch.repeat_interleave(torch.arange(G), torch.as_tensor(np.diff(z["nbh_ptr"]), dtype=torch.long))
    nf = torch.zeros(V, 1)
    centre_rows = torch.as_tensor(z["nbh_ptr"][1:] - 1, dtype=torch.long)
    nf[centre_rows] = 1.0
    from types import SimpleNamespace

    views = SimpleNamespace(rows={"n": torch.arange(V)}, batch={"n": bvec}, edges={("n", "r", "n"): torch.stack([src, dst])},
                            num_graphs=G, meta=meta)
    with torch.no_grad():
        emb = oc.gnn_core({"n": nf}, views.edges)["n"]
        emb[centre_rows] = oc.anchor_mlp(emb[centre_rows])  # gnn_model.py:77-83
        pooled = torch.zeros(G, emb.shape[1]).index_add_(0, bvec, emb)
        out = oc.post_mp(pooled)
    assert torch.allclose(out, torch.from_numpy(z["out"]), rtol=1e-5, atol=1e-6)


def test_large_graph_views_agree_with_the_whole_graph_oracle():
    """oracle/large.py (ball view for the partition, 2-hop closure for gossip) == the oracle on the whole graph."""
    import torch

    from desco_b200.graph import gen_powerlaw
    from oracle import model as M
    from oracle.large import BallView, gossip_closure

    csr = gen_powerlaw(1500, 6000, seed=2)
    centres = np.array([3, 700, 1499, 42])
    full = P.partition_dataset(csr, 2, mode="hetero", centres=centres)
    view = P.partition_dataset(BallView(csr.rowptr, csr.col, centres, 2), 2, mode="hetero", centres=centres)
    for k in full:
        assert np.array_equal(full[k], view[k]), k
    torch.manual_seed(0)
    og = M.GossipCountingModel()
    Q = 3
    x = torch.floor(torch.exp(torch.randn(csr.num_nodes, Q)))
    og.set_query_emb(torch.randn(Q, 64))
    sample = np.array([5, 900, 1499, 0])
    nodes, ei, pos = gossip_closure(csr.rowptr, csr.col, sample)
    with torch.no_grad():
        ref = og.graph_to_count(x, torch.from_numpy(csr.edge_index()))[torch.as_tensor(sample)]
        got = og.graph_to_count(x[torch.as_tensor(nodes)], torch.from_numpy(ei))[torch.as_tensor(pos)]
    assert (ref - got).abs().max().item() <= 1e-5


def test_oracle_matches_the_reference_pipeline_golden(golden_dir):
    """shmp_pipeline_ref.npz was produced by the reference's own get_neigh_hetero -> NetworkxToHetero -> ToTconvHetero ->
    collate -> to_hetero_old'd BaseGNN -> graph_to_count on the PyG stand-in (tests/golden/make_golden.py): pins the
    to_hetero wiring, the SHMP typing and the remove_self_loops quirk of oracle/model.py + oracle/shmp_types.py."""
    z = np.load(os.path.join(golden_dir, "shmp_pipeline_ref.npz"))
    b = {k[2:]: z[k] for k in z.files if k.startswith("b_")}
    assert np.array_equal(type_batch(b), b["edge_tri"])  # the reference's ToTconvHetero == the literal formulation
    torch.manual_seed(int(z["seed"]))
    om = M.NeighborhoodCountingModel().eval()
    assert abs(sum(float(v.double().abs().sum()) for v in om.state_dict().values()) - float(z["checksum"])) < 1e-6 * float(z["checksum"])
    qb = M.query_batch()
    with torch.no_grad():
        c = om.graph_to_count(b, qb, pyg_batch_size=int(z["pyg_batch_size"]))
        qe = om.get_query_emb(qb)
        c_off = om.graph_to_count(b, qb, pyg_batch_size=int(z["pyg_batch_size"]), self_loop_quirk=False)
    assert (qe - torch.from_numpy(z["query_emb"])).abs().max().item() <= 1e-6
    assert (c - torch.from_numpy(z["count"])).abs().max().item() <= 1e-6
    assert (c_off - torch.from_numpy(z["count"])).abs().max().item() > 1e-5  # the fixture does exercise the quirk
    # state-dict keys of the reference's hetero modules (PyG to_hetero naming) == the oracle's == the product's
    assert sorted(z["keys"].tolist()) == sorted(om.emb_model.state_dict().keys())
