"""CPU: the PyG-shaped containers and NetworkxToHetero of desco_b200 against the reference's own transforms.py run on the
PyG stand-in (oracle/ref_shim) - builder container only (the reference tree is absent on the GPU box; the GPU suite pins
the same surface through tests/golden/shmp_pipeline_ref.npz)."""
import os

import networkx as nx
import numpy as np
import pytest
import torch

from desco_b200 import hetero as H
from desco_b200.graph import gen_enzymes_shaped
from desco_b200.transforms import NetworkxToHetero
from oracle import partition as P

needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")


def _neighborhoods(n=40):
    csr = gen_enzymes_shaped(seed=3, num_graphs=3)
    b = P.partition_dataset(csr, 3)
    return P.neighborhoods_as_networkx(b)[:n]


@needs_reference
def test_networkx_to_hetero_equals_the_reference_transform():
    from oracle import ref_shim

    _, tr = ref_shim.install_full()
    graphs = _neighborhoods() + [nx.graph_atlas(i) for i in (6, 7, 31, 52)]  # typed neighborhoods and untyped queries
    for g in graphs:
        ref = tr.NetworkxToHetero(g.copy(), type_key="type", feat_key="feat")
        got = NetworkxToHetero(g.copy(), type_key="type", feat_key="feat")
        assert got.metadata() == ref.metadata()
        for t in ref.metadata()[0]:
            assert torch.equal(got[t].node_feature, ref[t].node_feature)
            assert got[t].num_nodes == ref[t].num_nodes
        for et in ref.metadata()[1]:
            assert torch.equal(got[et].edge_index, ref[et].edge_index)


def test_networkx_to_hetero_keeps_features_and_extra_attributes():
    g = nx.path_graph(4)
    for u in g.nodes:
        g.nodes[u]["type"] = "canonical" if u == 3 else "count"
        g.nodes[u]["feat"] = torch.tensor([float(u), 1.0])
        g.nodes[u]["label"] = u * 10
    d = NetworkxToHetero(g)
    assert d.metadata()[0] == ["count", "canonical"]
    assert torch.equal(d["count"].node_feature, torch.tensor([[0.0, 1.0], [1.0, 1.0], [2.0, 1.0]]))
    assert torch.equal(d["canonical"].node_feature, torch.tensor([[3.0, 1.0]]))
    assert d["count"].label.view(-1).tolist() == [0, 10, 20]
    assert d["count", "union", "canonical"].edge_index.tolist() == [[2], [0]]
    assert set(d.node_feature_dict) == {"count", "canonical"} and len(d.edge_index_dict) == 3


@needs_reference
def test_batch_collate_equals_the_stand_in_of_pyg_collate():
    from oracle import ref_shim

    datas = [NetworkxToHetero(g) for g in _neighborhoods(12)]
    for d in datas:  # workload.py:275-282: pad missing edge types
        for et in [("count", "union", "count"), ("count", "union", "canonical"), ("canonical", "union", "count")]:
            if et not in d.metadata()[1]:
                d[et].edge_index = torch.empty((2, 0), dtype=torch.long)
    got = H.Batch.from_data_list(datas)
    shim = []
    for d in datas:
        s = ref_shim.HeteroData()
        for t in d.metadata()[0]:
            s[t].node_feature = d[t].node_feature
        for et in d.metadata()[1]:
            s[et].edge_index = d[et].edge_index
        shim.append(s)
    ref = ref_shim.Batch.from_data_list(shim)
    assert got.metadata() == ref.metadata() and got.num_graphs == 12
    for t in ref.metadata()[0]:
        assert torch.equal(got[t].batch, ref[t].batch) and torch.equal(got[t].node_feature, ref[t].node_feature)
    for et in ref.metadata()[1]:
        assert torch.equal(got[et].edge_index, ref[et].edge_index)


def test_homogeneous_edges_offsets_follow_store_order():
    d = H.HeteroData()
    d["canonical"].node_feature = torch.zeros(1, 1)  # canonical store first: offsets must follow the store order
    d["count"].node_feature = torch.zeros(3, 1)
    d["count", "union", "canonical"].edge_index = torch.tensor([[2], [0]])
    d["canonical", "union", "count"].edge_index = torch.tensor([[0], [2]])
    ei, ns, es = H.homogeneous_edges(d)
    assert ns == {"canonical": (0, 1), "count": (1, 4)}
    assert ei.tolist() == [[3, 0], [0, 3]] and es[("canonical", "union", "count")] == (1, 2)
