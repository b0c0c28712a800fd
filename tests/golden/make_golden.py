"""Generate the committed golden fixtures from the REFERENCE ITSELF (run in the builder container only).

    python tests/golden/make_golden.py

* ``partition_*.npz``  : the reference's own ``get_neigh_hetero`` / ``get_neigh_canonical`` (executed from
  ``/root/reference/subgraph_counting/data.py:329-396``) driven like ``NeighborhoodDataset.process``
  (``workload.py:243-260``) on seeded synthetic targets + the KAT graphs of SURVEY.md App. C; edge types from the
  literal sparse ``A*A@A+A > 1`` of ``transforms.py:201-225``.
* ``gossip_ref.npz``   : the reference's own ``BaseGNN`` (GOSSIP path, ``gnn_model.py``) imported unmodified on the
  PyG stand-in of ``oracle/ref_shim``; weights = ``torch.manual_seed`` default init, stored by seed + checksum.
* ``sage_homog_ref.npz``: the reference's own homogeneous SAGE ``BaseGNN`` (BaseGNNCore.forward + anchor_mlp +
  global_add_pool + post_mp) on the shim, on real canonical neighborhoods (centre marked by node_feature = 1).
* ``shmp_hetero_ref.npz``: hetero SHMP forward where every leaf module is the reference's own ``SAGEConv`` /
  ``nn.Linear`` and the ``to_hetero`` wiring is restated (kept as a second, independent pin).
* ``shmp_pipeline_ref.npz``: the reference's OWN pipeline end to end on the extended PyG stand-in
  (``oracle/ref_shim.install_full``): ``get_neigh_hetero`` -> ``NetworkxToHetero`` -> ``ToTconvHetero``
  (``transforms.py:180-255,319-412``, imported unmodified) -> collate in batches -> ``BaseGNN.forward`` of
  ``gnn_model.py`` with ``gnn_core`` converted by the reference's ``to_hetero_old`` (``lightning_model.py:371-421``,
  executed from source; ``pyg.nn.to_hetero`` = a torch.fx trace of the reference's ``BaseGNNCore.forward`` re-executed
  per type) -> the reference's ``graph_to_count`` (``lightning_model.py:198-222``).  Also holds the SHMP edge types the
  reference's ``ToTconvHetero`` assigned, mapped back to the packed edge order.
* ``partition_*.npz`` edge types (``hetero_d*_edge_tri``) come from the reference's ``ToTconvHetero`` itself as well.

The fixtures travel to the GPU box; ``/root/reference`` does not.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import networkx as nx
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from desco_b200.graph import csr_from_networkx, gen_enzymes_shaped, gen_imdb_shaped, gen_mutag_shaped  # noqa: E402
from oracle import model as M  # noqa: E402
from oracle import partition as P  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.shmp_types import type_batch  # noqa: E402


def kat_graphs():
    """SURVEY.md App. C known-answer graphs (node ids made contiguous 0..n-1, order preserved)."""
    gs = []
    for edges in (
        [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4), (3, 5), (3, 6), (5, 6)],  # workflow figure
        [(4, 9), (9, 3), (3, 8), (8, 2), (2, 4)],  # cycle: component step drops 3
        [(5, 7), (7, 1), (5, 4)],  # G3
        [(9, 10), (10, 0), (9, 1), (1, 2), (2, 3), (3, 4), (4, 0)],  # G4: hetero != canonical at k=3
    ):
        g = nx.Graph()
        g.add_nodes_from(range(max(max(e) for e in edges) + 1))
        g.add_edges_from(edges)
        gs.append(g)
    return gs


def reference_typed_neighborhoods(csr, depth, ref_funcs, tr, centres=None):
    """``NeighborhoodDataset.process`` + ``__getitem__`` of the reference on the stand-in (``workload.py:243-290``):
    reference ``get_neigh_hetero`` per node, drop edge-free neighborhoods, reference ``NetworkxToHetero``, pad missing
    edge types, reference ``ToTconvHetero``.  Returns (packed batch with ``edge_tri`` FROM the reference transform,
    list of transformed HeteroData in dataset order, per-neighborhood first-count-node info)."""
    b = P.partition_dataset(csr, depth, mode="hetero", funcs=ref_funcs, with_types=False, centres=centres)
    get_neigh = ref_funcs["get_neigh_hetero"]
    datas, first_is_lowest = [], []
    edge_tri = np.zeros(len(b["edge_col"]), dtype=np.uint8)
    seen = np.zeros(len(b["edge_col"]), dtype=bool)
    cache = {}
    for g, (gid, local) in enumerate(b["index"]):
        gid, local = int(gid), int(local)
        if gid not in cache:
            cache[gid] = csr.to_networkx(gid)
        neigh = get_neigh(cache[gid], local, depth)
        order = {"count": [], "canonical": []}
        for u in neigh.to_directed().nodes:  # NetworkxToHetero assigns per-type ids in this iteration order (:342-348)
            order[neigh.nodes[u]["type"]].append(u)
        data = tr.NetworkxToHetero(neigh, type_key="type", feat_key="feat")
        data.y = torch.empty([1], dtype=torch.double).reshape(1, 1)
        datas.append((data, order))
        first_is_lowest.append(order["count"][0] == min(order["count"]))
    all_types = []
    for data, _ in datas:  # workload.py:275-282 (a list instead of a set: the order of padded types is immaterial)
        for et in data.metadata()[1]:
            if et not in all_types:
                all_types.append(et)
    out = []
    for g, (data, order) in enumerate(datas):
        for et in all_types:
            if et not in data.metadata()[1]:
                data[et].edge_index = torch.empty((2, 0), dtype=torch.long)
        data = tr.ToTconvHetero()(data)  # the dataset `transform` (main.py:85)
        lo = int(b["nbh_ptr"][g])
        base = int(csr.graph_ptr[int(b["index"][g][0])])
        row_of = {int(b["node_gid"][r]) - base: r for r in range(lo, int(b["nbh_ptr"][g + 1]))}
        for (s_t, rel, d_t), st in data._edge_store_dict.items():
            tri = 1 if rel.endswith("_triangle") else 0
            assert rel.endswith("_triangle") or rel.endswith("_tride")
            for a, c in st["edge_index"].T.tolist():
                ru, rv = row_of[order[s_t][a]], row_of[order[d_t][c]]
                es = np.arange(b["edge_ptr"][rv], b["edge_ptr"][rv + 1])  # packed rows hold INCOMING edges: dst row rv, col = src
                e = es[b["edge_col"][es] == ru]
                assert len(e) == 1 and not seen[e[0]]
                edge_tri[e[0]] = tri
                seen[e[0]] = True
        out.append(data)
    assert seen.all()
    b["edge_tri"] = edge_tri
    return b, out, np.asarray(first_is_lowest)


def weights_checksum(module) -> float:
    return float(sum(v.double().abs().sum() for v in module.state_dict().values()))


def main():
    ref_funcs = P.load_reference_functions()
    assert ref_funcs is not None, "run this where /root/reference is mounted"
    ref, tr = ref_shim.install_full()
    lm = ref_shim.load_lightning_methods()

    # ---------------- partition ----------------
    sets = {
        "kat": csr_from_networkx(kat_graphs()),
        "mutag24": gen_mutag_shaped(seed=3, num_graphs=24),
        "enzymes12": gen_enzymes_shaped(seed=3, num_graphs=12),
        "imdb6": gen_imdb_shaped(seed=3, num_graphs=6),
    }
    for name, csr in sets.items():
        out = {"rowptr": csr.rowptr, "col": csr.col, "graph_ptr": csr.graph_ptr}
        for depth in (1, 2, 3, 4):
            for mode in ("hetero", "canonical"):
                if mode == "hetero":  # types from the reference's own ToTconvHetero on its own HeteroData
                    b, _, _ = reference_typed_neighborhoods(csr, depth, ref_funcs, tr)
                    assert np.array_equal(b["edge_tri"], type_batch(b)), "oracle/shmp_types.py disagrees with the reference"
                else:  # homogeneous neighborhoods are not typed by the reference pipeline: literal sparse formulation
                    b = P.partition_dataset(csr, depth, mode=mode, funcs=ref_funcs, with_types=False)
                    b["edge_tri"] = type_batch(b)
                for k, v in b.items():
                    out[f"{mode}_d{depth}_{k}"] = v
        np.savez_compressed(os.path.join(HERE, f"partition_{name}.npz"), **out)
        print(name, csr.num_nodes, "nodes ->", int(out["hetero_d4_indicator"].sum()), "neighborhoods at depth 4")

    # ---------------- gossip (reference BaseGNN on the shim) ----------------
    torch.manual_seed(11)
    rg = ref.BaseGNN(1, 64, 1, M.default_gossip_args(), baseline="gossip", emb_channels=64, input_pattern_emb=True).eval()
    csr = gen_mutag_shaped(seed=5, num_graphs=10)
    ei = torch.from_numpy(csr.edge_index())
    N, Q = csr.num_nodes, 7
    g = torch.Generator().manual_seed(12)
    x = torch.floor(torch.exp(torch.randn(N, Q, generator=g)))
    qe = torch.randn(Q, 64, generator=g)
    outs = []
    with torch.no_grad():
        for q in range(Q):
            data = SimpleNamespace(node_feature=x[:, q].view(-1, 1), edge_index=ei, batch=torch.zeros(N, dtype=torch.long))
            outs.append(x[:, q].view(-1, 1) + rg(data, query_emb=qe[q].view(1, -1)))
        gates = torch.stack([c._gate_value(qe) for c in rg.gnn_core.convs], 0)
    np.savez_compressed(
        os.path.join(HERE, "gossip_ref.npz"), rowptr=csr.rowptr, col=csr.col, graph_ptr=csr.graph_ptr, x=x.numpy(),
        query_emb=qe.numpy(), out=torch.cat(outs, -1).numpy(), gates=gates.numpy(), seed=11,
        checksum=weights_checksum(rg), keys=np.array(list(rg.state_dict().keys())),
    )
    print("gossip golden", torch.cat(outs, -1).abs().max().item())

    # ---------------- homogeneous SAGE BaseGNN (reference, end to end on the shim) ----------------
    torch.manual_seed(21)
    rs = ref.BaseGNN(1, 64, 64, M.default_args(use_hetero=False)).eval()
    csr = gen_enzymes_shaped(seed=7, num_graphs=6)
    b = P.partition_dataset(csr, 4, mode="canonical", funcs=ref_funcs, with_types=False)
    V, G = int(b["nbh_ptr"][-1]), len(b["centre"])
    dst = torch.repeat_interleave(torch.arange(V), torch.as_tensor(np.diff(b["edge_ptr"]), dtype=torch.long))
    src = torch.as_tensor(b["edge_col"], dtype=torch.long)
    bvec = torch.repeat_interleave(torch.arange(G), torch.as_tensor(np.diff(b["nbh_ptr"]), dtype=torch.long))
    nf = torch.zeros(V, 1)
    nf[torch.as_tensor(b["nbh_ptr"][1:] - 1, dtype=torch.long)] = 1.0  # get_neigh_canonical marks the centre (data.py:369-371)
    with torch.no_grad():
        out = rs(SimpleNamespace(node_feature=nf, edge_index=torch.stack([src, dst]), batch=bvec))
    np.savez_compressed(
        os.path.join(HERE, "sage_homog_ref.npz"), **{k: v for k, v in b.items()}, out=out.numpy(), seed=21,
        checksum=weights_checksum(rs), keys=np.array(list(rs.state_dict().keys())),
    )
    print("homog SAGE golden", out.abs().max().item())

    # ---------------- hetero SHMP with reference leaf modules ----------------
    torch.manual_seed(31)
    om = M.NeighborhoodCountingModel().eval()  # supplies the weights (seeded default init)
    csr = gen_enzymes_shaped(seed=9, num_graphs=5)
    b = P.partition_dataset(csr, 4, mode="hetero", funcs=ref_funcs, with_types=False)
    b["edge_tri"] = type_batch(b)
    qb = M.query_batch()

    def ref_leaf_forward(base: M.BaseGNN, views):
        """BaseGNNCore.forward statement order (gnn_model.py:230-277) with reference SAGEConv leaves."""
        core = base.gnn_core
        x = {t: core.pre_mp[0][t](torch.zeros(len(views.rows[t]), 1)) for t in views.meta[0]}
        emb = dict(x)
        for l in range(core.layer_num):
            outs = {t: [] for t in views.meta[0]}
            for et in views.meta[1]:
                s, _, d = et
                conv = ref.SAGEConv(64, 64)
                conv.lin.load_state_dict(core.convs[l]["__".join(et)].lin.state_dict())
                xin = x[s] if s == d else (x[s], x[d])  # App. B.2; SAGEConv.forward then runs remove_self_loops
                # on this bipartite edge_index too (gnn_model.py:389-390) - the quirk oracle/model.py documents
                outs[d].append(conv(xin, views.edges[et]))
            nx_ = {}
            for t in views.meta[0]:
                v = outs[t]
                agg = (v[0] + v[1]) + (v[2] + v[3]) if len(v) == 4 else v[0] + v[1]
                nx_[t] = torch.relu(core.updates[l][t](torch.cat((agg, x[t]), 1)))
            x = nx_
            emb = {t: torch.cat((emb[t], x[t]), 1) for t in emb}
        if "canonical" in emb:
            emb["canonical"] = base.anchor_mlp(emb["canonical"])
        cat = torch.cat([emb[t] for t in views.meta[0]], 0)
        bv = torch.cat([views.batch[t] for t in views.meta[0]], 0)
        return base.post_mp(ref_shim.global_add_pool(cat, bv, views.num_graphs))

    with torch.no_grad():
        t_emb = ref_leaf_forward(om.emb_model, M.hetero_views(b, True, self_loop_quirk=False))  # the reference SAGEConv drops them itself
        q_emb = ref_leaf_forward(om.emb_model_query, M.hetero_views(qb, False))
        pred = torch.cat([om.count_model(torch.cat((t_emb, q.expand_as(t_emb)), -1)) for q in q_emb], -1)
        count = 2 ** pred - 1
    np.savez_compressed(
        os.path.join(HERE, "shmp_hetero_ref.npz"), rowptr=csr.rowptr, col=csr.col, graph_ptr=csr.graph_ptr,
        **{f"b_{k}": v for k, v in b.items()}, target_emb=t_emb.numpy(), query_emb=q_emb.numpy(), pred=pred.numpy(),
        count=count.numpy(), seed=31, checksum=weights_checksum(om),
    )
    print("hetero SHMP golden", len(b["centre"]), "neighborhoods, pred range", pred.min().item(), pred.max().item())

    # ---------------- the reference's own pipeline end to end (transforms + to_hetero_old + graph_to_count) ----------------
    torch.manual_seed(41)
    om = M.NeighborhoodCountingModel().eval()  # supplies the weights; loaded STRICTLY into the reference modules below,
    args = M.default_args()                    # which also checks the state-dict key names of SURVEY App. B.3
    self = SimpleNamespace(
        emb_model=ref.BaseGNN(1, 64, 64, args, emb_channels=64), emb_model_query=ref.BaseGNN(1, 64, 64, args, emb_channels=64),
        count_model=torch.nn.Sequential(torch.nn.Linear(128, 256), torch.nn.LeakyReLU(), torch.nn.Linear(256, 1)),
        kwargs={}, device="cpu")
    lm["to_hetero_old"](self, tconv_target=True, tconv_query=True)
    self.emb_model.load_state_dict(om.emb_model.state_dict(), strict=True)
    self.emb_model_query.load_state_dict(om.emb_model_query.state_dict(), strict=True)
    self.count_model.load_state_dict(om.count_model.state_dict(), strict=True)
    self.emb_model.eval(), self.emb_model_query.eval()
    self.embed_to_count = lambda embs: lm["embed_to_count"](self, embs)
    queries = [tr.ToTconvHetero()(tr.NetworkxToHetero(nx.graph_atlas(i), type_key="type", feat_key="feat"))
               for i in M.STANDARD_QUERY_IDS]  # gen_queries, lightning_model.py:37-87 (+ set_queries' DataLoader(batch_size=64))
    self.query_loader = [ref_shim.Batch.from_data_list(queries)]
    csr = gen_enzymes_shaped(seed=14, num_graphs=8)  # (seed 13: the first neighborhood, {5, 8}, iterates [8, 5] - the F10 trap)
    pyg_batch = 64  # several collated batches: the remove_self_loops quirk is evaluated per batch
    b, datas, first_low = reference_typed_neighborhoods(csr, 4, ref_funcs, tr)
    store_order = datas[0].metadata()[0]
    assert store_order == ["count", "canonical"], "first neighborhood iterates its canonical node first (SURVEY F10): pick another seed"
    for d in datas:  # InMemoryDataset.collate/separate rebuild every element in the FIRST element's store order
        for t in store_order:
            d._node_store_dict.move_to_end(t)
    counts = []
    with torch.no_grad():
        for i in range(0, len(datas), pyg_batch):
            counts.append(lm["graph_to_count"](self, ref_shim.Batch.from_data_list(datas[i:i + pyg_batch])))
        q_emb = torch.cat([self.emb_model_query(qb_) for qb_ in self.query_loader], 0)
    counts = torch.cat(counts, 0)
    np.savez_compressed(
        os.path.join(HERE, "shmp_pipeline_ref.npz"), rowptr=csr.rowptr, col=csr.col, graph_ptr=csr.graph_ptr,
        **{f"b_{k}": v for k, v in b.items()}, count=counts.numpy(), query_emb=q_emb.numpy(), seed=41, pyg_batch_size=pyg_batch,
        checksum=weights_checksum(om), first_count_node_is_lowest_id=first_low,
        keys=np.array(list(self.emb_model.state_dict().keys())),
    )
    print("pipeline golden", len(b["centre"]), "neighborhoods; first count node in networkx order is the lowest id in",
          int(first_low.sum()), "of", len(first_low), "; count range", counts.min().item(), counts.max().item())

    # ---------------- ground-truth canonical counts: the reference's own VF2 worker + symmetry factor ----------------
    from oracle import groundtruth as GT

    gt_funcs = GT.reference_functions()
    queries = [nx.graph_atlas(i) for i in M.STANDARD_QUERY_IDS]
    dense = [nx.gnm_random_graph(12, 34, seed=3), nx.complete_graph(6), nx.star_graph(7), nx.cycle_graph(9)]
    gsets = {"kat": csr_from_networkx(kat_graphs() + dense), "mutag": gen_mutag_shaped(seed=8, num_graphs=10),
             "enzymes": gen_enzymes_shaped(seed=8, num_graphs=3), "imdb": gen_imdb_shaped(seed=8, num_graphs=2)}
    out = {}
    for name, csr in gsets.items():
        truth = GT.canonical_count_truth(csr, queries, gt_funcs)
        assert np.array_equal(truth, np.round(truth))
        out.update({f"{name}_rowptr": csr.rowptr, f"{name}_col": csr.col, f"{name}_graph_ptr": csr.graph_ptr,
                    f"{name}_truth": truth.astype(np.int64)})
        print("ground truth golden", name, csr.num_nodes, "nodes, total occurrences", int(truth.sum()))
    np.savez_compressed(os.path.join(HERE, "groundtruth_ref.npz"), query_ids=np.asarray(M.STANDARD_QUERY_IDS), **out)


if __name__ == "__main__":
    main()
