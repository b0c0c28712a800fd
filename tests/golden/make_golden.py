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
  ``nn.Linear`` and only the ``to_hetero`` wiring (which cannot be run without PyG) is restated.

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


def weights_checksum(module) -> float:
    return float(sum(v.double().abs().sum() for v in module.state_dict().values()))


def main():
    ref_funcs = P.load_reference_functions()
    assert ref_funcs is not None, "run this where /root/reference is mounted"
    ref = ref_shim.install()

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
                b = P.partition_dataset(csr, depth, mode=mode, funcs=ref_funcs, with_types=False)
                b["edge_tri"] = type_batch(b)  # literal sparse formulation
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


if __name__ == "__main__":
    main()
