"""GPU: the counting driver end to end (Workload -> canonical partition -> SHMP counts -> gossip -> graph-level sums)
against the same sequence run through the oracle, on the MUTAG-shaped config (BASELINE.json configs[0])."""
import numpy as np
import pytest
import torch

from desco_b200.graph import gen_mutag_shaped

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a - b).abs() / b.abs().clamp(min=1.0)).max().item()


def _models(seed):
    from desco_b200.lightning_model import GossipCountingModel, NeighborhoodCountingModel, STANDARD_QUERY_IDS
    from oracle import model as M

    torch.manual_seed(seed)
    om, og = M.NeighborhoodCountingModel().eval(), M.GossipCountingModel()
    pm, pg = NeighborhoodCountingModel().eval(), GossipCountingModel()
    pm.load_state_dict(om.state_dict())
    pg.emb_model.load_state_dict(og.emb_model.state_dict())
    pm, pg = pm.cuda(), pg.cuda()
    pm.set_queries(STANDARD_QUERY_IDS)
    return om, og, pm, pg


def test_pipeline_matches_oracle_mutag_shaped(cuda_device, tmp_path):
    from desco_b200.workload import Workload, count_subgraphs
    from oracle import model as M
    from oracle import partition as P

    csr = gen_mutag_shaped(seed=0, num_graphs=60)
    om, og, pm, pg = _models(3)
    # ---- oracle: workload.py:250-260 loop, lightning_model.py:198-222, workload.py:107-112, :613-628, :136-148 ----
    b = P.partition_dataset(csr, 4)
    qb = M.query_batch()
    with torch.no_grad():
        ref_counts = om.graph_to_count(b, qb, pyg_batch_size=512)
        x = torch.zeros(csr.num_nodes, ref_counts.shape[1])
        x[torch.as_tensor(b["indicator"])] = ref_counts
        og.set_query_emb(om.get_query_emb(qb))
        ref_nodes = og.graph_to_count(x, torch.from_numpy(csr.edge_index()))
    gid = torch.as_tensor(csr.graph_of(np.arange(csr.num_nodes)))
    ref_graph = torch.zeros(csr.num_graphs, x.shape[1]).index_add_(0, gid, ref_nodes)
    # ---- product ----
    wl = Workload(csr, str(tmp_path))
    counts, nodes, graphs = count_subgraphs(wl, pm, pg, depth=4, batch_size=512)
    torch.cuda.synchronize()
    pm.emb_model.check_status()
    nd = wl.neighborhood_dataset
    assert np.array_equal(nd.nx_neighs_index, b["index"]) and np.array_equal(nd.nx_neighs_indicator, b["indicator"])
    assert (tmp_path / "NeighborhoodDataset" / "processed" / "neighs_index_depth_4.npy").exists()
    assert _rel(counts.cpu(), ref_counts) <= 1e-4
    assert _rel(nodes.cpu(), ref_nodes) <= 1e-4
    assert graphs.shape == (60, 29) and _rel(graphs.cpu(), ref_graph) <= 1e-4
    # neighborhood-level aggregation (workload.py:303-324)
    agg = nd.aggregate_neighborhood_count(counts)
    ref_agg = torch.zeros(60, 29).index_add_(0, torch.as_tensor(b["index"][:, 0]), ref_counts)
    assert _rel(agg, ref_agg) <= 1e-4


def test_loader_batches_equal_whole_dataset(cuda_device):
    """DataLoader-style chunks of 512 neighborhoods (config.py:255) give the same counts as the one-pass batch when the
    reference quirk is evaluated per chunk either way."""
    from desco_b200.workload import Workload

    _, _, pm, _ = _models(4)
    wl = Workload(gen_mutag_shaped(seed=1, num_graphs=80), None)
    wl.generate_pipeline_datasets(depth_neigh=4)
    nd = wl.neighborhood_dataset
    pm.set_pyg_batch_size(512)
    with torch.no_grad():
        whole = pm.graph_to_count(nd.batch)
        parts = torch.cat([pm.graph_to_count(bt) for bt in nd.loader(512)], 0)
    assert len(nd) > 512 and parts.shape == whole.shape
    assert (parts - whole).abs().max().item() <= 1e-6


def test_shuffled_loader_and_select(cuda_device):
    """LightningDataLoader(shuffle=True) (lightning_data.py:59-100): every neighborhood exactly once per pass, in a fresh
    seeded permutation; a gathered chunk (NeighborhoodBatch.select) is the same packed structure a contiguous slice gives
    and counts to the same values as the rows of the one-pass batch; the truth counts ride along as batch.y."""
    from desco_b200.lightning_data import LightningDataLoader
    from desco_b200.workload import Workload

    _, _, pm, _ = _models(4)
    wl = Workload(gen_mutag_shaped(seed=2, num_graphs=60), None)
    wl.generate_pipeline_datasets(depth_neigh=4)
    nd = wl.neighborhood_dataset
    G = len(nd)
    nd.y = torch.arange(G * 3, dtype=torch.float32, device="cuda").reshape(G, 3)
    pm.set_pyg_batch_size(-1)  # the reference quirk depends on the collated order: off for an order-free comparison
    with torch.no_grad():
        whole = pm.graph_to_count(nd.batch)
    # select on a contiguous range == slice
    a, b = nd.batch.slice(100, 300).to_numpy(), nd.batch.select(torch.arange(100, 300)).to_numpy()
    for k in ("nbh_ptr", "node_gid", "edge_ptr", "edge_col", "edge_tri", "centre"):
        assert np.array_equal(a[k], b[k]), k
    gen = torch.Generator().manual_seed(7)
    dl = LightningDataLoader(train_dataset=nd, batch_size=256, shuffle=True, generator=gen)
    seen, passes = [], []
    for _ in range(2):
        order = []
        for bt in dl.train_dataloader():
            idx = bt.index_in_dataset
            assert torch.equal(bt.y, nd.y[idx]) and torch.equal(bt.centre, nd.batch.centre[idx])
            with torch.no_grad():
                got = pm.graph_to_count(bt)
            assert _rel(got, whole[idx]) <= 1e-5
            order.append(idx)
        order = torch.cat(order)
        assert torch.equal(torch.sort(order).values, torch.arange(G, device=order.device))
        passes.append(order)
    assert not torch.equal(passes[0], passes[1]) and not torch.equal(passes[0], torch.arange(G, device="cuda"))
    # in-order loader hands out y as well
    ys = torch.cat([bt.y for bt in LightningDataLoader(test_dataset=nd, batch_size=256).test_dataloader()])
    assert torch.equal(ys, nd.y)


def test_sharded_pipeline_single_rank_equals_direct(cuda_device):
    """distributed.ShardedPipeline with world size 1 (no process group): same numbers as the direct calls, and the
    node-range gossip entry with several ranges stitched by hand equals the single call."""
    from types import SimpleNamespace

    from desco_b200.data import DeviceCSR, partition_batch
    from desco_b200.distributed import ShardedPipeline, node_ranges

    _, _, pm, pg = _models(5)
    csr = gen_mutag_shaped(seed=2, num_graphs=40)
    g = DeviceCSR.from_host(csr)
    sp = ShardedPipeline(g, pm, pg, depth=4)
    centres, counts = sp.count_neighborhoods()
    with torch.no_grad():
        direct = pm.graph_to_count(partition_batch(g, None, 4))
    assert torch.equal(counts, direct)
    x = sp.gather_node_counts(centres, counts)
    qe = pm.get_query_emb()
    out = sp.gossip(x, qe)
    pg.set_query_emb(qe)
    with torch.no_grad():
        ref = pg.graph_to_count(SimpleNamespace(graph=g, x=x))
    assert torch.equal(out, ref)
    # three "ranks" emulated in one process (LocalComm): every rank runs the pipelined node-range forward over query
    # groups; the in-place block exchange stands in for the NCCL all-gathers
    from desco_b200.distributed import LocalComm
    from desco_b200.gnn_model import GossipShardedRun

    for qg in (4, 29, 5):
        comm = LocalComm(3)
        runs = [GossipShardedRun(pg.emb_model, g.rowptr, g.col, x, qe, comm.for_rank(r), query_group=qg) for r in range(3)]
        for r in runs:
            r.start()
        for r in runs:
            r.finish()
        for r in runs:
            assert torch.equal(r.result(), ref)
    comm = LocalComm(2)
    runs = [GossipShardedRun(pg.emb_model, g.rowptr, g.col, x, qe, comm.for_rank(r), gather_output=False).start() for r in range(2)]
    own = torch.cat([r.finish().result() for r in runs], 0)
    assert torch.equal(own, ref)
