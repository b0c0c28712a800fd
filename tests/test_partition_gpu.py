"""GPU parity: canonical partition + SHMP typing kernels (through the C ABI) vs the oracle and the golden fixtures."""
import os

import numpy as np
import pytest
import torch

from desco_b200.graph import (TargetCSR, csr_from_networkx, first_nonempty_centres, gen_cox2_shaped,
                              gen_enzymes_shaped, gen_imdb_shaped, gen_mutag_shaped, gen_syn1827_shaped)

pytestmark = pytest.mark.gpu

KEYS = ("nbh_ptr", "node_gid", "edge_ptr", "edge_col", "edge_tri", "centre", "index", "indicator")


def _gpu_partition(csr, depth, mode, centres=None, large=None):
    from desco_b200.data import DeviceCSR, partition_batch

    d = DeviceCSR.from_host(csr)
    c = None if centres is None else torch.as_tensor(centres, dtype=torch.int32, device=d.rowptr.device)
    return partition_batch(d, c, depth, mode, large=large).to_numpy()


DEFAULT_CAPS = (13, 5120, 4096, 19, 1 << 18, 1 << 18)


@pytest.fixture
def large_caps():
    """Shrink / restore the tier capacities of the large-graph partition (so small graphs exercise every tier)."""
    from desco_b200 import _lib

    lib = _lib.load()

    def set_caps(*caps):
        _lib.check(lib.desco_partition_large_set_caps(*caps), "desco_partition_large_set_caps")

    yield set_caps
    set_caps(*DEFAULT_CAPS)


@pytest.mark.parametrize("name", ["kat", "mutag24", "enzymes12", "imdb6"])
@pytest.mark.parametrize("mode", ["hetero", "canonical"])
def test_partition_matches_reference_golden(cuda_device, golden_dir, name, mode):
    z = np.load(os.path.join(golden_dir, f"partition_{name}.npz"))
    csr = TargetCSR(z["rowptr"], z["col"], z["graph_ptr"])
    for depth in (1, 2, 3, 4):
        b = _gpu_partition(csr, depth, mode)
        for k in KEYS:
            assert np.array_equal(b[k], z[f"{mode}_d{depth}_{k}"]), (name, mode, depth, k)


@pytest.mark.parametrize("gen,kw", [(gen_mutag_shaped, {}), (gen_cox2_shaped, dict(num_graphs=60)),
                                    (gen_enzymes_shaped, dict(num_graphs=80)), (gen_imdb_shaped, dict(num_graphs=40)),
                                    (gen_syn1827_shaped, dict(stride=160))])
def test_partition_matches_oracle_on_config_shapes(cuda_device, gen, kw):
    from oracle import partition as P

    csr = gen(seed=0, **kw)
    ref = P.partition_dataset(csr, 4, mode="hetero")
    b = _gpu_partition(csr, 4, "hetero")
    for k in KEYS:
        assert np.array_equal(b[k], ref[k]), (gen.__name__, k)


def test_partition_cta_regime_large_graph(cuda_device):
    """A single graph above 2048 nodes takes the CTA-per-centre kernel; compare a centre sample with the oracle."""
    from oracle import partition as P

    from desco_b200.graph import gen_powerlaw

    csr = gen_powerlaw(6000, 24000, seed=1)
    rng = np.random.default_rng(0)
    centres = np.sort(rng.choice(csr.num_nodes, size=96, replace=False)).astype(np.int32)
    for depth in (2, 3):
        ref = P.partition_dataset(csr, depth, mode="hetero", centres=centres)
        b = _gpu_partition(csr, depth, "hetero", centres)
        for k in KEYS:
            assert np.array_equal(b[k], ref[k]), (depth, k)


def test_partition_edge_cases(cuda_device):
    import networkx as nx

    # isolated node, single edge, star, empty centre list, depth 0
    g = nx.Graph()
    g.add_nodes_from(range(6))
    g.add_edges_from([(1, 2), (3, 4), (3, 5)])
    csr = csr_from_networkx([g, nx.path_graph(3)])
    b = _gpu_partition(csr, 4, "hetero")
    assert b["indicator"].tolist() == [False, False, True, False, True, True, False, True, True]
    b0 = _gpu_partition(csr, 0, "hetero")
    assert b0["indicator"].sum() == 0 and b0["nbh_ptr"].tolist() == [0]
    be = _gpu_partition(csr, 4, "hetero", centres=np.zeros(0, dtype=np.int32))
    assert be["nbh_ptr"].tolist() == [0] and len(be["edge_col"]) == 0


def test_partition_properties_full_size(cuda_device):
    """Config-2 size (whole ENZYMES-shaped pool): size-independent invariants."""
    csr = gen_enzymes_shaped(seed=0)
    b = _gpu_partition(csr, 4, "hetero")
    assert np.array_equal(np.nonzero(b["indicator"])[0][:4096], first_nonempty_centres(csr, 4096))
    assert np.array_equal(b["node_gid"][b["nbh_ptr"][1:] - 1], b["centre"])  # canonical = last row = max id
    V = int(b["nbh_ptr"][-1])
    dst = np.repeat(np.arange(V, dtype=np.int64), np.diff(b["edge_ptr"]))
    key = dst * V + b["edge_col"]
    rev = b["edge_col"].astype(np.int64) * V + dst
    order, rorder = np.argsort(key), np.argsort(rev)
    assert np.array_equal(key[order], rev[rorder])  # symmetric edge set
    assert np.array_equal(b["edge_tri"][order], b["edge_tri"][rorder])  # symmetric types
    nb = np.repeat(np.arange(len(b["centre"])), np.diff(b["nbh_ptr"]))
    assert np.array_equal(nb[dst], nb[b["edge_col"]])  # edges never leave their neighborhood
    assert (b["node_gid"][dst] != b["node_gid"][b["edge_col"]]).all()


def test_standalone_edge_typing(cuda_device):
    from desco_b200.data import shmp_edge_types
    from oracle import partition as P

    csr = gen_imdb_shaped(seed=2, num_graphs=20)
    ref = P.partition_dataset(csr, 4)
    tri = shmp_edge_types(torch.as_tensor(ref["edge_ptr"], device="cuda"), torch.as_tensor(ref["edge_col"], device="cuda"))
    assert np.array_equal(tri.cpu().numpy(), ref["edge_tri"])


def test_drop_in_single_centre_api(cuda_device):
    import networkx as nx

    from desco_b200 import data as D
    from oracle import partition as P

    G4 = nx.Graph([(9, 10), (10, 0), (9, 1), (1, 2), (2, 3), (3, 4), (4, 0)])
    for k in (2, 3, 5):
        a, b = D.get_neigh_hetero(G4, 9, k), P.get_neigh_hetero(G4, 9, k)
        assert sorted(a.nodes) == sorted(b.nodes) and {frozenset(e) for e in a.edges} == {frozenset(e) for e in b.edges}
        assert a.nodes[9]["type"] == "canonical" and all(a.nodes[u]["type"] == "count" for u in a.nodes if u != 9)
        a, b = D.get_neigh_canonical(G4, 9, k), P.get_neigh_canonical(G4, 9, k)
        assert sorted(a.nodes) == sorted(b.nodes)
        assert sorted(D.k_neigh(G4, 9, k)) == sorted(P.k_neigh(G4, 9, k))
    assert sorted(D.get_neigh_hetero(G4, 0, 4).nodes) == [0]


# ---------------------------------------------------------------------------------------------
# large-graph regime (config 5): shared-memory hash-set tier + cooperative team tier (global bitmaps)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["kat", "mutag24", "enzymes12", "imdb6"])
@pytest.mark.parametrize("mode", ["hetero", "canonical"])
def test_large_path_matches_reference_golden(cuda_device, golden_dir, name, mode):
    z = np.load(os.path.join(golden_dir, f"partition_{name}.npz"))
    csr = TargetCSR(z["rowptr"], z["col"], z["graph_ptr"])
    for depth in (1, 2, 3, 4):
        b = _gpu_partition(csr, depth, mode, large=True)
        for k in KEYS:
            assert np.array_equal(b[k], z[f"{mode}_d{depth}_{k}"]), (name, mode, depth, k)


@pytest.mark.parametrize("caps,label", [
    (DEFAULT_CAPS, "tier0"),
    ((11, 16, 8, 12, 512, 256), "tier0+team"),     # tier 0 holds 16 members / 8 rows: most centres go to the team tier
    ((11, 1, 1, 9, 24, 16), "team-only"),          # tier 0 holds one member: every centre with a neighbour is a team job
])
def test_large_path_every_tier_matches_oracle(cuda_device, large_caps, caps, label):
    from oracle import partition as P

    large_caps(*caps)
    for gen, kw in ((gen_enzymes_shaped, dict(num_graphs=40)), (gen_imdb_shaped, dict(num_graphs=25))):
        csr = gen(seed=3, **kw)
        for mode in ("hetero", "canonical"):
            ref = P.partition_dataset(csr, 4, mode=mode)
            b = _gpu_partition(csr, 4, mode, large=True)
            for k in KEYS:
                assert np.array_equal(b[k], ref[k]), (label, gen.__name__, mode, k)


def test_large_path_powerlaw_sample_matches_oracle(cuda_device, large_caps):
    """Config-5-shaped target (power law, hubs, random labels), depths 2 and 3, default and shrunken tiers; also the
    plain k-hop ball (k_neigh)."""
    from oracle import partition as P

    from desco_b200.graph import gen_powerlaw

    csr = gen_powerlaw(30000, 150000, seed=2)
    rng = np.random.default_rng(1)
    centres = np.sort(rng.choice(csr.num_nodes, size=64, replace=False)).astype(np.int32)
    for caps in (DEFAULT_CAPS, (12, 512, 256, 15, 16384, 8192)):
        large_caps(*caps)
        for depth in (2, 3):
            ref = P.partition_dataset(csr, depth, mode="hetero", centres=centres)
            b = _gpu_partition(csr, depth, "hetero", centres, large=True)
            for k in KEYS:
                assert np.array_equal(b[k], ref[k]), (caps, depth, k)
    small = _gpu_partition(csr, 2, "khop", centres[:16], large=False)
    big = _gpu_partition(csr, 2, "khop", centres[:16], large=True)
    for k in KEYS:
        assert np.array_equal(small[k], big[k]), k


def test_large_path_equals_bitset_path_full_sweep(cuda_device):
    """Every centre of a 20k-node power-law target, depth 2: the sparse tiers and the shared-memory bitset kernel agree."""
    from desco_b200.graph import gen_powerlaw

    csr = gen_powerlaw(20000, 100000, seed=5)
    a = _gpu_partition(csr, 2, "hetero", large=False)
    b = _gpu_partition(csr, 2, "hetero", large=True)
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("large", [False, True])
def test_count_only_sizes_equal_packed_batch(cuda_device, large):
    """partition_sizes (streaming count-only mode, chunked) = the row / edge counts of the packed neighborhoods; dropped
    edge-free neighborhoods read (0, 0)."""
    from desco_b200.data import DeviceCSR, partition_batch, partition_sizes
    from desco_b200.graph import gen_powerlaw

    csr = gen_powerlaw(6000, 24000, seed=11)
    d = DeviceCSR.from_host(csr)
    nv, ne = partition_sizes(d, None, 2, "hetero", large=large, max_centres=1000)
    b = partition_batch(d, None, 2, "hetero", large=large)
    nv, ne = nv.cpu().numpy(), ne.cpu().numpy()
    got = b.to_numpy()
    kept = got["indicator"]
    want_nv, want_ne = np.zeros_like(nv), np.zeros_like(ne)
    want_nv[kept] = np.diff(got["nbh_ptr"])
    want_ne[kept] = np.diff(got["edge_ptr"][got["nbh_ptr"]])
    assert np.array_equal(nv, want_nv) and np.array_equal(ne, want_ne)
    assert int((nv == 0).sum()) == int((~kept).sum())


def test_one_call_partition_grows_its_buffers(cuda_device):
    """desco_partition_batch reports ENOBUFS with the exact sizes when the caller's capacity is too small; the Python
    wrapper re-allocates once and the result is still bit-exact (here the capacity estimate is forced down to 1 row per
    centre and 1 edge per row on a dense IMDB-shaped set)."""
    from oracle import partition as P

    from desco_b200 import data as D

    csr = gen_imdb_shaped(seed=4, num_graphs=12)
    ref = P.partition_dataset(csr, 4, mode="hetero")
    saved = dict(D._CAPACITY)
    try:
        D._CAPACITY.update(rows_per_centre=0.5, edges_per_row=0.5)
        b = _gpu_partition(csr, 4, "hetero", large=False)
        assert D._CAPACITY["rows_per_centre"] > 1.0  # learnt from the retry
    finally:
        D._CAPACITY.update(saved)
    for k in KEYS:
        assert np.array_equal(b[k], ref[k]), k


def test_one_call_partition_empty_and_edge_free(cuda_device):
    """No centres at all, and centres whose neighborhoods are all edge-free (dropped, workload.py:253-256)."""
    import networkx as nx

    from desco_b200.graph import csr_from_networkx

    g = nx.Graph()
    g.add_nodes_from(range(4))
    g.add_edge(2, 3)
    csr = csr_from_networkx([g])
    b = _gpu_partition(csr, 4, "hetero", centres=np.zeros(0, dtype=np.int32), large=False)
    assert len(b["centre"]) == 0 and b["nbh_ptr"].tolist() == [0] and len(b["edge_col"]) == 0
    b = _gpu_partition(csr, 4, "hetero", centres=np.array([0, 1, 2], dtype=np.int32), large=False)
    assert len(b["centre"]) == 0 and b["indicator"].tolist() == [False, False, False]
    b = _gpu_partition(csr, 4, "hetero", large=False)
    assert b["centre"].tolist() == [3] and b["node_gid"].tolist() == [2, 3] and b["indicator"].tolist() == [False, False, False, True]


def test_large_path_hub_rows_equal_bitset_path(cuda_device):
    """A power-law target with 1000-neighbour hubs: rows above 256 entries are walked by all warps of a shared-memory-tier
    CTA together, and the depth-2 balls around hubs overflow into the team tier.  Both must equal the bitset kernel."""
    from desco_b200.graph import gen_powerlaw

    csr = gen_powerlaw(20000, 200000, seed=9, max_deg_frac=0.05)
    deg = np.diff(csr.rowptr)
    assert deg.max() > 512
    rng = np.random.default_rng(2)
    hubs = np.argsort(deg)[-24:]
    centres = np.unique(np.concatenate([rng.choice(csr.num_nodes, size=1500, replace=False), hubs,
                                        csr.col[csr.rowptr[hubs[-1]]:csr.rowptr[hubs[-1]] + 64]])).astype(np.int32)
    a = _gpu_partition(csr, 2, "hetero", centres, large=False)
    b = _gpu_partition(csr, 2, "hetero", centres, large=True)
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    a = _gpu_partition(csr, 3, "canonical", centres[::7], large=False)
    b = _gpu_partition(csr, 3, "canonical", centres[::7], large=True)
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
