"""GPU parity: ground-truth canonical counts (csrc/groundtruth.cu through the C ABI) - exact integers - against the
fixture produced by the reference's own MatchSubgraphWorker + SymmetricFactor and against the oracle on seeded sets."""
import os

import networkx as nx
import numpy as np
import pytest
import torch

from desco_b200.graph import TargetCSR, csr_from_networkx, gen_enzymes_shaped, gen_mutag_shaped, gen_syn1827_shaped

pytestmark = pytest.mark.gpu


def test_ground_truth_matches_reference_fixture(cuda_device, golden_dir):
    from desco_b200.data import DeviceCSR
    from desco_b200.groundtruth import canonical_count_truth

    z = np.load(os.path.join(golden_dir, "groundtruth_ref.npz"))
    ids = [int(i) for i in z["query_ids"]]
    for name in sorted({k.rsplit("_", 1)[0] for k in z.files if k.endswith("_truth")}):
        csr = TargetCSR(z[f"{name}_rowptr"], z[f"{name}_col"], z[f"{name}_graph_ptr"])
        got = canonical_count_truth(DeviceCSR.from_host(csr), query_ids=ids).cpu().numpy()
        assert np.array_equal(got, z[f"{name}_truth"].astype(np.float32)), name


def test_ground_truth_matches_oracle_and_graphlet_totals(cuda_device, tmp_path):
    from desco_b200.workload import Workload
    from oracle import groundtruth as GT
    from oracle import model as M

    ids = [6, 7, 13, 14, 15, 16, 17, 18]  # 3- and 4-node queries: VF2 stays in seconds on these sets
    queries = [nx.graph_atlas(i) for i in ids]
    for csr in (gen_mutag_shaped(seed=3, num_graphs=30), gen_enzymes_shaped(seed=3, num_graphs=10)):
        wl = Workload(csr, str(tmp_path))
        truth = wl.compute_groundtruth(query_ids=ids)
        assert truth.shape == (csr.num_nodes, len(ids)) and wl.exist_groundtruth(ids)
        assert np.array_equal(truth.numpy(), GT.canonical_count_truth(csr, queries).astype(np.float32))
        assert torch.equal(Workload(csr, str(tmp_path)).load_groundtruth(ids), truth)  # the reference's cache-file scheme
    # size-independent property on graphs VF2 cannot finish: every occurrence is credited to exactly one node, so the
    # per-graph sum of the triangle column is the triangle count of the graph (networkx, one matrix product)
    csr = gen_syn1827_shaped(seed=1, stride=150)
    wl = Workload(csr, None)
    truth = wl.compute_groundtruth(query_ids=M.STANDARD_QUERY_IDS, save_to_file=False)
    tri_col = M.STANDARD_QUERY_IDS.index(7)
    for g in range(0, csr.num_graphs, 3):
        lo, hi = int(csr.graph_ptr[g]), int(csr.graph_ptr[g + 1])
        assert int(truth[lo:hi, tri_col].sum()) == sum(nx.triangles(csr.to_networkx(g)).values()) // 3
    assert (truth >= 0).all() and truth[:, tri_col].sum() > 0


def test_ground_truth_rejects_what_it_does_not_count(cuda_device):
    from desco_b200.data import DeviceCSR
    from desco_b200.groundtruth import canonical_count_truth

    d = DeviceCSR.from_host(csr_from_networkx([nx.path_graph(5)]))
    with pytest.raises(NotImplementedError):
        canonical_count_truth(d, queries=[nx.path_graph(6)])
    with pytest.raises(NotImplementedError):
        canonical_count_truth(d, queries=[nx.empty_graph(3)])
    with pytest.raises(ValueError):
        canonical_count_truth(d)
    got = canonical_count_truth(d, queries=[nx.path_graph(3), nx.path_graph(5)]).cpu().numpy()
    assert got[:, 0].tolist() == [0, 0, 1, 1, 1] and got[:, 1].tolist() == [0, 0, 0, 0, 1]
