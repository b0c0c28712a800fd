"""Host-side ingest and reporting: TU-format files -> CSR, norm_mse / mse / mae.  CPU only."""
import os

import numpy as np
import pytest

from desco_b200.analysis import mae, mse, norm_mse, round_counts
from desco_b200.graph import csr_from_graph_list, load_tu_dataset


def _write_tu(root, name, graphs):
    d = os.path.join(root, name, "raw")
    os.makedirs(d)
    base, lines, ind = 0, [], []
    for gid, (n, edges) in enumerate(graphs, start=1):
        for u, v in edges:  # TU files list both directions
            lines += [f"{base + u + 1}, {base + v + 1}", f"{base + v + 1}, {base + u + 1}"]
        ind += [str(gid)] * n
        base += n
    open(os.path.join(d, f"{name}_A.txt"), "w").write("\n".join(lines) + "\n")
    open(os.path.join(d, f"{name}_graph_indicator.txt"), "w").write("\n".join(ind) + "\n")


def test_tu_dataset_round_trip(tmp_path):
    graphs = [(4, [(0, 1), (1, 2), (2, 0), (2, 3)]), (3, [(0, 2)]), (5, [(4, 0), (0, 1), (1, 4), (2, 3)])]
    _write_tu(str(tmp_path), "TOY", graphs)
    got = load_tu_dataset(str(tmp_path), "TOY")
    ref = csr_from_graph_list([(n, np.array(e)) for n, e in graphs])
    assert np.array_equal(got.rowptr, ref.rowptr) and np.array_equal(got.col, ref.col)
    assert got.graph_ptr.tolist() == [0, 4, 7, 12]
    g = got.to_networkx(2)
    assert sorted(g.edges()) == [(0, 1), (0, 4), (1, 4), (2, 3)]


def test_tu_dataset_rejects_cross_graph_edges(tmp_path):
    d = os.path.join(str(tmp_path), "BAD", "raw")
    os.makedirs(d)
    open(os.path.join(d, "BAD_A.txt"), "w").write("1, 3\n3, 1\n")
    open(os.path.join(d, "BAD_graph_indicator.txt"), "w").write("1\n1\n2\n")
    with pytest.raises(ValueError):
        load_tu_dataset(str(tmp_path), "BAD")
    with pytest.raises(FileNotFoundError):
        load_tu_dataset(str(tmp_path), "MISSING")


def test_metrics_match_their_definitions():
    rng = np.random.default_rng(0)
    truth = np.floor(np.exp(rng.normal(0, 1.5, size=(50, 6))))
    pred = truth + rng.normal(0, 0.7, size=truth.shape)
    groups = [[0, 1], [2, 3, 4, 5]]
    m = mse(pred, truth, groups)
    assert m == pytest.approx([np.mean((pred[:, g] - truth[:, g]) ** 2) for g in groups])
    nm = norm_mse(pred, truth, groups)
    assert nm == pytest.approx([m[i] / np.var(truth[:, g]) for i, g in enumerate(groups)])
    assert norm_mse(pred, truth) == pytest.approx([np.mean((pred - truth) ** 2) / np.var(truth)])
    assert mae(pred, truth, groups) == pytest.approx([np.mean(np.abs(pred[:, g] - truth[:, g])) for g in groups])
    assert norm_mse(truth, truth, groups) == [0.0, 0.0]
    assert round_counts(np.array([-0.4, 0.49, 2.5, 3.51])).tolist() == [0.0, 0.0, 2.0, 4.0]
