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


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
def test_metrics_equal_the_reference_functions_executed_from_source():
    """norm_mse / mse / mae of subgraph_counting/analysis.py:22-83 (pure numpy; the module itself imports the whole package)
    compiled from the reference source and run on the same arrays."""
    import ast

    path = "/root/reference/subgraph_counting/analysis.py"
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("norm_mse", "mse", "mae")]
    for fn in body:
        fn.returns = None
        for a in fn.args.args:
            a.annotation = None
    ns = {"np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    rng = np.random.default_rng(1)
    truth = np.floor(np.exp(rng.normal(0, 1.5, size=(80, 29)))).astype(np.float32)
    pred = (truth + rng.normal(0, 1.0, size=truth.shape)).astype(np.float32)
    groups = [[0, 1], list(range(2, 8)), list(range(8, 29))]  # the paper's 3- / 4- / 5-node query groups
    for ours, name in ((norm_mse, "norm_mse"), (mse, "mse"), (mae, "mae")):
        assert ours(pred, truth, groups) == pytest.approx(ns[name](pred, truth, groups), rel=1e-12)
    assert norm_mse(pred, truth) == pytest.approx(ns["norm_mse"](pred, truth), rel=1e-12)
    assert mse(pred, truth) == pytest.approx(ns["mse"](pred, truth), rel=1e-12)
