"""CPU: the ground-truth oracle (the reference's own VF2 worker executed from source when the tree is mounted, its literal
restatement otherwise) against the committed fixture, the Python restatement of the kernel's ESU enumeration against VF2,
and the host-built pattern tables."""
import os

import networkx as nx
import numpy as np

from desco_b200.graph import TargetCSR
from desco_b200.groundtruth import SymmetricFactor, pattern_tables
from oracle import groundtruth as GT
from oracle import model as M


def _sets(golden_dir):
    z = np.load(os.path.join(golden_dir, "groundtruth_ref.npz"))
    names = sorted({k.rsplit("_", 1)[0] for k in z.files if k.endswith("_truth")})
    return z, {n: (TargetCSR(z[f"{n}_rowptr"], z[f"{n}_col"], z[f"{n}_graph_ptr"]), z[f"{n}_truth"]) for n in names}


def test_oracle_reproduces_the_reference_fixture(golden_dir):
    z, sets = _sets(golden_dir)
    queries = [nx.graph_atlas(int(i)) for i in z["query_ids"]]
    assert list(z["query_ids"]) == M.STANDARD_QUERY_IDS
    funcs = GT.reference_functions()  # the reference's own functions in the builder container
    for name in ("kat", "mutag"):
        csr, truth = sets[name]
        assert np.array_equal(GT.canonical_count_truth(csr, queries, funcs), truth), name
        assert np.array_equal(GT.canonical_count_truth(csr, queries), truth), name  # the restatement


def test_esu_enumeration_equals_vf2(golden_dir):
    """The algorithm of csrc/groundtruth.cu (ESU rooted at the largest node, exclusive-neighbourhood extension) counts
    every occurrence exactly once - checked against the fixture on sparse, dense, star and clique graphs."""
    z, sets = _sets(golden_dir)
    queries = [nx.graph_atlas(int(i)) for i in z["query_ids"]]
    for name in ("kat", "imdb"):
        csr, truth = sets[name]
        got = np.concatenate([GT.esu_counts(csr.to_networkx(g), queries) for g in range(csr.num_graphs)])
        assert np.array_equal(got, truth), name
    # every node set is credited to exactly one node: the sum over nodes is the graphlet count of the graph
    csr, truth = sets["kat"]
    g = csr.to_networkx(csr.num_graphs - 3)  # K6
    lo, hi = csr.graph_ptr[csr.num_graphs - 3], csr.graph_ptr[csr.num_graphs - 2]
    tot = truth[lo:hi].sum(0)
    k5 = M.STANDARD_QUERY_IDS.index(52)  # atlas 52 = K5
    assert g.number_of_edges() == 15 and tot[k5] == 6 and tot.sum() == 20 + 15 + 6  # C(6,3) triangles, C(6,4) K4s, C(6,5) K5s


def test_pattern_tables_and_symmetry_factors():
    queries = [nx.graph_atlas(i) for i in M.STANDARD_QUERY_IDS]
    l3, l4, l5 = pattern_tables(queries)
    assert [(t != 255).sum() for t in (l3, l4, l5)] == [4, 38, 728]  # connected labelled graphs on 3 / 4 / 5 nodes
    assert l3[0b111] == M.STANDARD_QUERY_IDS.index(7) and l3[0b011] == M.STANDARD_QUERY_IDS.index(6)
    assert l5[(1 << 10) - 1] == M.STANDARD_QUERY_IDS.index(52)
    for q in queries:
        assert SymmetricFactor(q) == GT.SymmetricFactor(q)
    # labelled occurrences x automorphisms = k! for every class
    import math

    for k, lut in ((3, l3), (4, l4), (5, l5)):
        for qi, q in enumerate(queries):
            if q.number_of_nodes() == k:
                assert (lut == qi).sum() * SymmetricFactor(q) == math.factorial(k)
