"""Oracle: ground-truth canonical counts (CPU, networkx VF2).  TEST INFRASTRUCTURE ONLY.

``reference_functions()`` executes the reference's OWN ``MatchSubgraphWorker`` (``subgraph_counting/workload.py:327-348``)
and ``SymmetricFactor`` / ``GenVMap`` (``subgraph_counting/data.py:61-88``) from source - they are pure networkx - and
``canonical_count_truth`` drives them like ``Workload.compute_groundtruth`` (``workload.py:600-699``): one VF2 run per
(target, query), every mapping credited to the largest matched node, counts divided by the query's symmetry factor.
When the reference tree is not mounted (the GPU box) the literal restatement below is used instead.
``esu_counts`` restates, in plain Python sets, the enumeration the CUDA kernel performs (csrc/groundtruth.cu) so that
the algorithm itself can be checked against VF2 without a GPU.
"""
from __future__ import annotations

import ast
import os
from collections import defaultdict
from typing import Callable, Dict, List, Optional

import networkx as nx
import numpy as np

REFERENCE_ROOT = "/root/reference/subgraph_counting"


def MatchSubgraphWorker(task):
    """``workload.py:327-348``."""
    tid, target, qid, query, node_feat_key = task
    gm = nx.algorithms.isomorphism.GraphMatcher(target, query)
    count = defaultdict(int)
    for vmap in gm.subgraph_isomorphisms_iter():
        count[max(vmap.keys())] += 1
    return tid, qid, tuple(count.items())


def SymmetricFactor(graph, node_feat_key=None) -> int:
    """``data.py:61-88``: number of mappings of the graph onto itself."""
    return sum(1 for _ in nx.algorithms.isomorphism.GraphMatcher(graph, graph).subgraph_isomorphisms_iter())


def reference_functions() -> Optional[Dict[str, Callable]]:
    """The reference's own three functions compiled from its source files (None when the tree is absent)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return None
    ns = {"nx": nx, "defaultdict": defaultdict}

    def take(path, names):
        tree = ast.parse(open(path).read())
        body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
        for fn in body:
            fn.returns = None
            for a in fn.args.args:
                a.annotation = None
        exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)

    take(os.path.join(REFERENCE_ROOT, "workload.py"), {"MatchSubgraphWorker"})
    take(os.path.join(REFERENCE_ROOT, "data.py"), {"SymmetricFactor", "GenVMap"})
    return {k: ns[k] for k in ("MatchSubgraphWorker", "SymmetricFactor", "GenVMap")}


def canonical_count_truth(csr, queries: List[nx.Graph], funcs: Optional[Dict[str, Callable]] = None) -> np.ndarray:
    """[num_nodes, num_queries] float64, dataset-global node order (``workload.py:600-699`` without the process pool)."""
    worker = (funcs or {}).get("MatchSubgraphWorker", MatchSubgraphWorker)
    sym = (funcs or {}).get("SymmetricFactor", SymmetricFactor)
    factors = [sym(q, None) for q in queries]
    out = np.zeros((csr.num_nodes, len(queries)), dtype=np.float64)
    for gid in range(csr.num_graphs):
        target = csr.to_networkx(gid)
        base = int(csr.graph_ptr[gid])
        for qi, q in enumerate(queries):
            _, _, counts = worker((gid, target, qi, q, None))
            for node, c in counts:
                out[base + node, qi] = c / factors[qi]
    return out


def esu_counts(g: nx.Graph, queries: List[nx.Graph]) -> np.ndarray:
    """The kernel's algorithm in Python: for every root v, ESU over nodes < v with the exclusive-neighbourhood rule
    enumerates each connected induced subgraph with maximum v once; classify by isomorphism with the queries."""
    n = g.number_of_nodes()
    out = np.zeros((n, len(queries)), dtype=np.int64)
    kmax = max(q.number_of_nodes() for q in queries)
    by_size = defaultdict(list)
    for qi, q in enumerate(queries):
        by_size[q.number_of_nodes()].append((qi, q))
    adj = {u: set(g.neighbors(u)) - {u} for u in g.nodes}

    def credit(v, sub):
        h = g.subgraph(sub)
        for qi, q in by_size.get(len(sub), ()):
            if nx.is_isomorphic(h, q):
                out[v, qi] += 1

    def extend(v, sub, ext, closed):
        if len(sub) >= 3:
            credit(v, sub)
        if len(sub) == kmax:
            return
        ext = sorted(ext)
        for i, w in enumerate(ext):
            rem = set(ext[i + 1:])
            new = {u for u in adj[w] if u < v and u not in closed}
            extend(v, sub + [w], rem | new, closed | adj[w] | {w})

    for v in g.nodes:
        extend(v, [v], {u for u in adj[v] if u < v}, adj[v] | {v})
    return out
