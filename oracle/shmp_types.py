"""Oracle: SHMP edge typing (triangle / tride).  TEST INFRASTRUCTURE ONLY.

Literal restatement of ``ToTconvHetero.__call__`` (``subgraph_counting/transforms.py:180-255``) on ONE neighborhood
given as a directed edge list over its own homogeneous node ids: A = ones on edges (:201-207), A2 = A@A (:208),
T = (A*A2 + A).coalesce() (:209-211), tri = T.values() > 1 (:221), un-sorted back to the input edge order (:190-195,
:223-225).  ``coalesce`` returns row-major sorted indices, which is what ``sort_edge_index`` produces (App. B.6).
"""
from __future__ import annotations

import warnings
from typing import Dict

import numpy as np
import torch


def tri_flags_sparse(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """edge_index [2,e] int64 (simple, symmetric) -> bool[e] in the INPUT edge order."""
    e = edge_index.shape[1]
    if e == 0:
        return torch.zeros(0, dtype=torch.bool)
    # sort_edge_index (row-major) with a carried permutation  (transforms.py:190-195)
    key = edge_index[0] * num_nodes + edge_index[1]
    sort_idx = torch.argsort(key, stable=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        A = torch.sparse_coo_tensor(edge_index, torch.ones(e), (num_nodes, num_nodes))
        A2 = torch.sparse.mm(A, A)
        T = (A * A2 + A).coalesce()
    assert T.indices().shape[1] == e, "A*A2+A must have exactly the pattern of A"
    tri_sorted = T.values() > 1
    # gather(index=sort_idx.argsort())  (transforms.py:223-225)
    return tri_sorted[torch.argsort(sort_idx)]


def type_batch(batch: Dict[str, np.ndarray]) -> np.ndarray:
    """Apply the literal transform to every neighborhood of a packed batch -> uint8[E] (same order as edge_col)."""
    out = np.zeros(len(batch["edge_col"]), dtype=np.uint8)
    nbh_ptr, edge_ptr, edge_col = batch["nbh_ptr"], batch["edge_ptr"], batch["edge_col"]
    for g in range(len(nbh_ptr) - 1):
        lo, hi = int(nbh_ptr[g]), int(nbh_ptr[g + 1])
        e0, e1 = int(edge_ptr[lo]), int(edge_ptr[hi])
        deg = np.diff(edge_ptr[lo : hi + 1])
        src = np.repeat(np.arange(hi - lo), deg)
        dst = edge_col[e0:e1].astype(np.int64) - lo
        ei = torch.from_numpy(np.stack([src, dst]).astype(np.int64))
        out[e0:e1] = tri_flags_sparse(ei, hi - lo).numpy().astype(np.uint8)
    return out
