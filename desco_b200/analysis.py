"""Accuracy metrics of the counting pipeline (``subgraph_counting/analysis.py:22-83``, used by ``main.py:425-447`` on the
graph-level counts): per query group, over arrays ``pred`` / ``truth`` of shape [num_graphs, num_queries]."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def _groups(num_queries: int, groupby: Optional[Sequence[Sequence[int]]]) -> List[List[int]]:
    return [list(range(num_queries))] if groupby is None else [list(g) for g in groupby]


def mse(pred: np.ndarray, truth: np.ndarray, groupby: Optional[Sequence[Sequence[int]]] = None, verbose: bool = False) -> List[float]:
    """Mean squared error per query group (``analysis.py:46-64``)."""
    p, t = np.asarray(pred, dtype=np.float64), np.asarray(truth, dtype=np.float64)
    out = [float(np.mean((p[:, g] - t[:, g]) ** 2)) for g in _groups(p.shape[1], groupby)]
    if verbose:
        for v in out:
            print("mean_mse: ", v)
    return out


def norm_mse(pred: np.ndarray, truth: np.ndarray, groupby: Optional[Sequence[Sequence[int]]] = None,
             verbose: bool = False) -> List[float]:
    """MSE divided by the variance of the truth of the group (``analysis.py:22-43``), the paper's headline metric."""
    t = np.asarray(truth, dtype=np.float64)
    out = [m / float(np.var(t[:, g])) for m, g in zip(mse(pred, truth, groupby), _groups(t.shape[1], groupby))]
    if verbose:
        for v in out:
            print("mean_norm_mse: ", v)
    return out


def mae(pred: np.ndarray, truth: np.ndarray, groupby: Optional[Sequence[Sequence[int]]] = None, verbose: bool = False) -> List[float]:
    """Mean absolute error per query group (``analysis.py:67-83``; the reference requires ``groupby``)."""
    p, t = np.asarray(pred), np.asarray(truth)
    out = [float(np.mean(np.abs(p[:, g] - t[:, g]))) for g in _groups(p.shape[1], groupby)]
    if verbose:
        for v in out:
            print("mean_mae: ", v)
    return out


def round_counts(pred) -> np.ndarray:
    """``main.py:425``: graph-level predictions are clamped at zero and rounded before they are scored."""
    import torch

    if isinstance(pred, torch.Tensor):
        pred = pred.detach().cpu().numpy()
    return np.round(np.maximum(np.asarray(pred, dtype=np.float64), 0.0))
