"""Counting driver of the DeSCo hot path.

Mirrors ``subgraph_counting/workload.py`` (reference @ 4508f7a): ``NeighborhoodDataset`` :153 (``process`` :215,
``aggregate_neighborhood_count`` :303), ``GossipDataset`` :48 (``apply_neighborhood_count`` :107,
``aggregate_neighborhood_count`` :136) and ``Workload`` :363 (``generate_pipeline_datasets`` :422,
``apply_neighborhood_count`` :728) - with the Python double loop over graphs x nodes replaced by the batched CUDA
partition (three launches for the whole dataset) and every tensor resident in HBM.

Targets may be given as a ``TargetCSR`` / ``DeviceCSR``, a list of networkx graphs (nodes 0..n-1) or a list of
PyG-style ``Data`` objects (duck-typed: ``edge_index``, ``num_nodes``); torch_geometric itself is optional.
Ground-truth generation (VF2, ``workload.py:551-726``) is out of scope.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Iterator, Optional

import numpy as np
import torch

from .data import DeviceCSR, NeighborhoodBatch, partition_batch
from .graph import TargetCSR, csr_from_graph_list, csr_from_networkx


def as_device_csr(dataset, device=None) -> DeviceCSR:
    """Ingest: anything the reference accepts as a list of target graphs -> one block-diagonal CSR in HBM
    (replaces ``pyg.utils.to_networkx(g, to_undirected=True)`` per graph, ``workload.py:223-231``)."""
    if isinstance(dataset, DeviceCSR):
        return dataset
    if isinstance(dataset, TargetCSR):
        return DeviceCSR.from_host(dataset, device)
    graphs = list(dataset)
    if not graphs:
        raise ValueError("empty dataset")
    g0 = graphs[0]
    if hasattr(g0, "edge_index") and hasattr(g0, "num_nodes"):  # PyG Data duck type
        gl = []
        for g in graphs:
            ei = g.edge_index.detach().cpu().numpy().T.reshape(-1, 2)
            ei = ei[ei[:, 0] <= ei[:, 1]]  # to_networkx(to_undirected=True) keeps u <= v only (workload.py:224; data._to_nx_graph)
            gl.append((int(g.num_nodes), ei))
        return DeviceCSR.from_host(csr_from_graph_list(gl), device)
    return DeviceCSR.from_host(csr_from_networkx(graphs), device)


class NeighborhoodDataset:
    """``workload.py:153-324``.  ``process`` runs the canonical partition + SHMP typing of EVERY node of every target
    on the GPU and keeps the result as one packed ``NeighborhoodBatch``; ``nx_neighs_index`` / ``nx_neighs_indicator``
    are the reference's bookkeeping arrays (:259, :293-294)."""

    def __init__(self, depth_neigh, root, dataset=None, nx_targets=None, transform=None, pre_transform=None,
                 pre_filter=None, hetero_graph=True, node_feat=False, node_feat_key="feat", device=None):
        self.depth_neigh, self.root, self.hetero_graph = depth_neigh, root, hetero_graph
        self.node_feat, self.node_feat_key = node_feat, node_feat_key
        self.transform, self.pre_transform, self.pre_filter = transform, pre_transform, pre_filter
        self.dataset = dataset if dataset is not None else nx_targets
        self.graph = as_device_csr(self.dataset, device)
        self.process()

    @property
    def processed_file_names(self):
        suffix = "_node_feat" if self.node_feat else ""
        suffix_homo = "_homo" if not self.hetero_graph else ""
        d = str(self.depth_neigh)
        return ["neighs_packed_depth_" + d + suffix + suffix_homo + ".npz",
                "neighs_index_depth_" + d + suffix + suffix_homo + ".npy",
                "neighs_indicator_depth_" + d + suffix + suffix_homo + ".npy"]

    def process(self):
        mode = "hetero" if self.hetero_graph else "canonical"  # workload.py:238-241
        self.batch = partition_batch(self.graph, None, self.depth_neigh, mode)
        self.nx_neighs_index = self.batch.index()  # (#neighborhood, 2) = (graph id, local node id)
        self.nx_neighs_indicator = self.batch.indicator.cpu().numpy().astype(bool)  # (#node,)
        if self.root is not None:  # the index / indicator files keep the reference's names and formats (:198-213)
            os.makedirs(os.path.join(self.root, "processed"), exist_ok=True)
            paths = [os.path.join(self.root, "processed", f) for f in self.processed_file_names]
            np.savez_compressed(paths[0], **{k: v for k, v in self.batch.to_numpy().items()})
            np.save(paths[1], self.nx_neighs_index)
            np.save(paths[2], self.nx_neighs_indicator)
        return self.batch

    def __len__(self) -> int:
        return self.batch.num_neighborhoods

    def loader(self, batch_size: int = 512, shuffle: bool = False, generator=None) -> Iterator[NeighborhoodBatch]:
        """Neighborhoods in chunks of ``batch_size`` (``config.py:255``) - the PyG ``DataLoader`` analogue: consecutive, or
        (``shuffle=True``) a fresh random permutation per pass, each chunk gathered into its own packed batch on the device.
        The truth counts of the chunk ride along as ``batch.y`` once ``apply_truth_from_dataset`` has set them."""
        G = len(self)
        y = getattr(self, "y", None)
        if shuffle:
            perm = torch.randperm(G, generator=generator, device="cpu").to(self.batch.nbh_ptr.device)
            step = G if batch_size <= 0 else batch_size
            for g0 in range(0, G, step):
                idx = perm[g0:g0 + step]
                b = self.batch.select(idx)
                b.index_in_dataset = idx
                if y is not None:
                    b.y = y.to(idx.device)[idx]
                yield b
            return
        if batch_size <= 0 or batch_size >= G:
            if y is not None:
                self.batch.y = y
            yield self.batch
            return
        for g0 in range(0, G, batch_size):
            b = self.batch.slice(g0, g0 + batch_size)
            if y is not None:
                b.y = y[g0:g0 + batch_size]
            yield b

    def apply_truth_from_dataset(self, truth: torch.Tensor):
        """``workload.py:296-301``: truth [#node, #query] -> y of the kept neighborhoods."""
        self.y = truth[torch.as_tensor(self.nx_neighs_indicator)]

    def aggregate_neighborhood_count(self, count: torch.Tensor) -> torch.Tensor:
        """``workload.py:303-324``: [#neighborhood, #query] -> [#graph, #query] (CPU float, like the reference)."""
        gid = torch.as_tensor(self.nx_neighs_index[:, 0], dtype=torch.long, device=count.device)
        out = torch.zeros((self.graph.num_graphs, count.shape[1]), dtype=torch.float32, device=count.device)
        out.index_add_(0, gid, count.detach().to(torch.float32))
        return out.cpu()


class GossipDataset:
    """``workload.py:48-150``: the original target graphs with the neighborhood counts as node features ``x``."""

    def __init__(self, dataset, root, transform=None, pre_transform=None, pre_filter=None, hetero_graph=True, device=None):
        self.dataset, self.root, self.hetero_graph = dataset, root, hetero_graph
        self.graph = as_device_csr(dataset, device)
        self.slices = {"y": self.graph.graph_ptr.to(torch.long)}  # per-graph node ranges (segment_csr pointer, :146-148)
        self.slices["x"] = self.slices["y"]
        self.x: Optional[torch.Tensor] = None
        self.y: Optional[torch.Tensor] = None
        self.data = self  # reference code reads dataset.data.x / dataset.data.y

    def __len__(self) -> int:
        return self.graph.num_graphs

    def apply_truth_from_dataset(self, truth: torch.Tensor):
        self.y = truth

    def apply_neighborhood_count(self, count: torch.Tensor, neighborhood_indicator):
        """``workload.py:107-126``: x = zeros[#node, #query]; x[indicator] = count."""
        num_query = count.shape[1]
        dev = self.graph.rowptr.device
        ind = torch.as_tensor(neighborhood_indicator, device=dev).to(torch.bool)
        self.x = torch.zeros((ind.numel(), num_query), dtype=torch.float32, device=dev)
        self.x[ind] = count.detach().to(device=dev, dtype=torch.float32)

    def apply_neighborhood_embeddings(self, embedding: torch.Tensor, neighborhood_indicator):
        """``workload.py:128-134``."""
        dev = self.graph.rowptr.device
        ind = torch.as_tensor(neighborhood_indicator, device=dev).to(torch.bool)
        e = torch.zeros((ind.numel(), embedding.shape[1]), dtype=torch.float32, device=dev)
        e[ind] = embedding.detach().to(dev)
        self.x = e if self.x is None else torch.cat([self.x, e], dim=1)

    def loader(self, batch_size: int = 256):
        """One batch for the whole dataset: the graphs are independent blocks of one CSR, so the reference's
        256-graph batches (``config.py:319``) only bounded memory; results are identical."""
        yield SimpleNamespace(graph=self.graph, x=self.x, y=self.y)

    def aggregate_neighborhood_count(self, count: torch.Tensor) -> torch.Tensor:
        """``workload.py:136-148``: segment sum of [#node, #query] over the graphs -> [#graph, #query]."""
        ptr = self.slices["y"].to(count.device)
        gid = torch.repeat_interleave(torch.arange(ptr.numel() - 1, device=count.device), ptr[1:] - ptr[:-1])
        out = torch.zeros((ptr.numel() - 1, count.shape[1]), dtype=count.dtype, device=count.device)
        return out.index_add_(0, gid, count)


class Workload:
    """``workload.py:363-471, 728-736``."""

    def __init__(self, dataset, root: Optional[str], hetero_graph: bool = True, node_feat_len: int = -1,
                 node_feat_key: str = "feat", device=None, **kwargs):
        self.dataset, self.root, self.hetero_graph = dataset, root, hetero_graph
        self.use_node_feat = node_feat_len != -1
        if self.use_node_feat:
            raise NotImplementedError("--use_node_feature is not on the default path (SURVEY.md section 2: utils.py)")
        self.node_feat_len, self.node_feat_key = 1, "feat"
        self.queries, self.query_ids = [], []
        self.canonical_count_truth = torch.tensor([[]])
        self.neighborhood_dataset: Optional[NeighborhoodDataset] = None
        self.gossip_dataset: Optional[GossipDataset] = None
        self.device = device
        self._graph = None

    @property
    def graph(self) -> DeviceCSR:
        if self._graph is None:
            self._graph = as_device_csr(self.dataset, self.device)
        return self._graph

    def generate_pipeline_datasets(self, depth_neigh, neighborhood_transform=None, gossip_transform=None,
                                   pre_transform=None, pre_filter=None):
        """``workload.py:422-471``.  Zero node features (:431-440) are implicit (the kernels read feat == NULL as
        zeros); ``neighborhood_transform=ToTconvHetero()`` is what the partition kernel already emits (edge_tri)."""
        sub = lambda name: None if self.root is None else os.path.join(self.root, name)
        self.neighborhood_dataset = NeighborhoodDataset(
            depth_neigh, sub("NeighborhoodDataset"), dataset=self.graph, transform=neighborhood_transform,
            pre_transform=pre_transform, pre_filter=pre_filter, hetero_graph=self.hetero_graph)
        self.gossip_dataset = GossipDataset(self.graph, sub("GossipDataset"), transform=gossip_transform,
                                            pre_transform=pre_transform, pre_filter=pre_filter,
                                            hetero_graph=self.hetero_graph)
        if self.canonical_count_truth.shape[1] != 0:
            self.neighborhood_dataset.apply_truth_from_dataset(self.canonical_count_truth)
            self.gossip_dataset.apply_truth_from_dataset(self.canonical_count_truth)

    # ---- ground truth (labels): workload.py:473-726 ----
    max_file_name_len = 30  # workload.py:412

    def _truth_file(self, query_ids, queries) -> str:
        """File name scheme of the reference (``workload.py:484-507``) so existing label caches are found."""
        if (query_ids is None) == (queries is None):
            raise ValueError("query_ids or queries must be given (and not both)")
        if query_ids is not None:
            name = "query_num_{:d}_".format(len(query_ids)) + "atlas_ids_" + "_".join(map(str, query_ids[: self.max_file_name_len])) + ".pt"
        else:
            name = "query_num_{:d}_".format(len(queries)) + "query_len_sum_" + str(sum(len(g) for g in queries)) + ".pt"
        return os.path.join(self.root, "CanonicalCountTruth", name)

    def exist_groundtruth(self, query_ids, queries=None) -> bool:
        """``workload.py:512-549``."""
        return self.root is not None and os.path.exists(self._truth_file(query_ids, queries))

    def load_groundtruth(self, query_ids, queries=None) -> torch.Tensor:
        """``workload.py:473-510``: the cached ``count_motif`` tensor (plain ``torch.save`` of a tensor)."""
        if not self.exist_groundtruth(query_ids, queries):
            raise NotImplementedError
        self.canonical_count_truth = torch.load(self._truth_file(query_ids, queries), weights_only=True)
        return self.canonical_count_truth

    def compute_groundtruth(self, query_ids=None, queries=None, num_workers=-1, save_to_file=True) -> torch.Tensor:
        """``workload.py:551-726``: canonical counts of every query at every node, [#node, #query] - on the GPU
        (csrc/groundtruth.cu) instead of one networkx VF2 process per (target, query); ``num_workers`` is ignored."""
        from .groundtruth import canonical_count_truth

        truth = canonical_count_truth(self.graph, query_ids=query_ids, queries=queries).cpu()
        self.query_ids = list(query_ids) if query_ids is not None else [-i for i in range(len(queries))]
        self.queries = queries
        self.canonical_count_truth = truth
        if save_to_file and self.root is not None:
            path = self._truth_file(query_ids, queries)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            torch.save(truth, path)
        return truth

    def apply_neighborhood_count(self, count):
        self.gossip_dataset.apply_neighborhood_count(count, self.neighborhood_dataset.nx_neighs_indicator)

    def apply_neighborhood_embeddings(self, embeddings):
        self.gossip_dataset.apply_neighborhood_embeddings(embeddings, self.neighborhood_dataset.nx_neighs_indicator)


def count_subgraphs(workload: Workload, neighborhood_model, gossip_model=None, depth: int = 4, batch_size: int = 512):
    """The inference sequence of the reference's ``main.py:285-302, 334, 418-447`` in one call: canonical partition ->
    neighborhood counts -> (node features of the gossip dataset) -> gossip propagation -> graph-level sums.
    Returns ``(count_neighborhood [G,Q], count_node [N,Q] or None, count_graph [#graph,Q])``."""
    if workload.neighborhood_dataset is None or workload.neighborhood_dataset.depth_neigh != depth:
        workload.generate_pipeline_datasets(depth_neigh=depth)
    nd = workload.neighborhood_dataset
    neighborhood_model.set_pyg_batch_size(batch_size)
    with torch.no_grad():
        counts = neighborhood_model.graph_to_count(nd.batch)  # whole dataset in one pass; batch_size only shapes the quirk
        workload.apply_neighborhood_count(counts)
        if gossip_model is None:
            return counts, None, nd.aggregate_neighborhood_count(counts)
        gossip_model.set_query_emb(neighborhood_model.get_query_emb())
        node_counts = torch.cat([gossip_model.graph_to_count(b) for b in workload.gossip_dataset.loader()], dim=0)
        return counts, node_counts, workload.gossip_dataset.aggregate_neighborhood_count(node_counts)
