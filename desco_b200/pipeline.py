"""The partition -> SHMP -> count-head step without host round trips, captured as ONE CUDA graph.

The reference computes the canonical partition on the CPU once per dataset (``workload.py:215-294``) and then runs a
Python loop of small launches per batch (``lightning_model.py:198-222``).  ``partition_batch`` + ``graph_to_count`` here
already fuse that into ~12 launches, but between the partition's scan and its fill the HOST reads the output sizes (one
stream synchronisation per step, then ~0.1 ms of launch latency while the GPU idles).  ``NeighborhoodCountStep`` removes
it: buffers are sized by a capacity learnt from a first eager pass, the batch's own sizes stay on the device
(``desco_partition_batch_async`` -> ``desco_shmp_forward_dev`` -> ``desco_count_head_dev``) and the whole step is replayed
as a CUDA graph.  Serves batches whose neighborhoods fit the fused tensor-core kernel (<= 128 rows: configs 1, 2, 4);
anything else raises at construction and the caller keeps using the eager path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .data import MODE_HETERO, DeviceCSR, LARGE_GRAPH_NODES, _ptr, partition_batch
from .gnn_model import PRECISION, TILE_ROWS


def pack_int32_block(arrays: Sequence, align: int = 128) -> Tuple[torch.Tensor, List[Tuple[int, int]]]:
    """Lay int32 arrays (the CSR parts and the centre list of a step) out back to back in ONE host buffer - pinned when
    CUDA is there - every part starting on a multiple of ``align`` elements.  The same layout on the device (``.to(device)``,
    then ``block[o:o + n]`` views as the step's CSR and centre storage) lets a single host->device copy refresh everything
    the step reads, instead of one copy per array.  Returns ``(block, [(offset, length), ...])``."""
    parts, total = [], 0
    for a in arrays:
        a = np.ascontiguousarray(a)
        if a.dtype != np.int32 or a.ndim != 1:
            raise ValueError("pack_int32_block takes one-dimensional int32 arrays")
        parts.append((total, int(a.shape[0])))
        total += -(-int(a.shape[0]) // align) * align
    block = torch.zeros(max(total, 1), dtype=torch.int32, pin_memory=torch.cuda.is_available())
    for (o, n), a in zip(parts, arrays):
        block[o:o + n] = torch.from_numpy(np.ascontiguousarray(a))
    return block, parts


class NeighborhoodCountStep:
    """``step(centres) -> (counts[C, Q], sizes)``: canonical partition + SHMP typing + SHMP counting + count head of ``C``
    centres (a fixed number per step) as one CUDA-graph replay.  ``counts`` rows ``>= G`` are meaningless; ``sizes`` is the
    device int32 block ``{G, V, E, max_rows}``; ``result()`` synchronises and returns ``counts[:G]`` (+ the kept centres).
    A step whose neighborhoods outgrow the capacities reports ENOBUFS/ERANGE through ``result()``."""

    def __init__(self, model, graph: DeviceCSR, example_centres: torch.Tensor, depth: int, margin: float = 1.5,
                 use_cuda_graph: bool = True, centres_storage: Optional[torch.Tensor] = None):
        """``centres_storage``: caller-owned int32 [C] device tensor the step reads its centres from (instead of a
        private copy) - e.g. a view of the block ``pack_int32_block`` describes, next to the CSR it is copied with."""
        if graph.max_graph_nodes > LARGE_GRAPH_NODES:
            raise NotImplementedError("the stream-ordered step serves the small-graph partition kernels")
        emb = model.emb_model
        if PRECISION[emb.precision] == 0:
            raise NotImplementedError("the stream-ordered step serves the tensor-core precisions")
        self.model, self.graph, self.depth = model, graph, depth
        self.lib = _lib.load()
        dev = self.dev = graph.rowptr.device
        centres = example_centres.to(device=dev, dtype=torch.int32).contiguous()
        self.C = C = centres.numel()
        probe = partition_batch(graph, centres, depth, "hetero")  # eager pass: learns the capacities (one sync)
        if probe.max_rows > TILE_ROWS:
            raise NotImplementedError("neighborhoods beyond 128 rows take the eager multi-tile path")
        self.cap_rows = int(max(probe.num_rows, 64) * margin) + 64
        self.cap_edges = int(max(probe.num_edges, 64) * margin) + 64
        i32 = dict(dtype=torch.int32, device=dev)
        if centres_storage is not None:
            if (centres_storage.dtype != torch.int32 or centres_storage.numel() != C or centres_storage.device != dev
                    or not centres_storage.is_contiguous()):
                raise ValueError(f"centres_storage must be a contiguous int32 [{C}] tensor on {dev}")
            if centres_storage.data_ptr() != centres.data_ptr():
                centres_storage.copy_(centres)
            self.centres = centres_storage
        else:
            self.centres = centres.clone()
        self.nbh_ptr = torch.zeros(C + 1, **i32)
        self.centre_out = torch.zeros(C, **i32)
        self.centre_graph = torch.zeros(C, **i32)
        self.indicator = torch.zeros(C, dtype=torch.uint8, device=dev)
        self.node_gid = torch.zeros(self.cap_rows + 1, **i32)
        self.edge_ptr = torch.zeros(self.cap_rows + 2, **i32)
        self.edge_col = torch.zeros(self.cap_edges, **i32)
        self.edge_tri = torch.zeros(self.cap_edges, dtype=torch.uint8, device=dev)
        self.sizes = torch.zeros(16, **i32)
        self.status = torch.zeros(1, **i32)
        self.pw_bytes = int(self.lib.desco_partition_batch_workspace_bytes(C))
        self.pwork = torch.empty(self.pw_bytes, dtype=torch.uint8, device=dev)
        core = emb.gnn_core
        self.sw_bytes = int(self.lib.desco_shmp_workspace_bytes(self.cap_rows, C, core.layer_num))
        self.swork = torch.empty(self.sw_bytes, dtype=torch.uint8, device=dev)
        self.emb = torch.zeros((C, core.hidden_dim), dtype=torch.float32, device=dev)
        self.qe = model.get_query_emb().contiguous()
        Q = self.qe.shape[0]
        if Q > 32:
            raise NotImplementedError("the stream-ordered count head serves up to 32 queries")
        self.counts = torch.zeros((C, Q), dtype=torch.float32, device=dev)
        self.hw_bytes = int(self.lib.desco_count_head_workspace_bytes(C, Q))
        self.hwork = torch.empty(max(self.hw_bytes, 1), dtype=torch.uint8, device=dev)
        self.w = emb.packed_weights()
        self.w_head = model._head_weights()
        self.precision = PRECISION[emb.precision]
        self.pyg_bs = int(emb.pyg_batch_size)
        self.graph_exec: Optional[torch.cuda.CUDAGraph] = None
        l0 = self.lib.desco_kernel_launches()
        self._enqueue()  # warm-up outside capture (first-launch attribute calls, lazy module loads)
        self.launches_per_step = int(self.lib.desco_kernel_launches() - l0)  # kernel nodes of the captured graph
        torch.cuda.synchronize(dev)
        if use_cuda_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.graph_exec = g

    def _enqueue(self):
        lib, g, w = self.lib, self.graph, self.w
        core = self.model.emb_model.gnn_core
        st = torch.cuda.current_stream().cuda_stream
        eff = self.sizes.data_ptr() + 12 * 4  # effective {G, V, E, max_rows}
        with torch.cuda.device(self.dev):
            _lib.check(lib.desco_partition_batch_async(
                _ptr(g.rowptr), _ptr(g.col), _ptr(g.graph_ptr), g.num_graphs, _ptr(self.centres), self.C, self.depth, MODE_HETERO,
                g.max_graph_nodes, _ptr(self.pwork), self.pw_bytes, _ptr(self.nbh_ptr), _ptr(self.centre_out), _ptr(self.indicator),
                _ptr(self.centre_graph), _ptr(self.node_gid), _ptr(self.edge_ptr), self.cap_rows, _ptr(self.edge_col),
                _ptr(self.edge_tri), self.cap_edges, _ptr(self.sizes), st), "desco_partition_batch_async")
            _lib.check(lib.desco_shmp_forward_dev(
                _ptr(self.nbh_ptr), _ptr(self.edge_ptr), _ptr(self.edge_col), _ptr(self.edge_tri), self.C, self.cap_rows, eff,
                self.pyg_bs, 0, core.input_dim, _ptr(w["pre"]), _ptr(w["layers_tc"]), _ptr(w["readout"]), _ptr(w["readout_tc"]),
                core.layer_num, core.hidden_dim, _ptr(self.emb), _ptr(self.swork), self.sw_bytes, self.precision,
                _ptr(self.status), st), "desco_shmp_forward_dev")
            _lib.check(lib.desco_count_head_dev(
                _ptr(self.emb), self.C, eff, _ptr(self.qe), self.qe.shape[0], _ptr(self.w_head), core.hidden_dim, 0,
                _ptr(self.counts), _ptr(self.hwork), self.hw_bytes, st), "desco_count_head_dev")

    def __call__(self, centres: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Enqueue one step on the current stream (no synchronisation).  ``centres``: int32 [C] on any device (copied into the
        step's static buffer, asynchronously from pinned host memory) or None to reuse the last ones."""
        if centres is not None:
            if centres.numel() != self.C:
                raise ValueError(f"this step was built for {self.C} centres")
            self.centres.copy_(centres, non_blocking=True)
        if self.graph_exec is not None:
            self.graph_exec.replay()
        else:
            self._enqueue()
        return self.counts, self.sizes[12:16]

    def result(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Synchronise; ``(counts[:G], centre[:G])`` of the last step.  Raises on a capacity overflow / device error."""
        s = self.sizes.cpu()
        code = int(s[4]) or int(self.status.item())
        if code:
            _lib.check(code, "NeighborhoodCountStep (device status)")
        G = int(s[12])
        return self.counts[:G], self.centre_out[:G]
