"""Canonical neighborhood partition - host side of the sm_100a kernels in csrc/partition.cu.

Mirrors the partition surface of ``subgraph_counting/data.py`` (reference @ 4508f7a):

    k_neigh / k_neigh_canonical          data.py:329-350
    get_neigh_canonical / get_neigh_hetero  data.py:353-396   (nx.Graph in, nx.Graph out - drop-in signatures)

and adds the batched entry point the datasets use (``partition_batch``): all centres of a dataset in three kernel
launches, emitted as one packed ``NeighborhoodBatch`` resident in HBM (this also subsumes ``NetworkxToHetero`` +
``ToTconvHetero`` + PyG ``collate`` of ``workload.py:265-290`` / ``transforms.py:180-255,319-412``).

No CPU fallback: everything here raises if the CUDA library or a CUDA device is unavailable.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Union

import ctypes

import numpy as np
import torch

from . import _lib
from .graph import TargetCSR, csr_from_networkx

MODE_HETERO = 0
MODE_CANONICAL = 1
MODE_KHOP = 2


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("desco_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


@dataclass
class DeviceCSR:
    """Target graphs resident in HBM (int32 CSR, see graph.TargetCSR)."""

    rowptr: torch.Tensor
    col: torch.Tensor
    graph_ptr: torch.Tensor
    max_graph_nodes: int
    host: Optional[TargetCSR] = None

    @property
    def num_nodes(self) -> int:
        return self.rowptr.numel() - 1

    @property
    def num_directed_edges(self) -> int:
        return self.col.numel()

    @property
    def num_graphs(self) -> int:
        return self.graph_ptr.numel() - 1

    @staticmethod
    def from_host(csr: TargetCSR, device=None, non_blocking: bool = False) -> "DeviceCSR":
        dev = _require_cuda(device)
        if csr.rowptr.dtype != np.int32:
            raise ValueError("target graph exceeds int32 edge offsets")

        def up(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            return t.to(dev, non_blocking=non_blocking)

        mg = int(np.diff(csr.graph_ptr).max()) if csr.num_graphs else 1
        return DeviceCSR(up(csr.rowptr), up(csr.col), up(csr.graph_ptr), max(mg, 1), csr)


def gen_powerlaw_device(n: int, m_undirected: int, seed: int = 0, gamma: float = 2.5, max_deg_frac: float = 0.002,
                        device=None) -> DeviceCSR:
    """Config 5 built directly in HBM (same Chung-Lu recipe as ``graph.gen_powerlaw``, torch RNG on the device, so a
    10M-node / 100M-edge target takes seconds instead of minutes of host numpy).  Benchmark input generation only."""
    dev = _require_cuda(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 505)
    w = torch.arange(1, n + 1, dtype=torch.float64, device=dev) ** (-1.0 / (gamma - 1.0))
    w = w * (2.0 * m_undirected / w.sum())
    w = torch.clamp(w, max=max(max_deg_frac * n, 8.0))
    cdf = torch.cumsum(w / w.sum(), 0)
    k = int(m_undirected)
    a = torch.searchsorted(cdf, torch.rand(k, dtype=torch.float64, device=dev, generator=g)).clamp_(0, n - 1)
    b = torch.searchsorted(cdf, torch.rand(k, dtype=torch.float64, device=dev, generator=g)).clamp_(0, n - 1)
    del cdf, w
    child = torch.arange(1, n, dtype=torch.int64, device=dev)
    parent = (torch.rand(n - 1, dtype=torch.float64, device=dev, generator=g) * child).to(torch.int64)
    perm = torch.randperm(n, device=dev, generator=g)
    u = perm[torch.cat([a, child])]
    v = perm[torch.cat([b, parent])]
    del a, b, child, parent, perm
    keep = u != v
    u, v = u[keep], v[keep]
    key = torch.unique(torch.cat([u * n + v, v * n + u]))  # sorted: row-major, ascending columns, deduplicated
    del u, v, keep
    src = key // n
    col = (key - src * n).to(torch.int32)
    del key
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(src, minlength=n), 0)
    if int(rowptr[-1]) >= 2**31:
        raise ValueError("target graph exceeds int32 edge offsets")
    graph_ptr = torch.tensor([0, n], dtype=torch.int32, device=dev)
    return DeviceCSR(rowptr.to(torch.int32), col, graph_ptr, n, None)


@dataclass
class NeighborhoodBatch:
    """Packed canonical neighborhoods in HBM.  Row order inside a neighborhood = ascending node id, so the canonical
    node is the last row (``nbh_ptr[g+1]-1``) and every other row is a "count" node."""

    nbh_ptr: torch.Tensor  # int32 [G+1]
    node_gid: torch.Tensor  # int32 [V]
    edge_ptr: torch.Tensor  # int32 [V+1]
    edge_col: torch.Tensor  # int32 [E]  batch-global row of the other endpoint
    edge_tri: torch.Tensor  # uint8 [E]  1 = union_triangle, 0 = union_tride
    centre: torch.Tensor  # int32 [G]
    indicator: Optional[torch.Tensor] = None  # uint8 [#centres]  == nx_neighs_indicator (workload.py:294)
    centre_graph: Optional[torch.Tensor] = None  # int32 [#centres]
    graph_ptr: Optional[torch.Tensor] = None
    num_neighborhoods: int = 0
    num_rows: int = 0
    num_edges: int = 0
    hetero: bool = True  # False: query graphs / homogeneous neighborhoods (single node type)
    max_rows: int = 1 << 30  # host-known upper bound on the rows of any one neighborhood (selects the fused SHMP kernel)
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def num_graphs(self) -> int:  # PyG Batch vocabulary: one "graph" per neighborhood
        return self.num_neighborhoods

    def to_numpy(self) -> Dict[str, np.ndarray]:
        out = {k: getattr(self, k).cpu().numpy() for k in ("nbh_ptr", "node_gid", "edge_ptr", "edge_col", "edge_tri", "centre")}
        if self.indicator is not None:
            out["indicator"] = self.indicator.cpu().numpy().astype(bool)
        out["index"] = self.index()
        return out

    def index(self) -> np.ndarray:
        """``nx_neighs_index``: (graph id, local node id) per kept neighborhood (``workload.py:259,293``)."""
        c = self.centre.cpu().numpy().astype(np.int64)
        if self.graph_ptr is None:
            return np.stack([np.zeros_like(c), c], axis=1)
        gp = self.graph_ptr.cpu().numpy().astype(np.int64)
        gid = np.searchsorted(gp, c, side="right") - 1
        return np.stack([gid, c - gp[gid]], axis=1)

    def slice(self, g0: int, g1: int) -> "NeighborhoodBatch":
        """Neighborhoods [g0, g1) as their own packed batch (what one DataLoader step of ``batch_size`` consecutive
        neighborhoods would collate, ``lightning_data.py:78-100``).  Needs the two offsets on the host: one small sync."""
        g1 = min(g1, self.num_neighborhoods)
        r0, r1 = (int(x) for x in self.nbh_ptr[[g0, g1]].cpu())
        e0, e1 = (int(x) for x in self.edge_ptr[[r0, r1]].cpu())
        out = NeighborhoodBatch(
            self.nbh_ptr[g0:g1 + 1] - r0, self.node_gid[r0:r1], self.edge_ptr[r0:r1 + 1] - e0, self.edge_col[e0:e1] - r0,
            self.edge_tri[e0:e1], self.centre[g0:g1], None, None, self.graph_ptr, g1 - g0, r1 - r0, e1 - e0,
            hetero=self.hetero, max_rows=self.max_rows,
        )
        if "centre_last" in self._cache:
            out._cache["centre_last"] = self._cache["centre_last"]
        return out

    def select(self, idx) -> "NeighborhoodBatch":
        """The neighborhoods ``idx`` (any order, repeats allowed) gathered into their own packed batch - what one step of a
        SHUFFLED DataLoader collates (``lightning_data.py:78-100`` with ``shuffle=True``).  Device-side gathers; the row and
        edge totals are read on the host (one small sync)."""
        dev = self.nbh_ptr.device
        idx = torch.as_tensor(idx, device=dev).long().reshape(-1)
        S = idx.numel()
        ar = lambda n: torch.arange(n, device=dev)
        ptr = self.nbh_ptr.long()
        r0 = ptr[idx]
        rows = ptr[idx + 1] - r0
        new_ptr = torch.cat([rows.new_zeros(1), torch.cumsum(rows, 0)])
        V2 = int(new_ptr[-1])
        seg = torch.repeat_interleave(ar(S), rows, output_size=V2)       # selected neighborhood of each new row
        old_row = r0[seg] + (ar(V2) - new_ptr[seg])
        ep = self.edge_ptr.long()
        e_first = ep[old_row]
        deg = ep[old_row + 1] - e_first
        new_eptr = torch.cat([deg.new_zeros(1), torch.cumsum(deg, 0)])
        E2 = int(new_eptr[-1])
        rseg = torch.repeat_interleave(ar(V2), deg, output_size=E2)       # new row of each new edge
        old_e = e_first[rseg] + (ar(E2) - new_eptr[rseg])
        shift = (new_ptr[:-1] - r0)[seg][rseg]                             # rows of a neighborhood move together
        out = NeighborhoodBatch(
            new_ptr.to(torch.int32), self.node_gid[old_row], new_eptr.to(torch.int32),
            (self.edge_col[old_e].long() + shift).to(torch.int32), self.edge_tri[old_e], self.centre[idx], None, None,
            self.graph_ptr, S, V2, E2, hetero=self.hetero, max_rows=self.max_rows,
        )
        if "centre_last" in self._cache:
            out._cache["centre_last"] = self._cache["centre_last"]
        return out

    @staticmethod
    def from_numpy(d: Dict[str, np.ndarray], device=None, hetero: bool = True) -> "NeighborhoodBatch":
        dev = _require_cuda(device)
        t = lambda k, dt: torch.from_numpy(np.ascontiguousarray(d[k]).astype(dt)).to(dev)
        V = int(d["nbh_ptr"][-1])
        node_gid = t("node_gid", np.int32) if "node_gid" in d else torch.arange(V, dtype=torch.int32, device=dev)
        centre = t("centre", np.int32) if "centre" in d else node_gid[(t("nbh_ptr", np.int64)[1:] - 1)]
        sizes = np.diff(np.asarray(d["nbh_ptr"]))
        return NeighborhoodBatch(
            t("nbh_ptr", np.int32), node_gid, t("edge_ptr", np.int32), t("edge_col", np.int32), t("edge_tri", np.uint8),
            centre, None, None, None, len(d["nbh_ptr"]) - 1, V, int(d["edge_ptr"][-1]), hetero,
            max_rows=int(sizes.max()) if len(sizes) else 0,
        )


LARGE_GRAPH_NODES = 409_600  # above this the four node bitsets of a centre no longer fit shared memory


def partition_batch(graph: DeviceCSR, centres: Optional[torch.Tensor], depth: int, mode: Union[int, str] = "hetero",
                    large: Optional[bool] = None) -> NeighborhoodBatch:
    """All canonical neighborhoods of ``centres`` (default: every node, in dataset order) in one packed batch.

    Replaces the double Python loop of ``NeighborhoodDataset.process`` (``workload.py:250-272``).  ``large`` selects
    the sparse (hash-set) kernels of the large-graph regime; default: by ``graph.max_graph_nodes``."""
    lib = _lib.load()
    dev = graph.rowptr.device
    if large is None:
        large = graph.max_graph_nodes > LARGE_GRAPH_NODES
    if isinstance(mode, str):
        mode = {"hetero": MODE_HETERO, "canonical": MODE_CANONICAL, "khop": MODE_KHOP}[mode]
    if centres is None:
        centres = torch.arange(graph.num_nodes, dtype=torch.int32, device=dev)
    centres = centres.to(device=dev, dtype=torch.int32).contiguous()
    C = centres.numel()
    i32 = dict(dtype=torch.int32, device=dev)
    if not large:
        return _partition_batch_one_call(lib, graph, centres, C, depth, mode)
    scratch = torch.empty((6, max(C, 1)), **i32)  # nv, ne, centre_graph, keep_rank, node_off, edge_off
    nv, ne, cg, rank, noff, eoff = scratch.unbind(0)
    nbh_ptr = torch.empty(C + 1, **i32)
    centre_out = torch.empty(max(C, 1), **i32)
    indicator = torch.empty(max(C, 1), dtype=torch.uint8, device=dev)
    small = torch.zeros(4, **i32)  # totals[3] + status
    totals, status = small[:3], small[3:]
    with torch.cuda.device(dev):
        st = _stream()
        if large:
            lbytes = int(lib.desco_partition_large_workspace_bytes(graph.max_graph_nodes, C))
            lwork = torch.empty(lbytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.desco_partition_large_count(
                _ptr(graph.rowptr), _ptr(graph.col), _ptr(graph.graph_ptr), graph.num_graphs, _ptr(centres), C, depth,
                mode, graph.max_graph_nodes, _ptr(nv), _ptr(ne), _ptr(cg), _ptr(status), _ptr(lwork), lbytes, st),
                "desco_partition_large_count")
        else:
            _lib.check(lib.desco_partition_count(
                _ptr(graph.rowptr), _ptr(graph.col), _ptr(graph.graph_ptr), graph.num_graphs, _ptr(centres), C, depth,
                mode, graph.max_graph_nodes, _ptr(nv), _ptr(ne), _ptr(cg), _ptr(status), st), "desco_partition_count")
        wbytes = int(lib.desco_partition_scan_workspace_bytes(C))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.desco_partition_scan(
            _ptr(centres), _ptr(nv), _ptr(ne), C, _ptr(rank), _ptr(noff), _ptr(eoff), _ptr(nbh_ptr), _ptr(centre_out),
            _ptr(indicator), _ptr(totals), _ptr(work), wbytes, st), "desco_partition_scan")
        mx = nv[:C].max().reshape(1) if C else torch.zeros(1, **i32)
        sums = torch.stack([nv[:C].sum(dtype=torch.int64), ne[:C].sum(dtype=torch.int64)])  # exact: the scans are int32
        G, V, E, code, max_nv, V64, E64 = (int(x) for x in torch.cat([small.long(), mx.long(), sums]).cpu())  # the one host sync
        if code != 0:
            _lib.check(code, "partition kernel (device status)")
        if V64 >= 2**31 or E64 >= 2**31:
            _lib.check(_lib.ERANGE, f"partition_batch: {V64} rows / {E64} edges do not fit one packed batch (int32 offsets); "
                                    "split the centre list (partition_batches does)")
        node_gid = torch.empty(V, **i32)
        edge_ptr = torch.zeros(V + 1, **i32)
        edge_col = torch.empty(E, **i32)
        edge_tri = torch.empty(E, dtype=torch.uint8, device=dev)
        if V > 0 and large:
            _lib.check(lib.desco_partition_large_fill(
                _ptr(graph.rowptr), _ptr(graph.col), _ptr(graph.graph_ptr), graph.num_graphs, _ptr(centres), C, depth,
                mode, graph.max_graph_nodes, _ptr(nv), _ptr(ne), _ptr(cg), _ptr(noff), _ptr(eoff), V, _ptr(node_gid),
                _ptr(edge_ptr), _ptr(edge_col), _ptr(edge_tri), _ptr(status), _ptr(lwork), lbytes, st),
                "desco_partition_large_fill")
        elif V > 0:
            _lib.check(lib.desco_partition_fill(
                _ptr(graph.rowptr), _ptr(graph.col), _ptr(graph.graph_ptr), graph.num_graphs, _ptr(centres), C, depth,
                mode, graph.max_graph_nodes, _ptr(nv), _ptr(ne), _ptr(cg), _ptr(noff), _ptr(eoff), _ptr(node_gid),
                _ptr(edge_ptr), _ptr(edge_col), _ptr(edge_tri), _ptr(status), st), "desco_partition_fill")
    batch = NeighborhoodBatch(
        nbh_ptr[: G + 1], node_gid, edge_ptr, edge_col, edge_tri, centre_out[:G], indicator[:C], cg[:C],
        graph.graph_ptr, G, V, E, hetero=(mode == MODE_HETERO),
        max_rows=max_nv,  # rows of the largest neighborhood (selects the fused SHMP kernel)
    )
    if mode == MODE_CANONICAL:
        batch._cache["centre_last"] = True  # get_neigh_canonical marks the centre (last row) with node_feature = 1
    if large:
        batch._cache["tier"] = lwork[:C].clone()  # which tier served each centre (0 shared-memory hash, 1 team bitmap)
    return batch


def partition_batches(graph: DeviceCSR, centres: Optional[torch.Tensor], depth: int, mode: Union[int, str] = "hetero",
                      max_centres: Optional[int] = None, large: Optional[bool] = None):
    """``partition_batch`` over consecutive chunks of ``centres`` (generator of packed batches, in centre order).

    A packed batch addresses rows and edges with int32, and on a power-law target the depth-2 ball of a hub holds 1e5-1e6
    rows, so a long centre list (a rank's whole centre range, ``distributed.ShardedPipeline``) cannot go through one
    call.  ``max_centres`` bounds a chunk (default: 4096 in the large-graph regime, 1 << 20 otherwise); a chunk whose
    exact row or edge total still overflows (``DESCO_ERANGE``) is halved and retried."""
    dev = graph.rowptr.device
    if large is None:
        large = graph.max_graph_nodes > LARGE_GRAPH_NODES
    if centres is None:
        centres = torch.arange(graph.num_nodes, dtype=torch.int32, device=dev)
    centres = centres.to(device=dev, dtype=torch.int32).contiguous()
    step = int(max_centres) if max_centres else (4096 if large else 1 << 20)
    todo = [(a, min(a + step, centres.numel())) for a in range(0, centres.numel(), step)][::-1]
    while todo:
        a, b = todo.pop()
        try:
            batch = partition_batch(graph, centres[a:b], depth, mode, large)
        except _lib.DescoError as e:
            if e.code != _lib.ERANGE or b - a <= 1:
                raise
            mid = (a + b) // 2
            todo += [(mid, b), (a, mid)]
            continue
        yield batch


def partition_sizes(graph: DeviceCSR, centres: Optional[torch.Tensor], depth: int, mode: Union[int, str] = "hetero",
                    large: Optional[bool] = None, max_centres: Optional[int] = None):
    """Streaming count-only mode: rows and directed edges of every centre's canonical neighborhood, nothing emitted.

    Only the count pass of the partition runs (``desco_partition_count`` / ``desco_partition_large_count``), chunk by
    chunk, so a centre list of any length goes through in bounded memory - what a full sweep over a 10^7-node target
    needs before anything is packed (the sizes decide the chunking: ``partition_batches``) and what the reference's
    dataset statistics compute by building every neighborhood (``analysis/dataset_statistics.py:53`` ->
    ``data.py:375-396``).  Returns (rows [C] int32, directed_edges [C] int32) in centre order; an edge-free
    neighborhood - which ``NeighborhoodDataset.process`` drops, ``workload.py:253-256`` - reads (0, 0)."""
    lib = _lib.load()
    dev = graph.rowptr.device
    if large is None:
        large = graph.max_graph_nodes > LARGE_GRAPH_NODES
    if isinstance(mode, str):
        mode = {"hetero": MODE_HETERO, "canonical": MODE_CANONICAL, "khop": MODE_KHOP}[mode]
    if centres is None:
        centres = torch.arange(graph.num_nodes, dtype=torch.int32, device=dev)
    centres = centres.to(device=dev, dtype=torch.int32).contiguous()
    C = centres.numel()
    i32 = dict(dtype=torch.int32, device=dev)
    nv, ne = torch.zeros(max(C, 1), **i32), torch.zeros(max(C, 1), **i32)
    step = int(max_centres) if max_centres else (4096 if large else 1 << 20)
    cg = torch.empty(min(step, max(C, 1)), **i32)
    status = torch.zeros(1, **i32)
    with torch.cuda.device(dev):
        st = _stream()
        lwork, lbytes = None, 0
        if large:
            lbytes = int(lib.desco_partition_large_workspace_bytes(graph.max_graph_nodes, min(step, max(C, 1))))
            lwork = torch.empty(lbytes, dtype=torch.uint8, device=dev)
        for a in range(0, C, step):
            n = min(step, C - a)
            if large:
                _lib.check(lib.desco_partition_large_count(
                    _ptr(graph.rowptr), _ptr(graph.col), _ptr(graph.graph_ptr), graph.num_graphs, _ptr(centres[a:]), n, depth,
                    mode, graph.max_graph_nodes, _ptr(nv[a:]), _ptr(ne[a:]), _ptr(cg), _ptr(status), _ptr(lwork), lbytes, st),
                    "desco_partition_large_count")
            else:
                _lib.check(lib.desco_partition_count(
                    _ptr(graph.rowptr), _ptr(graph.col), _ptr(graph.graph_ptr), graph.num_graphs, _ptr(centres[a:]), n, depth,
                    mode, graph.max_graph_nodes, _ptr(nv[a:]), _ptr(ne[a:]), _ptr(cg), _ptr(status), st), "desco_partition_count")
        code = int(status.item())
    if code != 0:
        _lib.check(code, "partition kernel (device status)")
    return nv[:C], ne[:C]


_CAPACITY = {"rows_per_centre": 24.0, "edges_per_row": 6.0}  # running maxima of the batches seen so far


def _partition_batch_one_call(lib, graph: DeviceCSR, centres: torch.Tensor, C: int, depth: int, mode: int) -> NeighborhoodBatch:
    """count -> scans -> (one stream sync for the sizes) -> fill inside ONE C call (``desco_partition_batch``).  The packed
    batch is emitted into buffers sized from the largest rows-per-centre / edges-per-row ratios seen so far; the call
    reports ENOBUFS with the exact sizes when they are too small and is then repeated once."""
    dev = graph.rowptr.device
    i32 = dict(dtype=torch.int32, device=dev)
    per = torch.empty((3, max(C, 1) + 1), **i32)  # nbh_ptr, centre_out, centre_graph
    nbh_ptr, centre_out, cg = per[0], per[1, :max(C, 1)], per[2, :max(C, 1)]
    indicator = torch.empty(max(C, 1), dtype=torch.uint8, device=dev)
    wbytes = int(lib.desco_partition_batch_workspace_bytes(C))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    totals = (ctypes.c_int32 * 4)()
    cap_rows = int(min(C * min(graph.max_graph_nodes, _CAPACITY["rows_per_centre"] * 1.25) + 64, 2**31 - 2))
    cap_edges = int(min(cap_rows * _CAPACITY["edges_per_row"] * 1.25 + 64, 2**31 - 2))
    with torch.cuda.device(dev):
        st = _stream()
        for attempt in range(2):
            rows = torch.empty((2, cap_rows + 1), **i32)  # node_gid, edge_ptr
            edge_col = torch.empty(cap_edges, **i32)
            edge_tri = torch.empty(cap_edges, dtype=torch.uint8, device=dev)
            rc = lib.desco_partition_batch(
                _ptr(graph.rowptr), _ptr(graph.col), _ptr(graph.graph_ptr), graph.num_graphs, _ptr(centres), C, depth, mode,
                graph.max_graph_nodes, _ptr(work), wbytes, _ptr(nbh_ptr), _ptr(centre_out), _ptr(indicator), _ptr(cg),
                _ptr(rows[0]), _ptr(rows[1]), cap_rows, _ptr(edge_col), _ptr(edge_tri), cap_edges, totals, st)
            G, V, E, max_nv = (int(x) for x in totals)
            if rc != _lib.ENOBUFS or attempt == 1:
                break
            cap_rows, cap_edges = V, E
        _lib.check(rc, "desco_partition_batch")
    if C and V:
        _CAPACITY["rows_per_centre"] = max(_CAPACITY["rows_per_centre"], V / C)
        _CAPACITY["edges_per_row"] = max(_CAPACITY["edges_per_row"], E / V)
    batch = NeighborhoodBatch(
        nbh_ptr[: G + 1], rows[0, :V], rows[1, : V + 1], edge_col[:E], edge_tri[:E], centre_out[:G], indicator[:C], cg[:C],
        graph.graph_ptr, G, V, E, hetero=(mode == MODE_HETERO), max_rows=max_nv,
    )
    if mode == MODE_CANONICAL:
        batch._cache["centre_last"] = True  # get_neigh_canonical marks the centre (last row) with node_feature = 1
    return batch


def shmp_edge_types(edge_ptr: torch.Tensor, edge_col: torch.Tensor) -> torch.Tensor:
    """Triangle / tride flag of every directed edge of a packed batch (``transforms.py:201-225``)."""
    lib = _lib.load()
    _require_cuda(edge_ptr.device)
    tri = torch.empty(edge_col.numel(), dtype=torch.uint8, device=edge_ptr.device)
    with torch.cuda.device(edge_ptr.device):
        _lib.check(lib.desco_shmp_edge_types(_ptr(edge_ptr), _ptr(edge_col), edge_ptr.numel() - 1, _ptr(tri), _stream()),
                   "desco_shmp_edge_types")
    return tri


# ---------------------------------------------------------------------------------------------
# drop-in single-centre signatures (nx.Graph in, nx.Graph out)
# ---------------------------------------------------------------------------------------------


def _to_nx_graph(graph):
    import networkx as nx

    if isinstance(graph, nx.Graph):
        return graph
    if hasattr(graph, "edge_index") and hasattr(graph, "num_nodes"):  # PyG Data duck type (data.py:358-359,380-381)
        g = nx.Graph()
        g.add_nodes_from(range(int(graph.num_nodes)))
        ei = graph.edge_index.cpu().numpy()
        g.add_edges_from((int(a), int(b)) for a, b in ei.T if a <= b)
        return g
    raise TypeError(f"unsupported graph type {type(graph)}")


def _device_graph_for(graph):
    """nx graph with arbitrary integer node labels -> (DeviceCSR over dense ids that preserve label order, labels)."""
    labels = sorted(graph.nodes)
    if labels and (labels[0] != 0 or labels[-1] != len(labels) - 1):
        import networkx as nx

        dense = nx.relabel_nodes(graph, {u: i for i, u in enumerate(labels)}, copy=True)
    else:
        dense = graph
    return DeviceCSR.from_host(csr_from_networkx([dense])), labels


def _neigh_as_nx(graph, labels, batch: NeighborhoodBatch, mark: str):
    import networkx as nx

    b = batch.to_numpy()
    out = nx.Graph()
    if batch.num_neighborhoods == 0:  # edge-free neighborhood: the reference still returns the lone centre
        return None
    rows = [labels[int(g)] for g in b["node_gid"]]
    for u in rows:
        out.add_node(u, **graph.nodes[u])
    for r, u in enumerate(rows):
        for e in range(int(b["edge_ptr"][r]), int(b["edge_ptr"][r + 1])):
            out.add_edge(u, rows[int(b["edge_col"][e])])
    return out


def _k_hop(graph, start_node, k, mode):
    graph = _to_nx_graph(graph)
    dg, labels = _device_graph_for(graph)
    c = torch.tensor([labels.index(start_node)], dtype=torch.int32, device=dg.rowptr.device)
    batch = partition_batch(dg, c, k, mode)
    return graph, labels, batch


def get_neigh_hetero(graph, node, radius: int):
    """Drop-in for ``data.py:375-396``: nx.Graph of the canonical neighborhood, node attr ``type``."""
    graph, labels, batch = _k_hop(graph, node, radius, MODE_HETERO)
    neigh = _neigh_as_nx(graph, labels, batch, "type")
    if neigh is None:
        import networkx as nx

        neigh = nx.Graph()
        neigh.add_node(node, **graph.nodes[node])
    for u in neigh.nodes:
        neigh.nodes[u]["type"] = "count"
    neigh.nodes[node]["type"] = "canonical"
    return neigh


def get_neigh_canonical(graph, node, radius: int):
    """Drop-in for ``data.py:353-372``: restricted-BFS variant, ``node_feature`` = 1 at the centre."""
    graph, labels, batch = _k_hop(graph, node, radius, MODE_CANONICAL)
    neigh = _neigh_as_nx(graph, labels, batch, "node_feature")
    if neigh is None:
        import networkx as nx

        neigh = nx.Graph()
        neigh.add_node(node, **graph.nodes[node])
    for u in neigh.nodes:
        neigh.nodes[u]["node_feature"] = torch.zeros(1)
    neigh.nodes[node]["node_feature"] = torch.ones(1)
    return neigh


def k_neigh(G, start_node, k):
    """Drop-in for ``data.py:329-338`` (node list of the unrestricted k-hop ball)."""
    graph, labels, batch = _k_hop(G, start_node, k, MODE_KHOP)
    if batch.num_neighborhoods == 0:
        return [start_node]
    return [labels[int(g)] for g in batch.node_gid.cpu().numpy()]


def k_neigh_canonical(G, start_node, k):
    """Drop-in for ``data.py:341-350`` (node list of the restricted k-hop BFS)."""
    return list(get_neigh_canonical(G, start_node, k).nodes)
