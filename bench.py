#!/usr/bin/env python
"""bench.py - headline benchmark of the DeSCo inference hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 (the BENCH line): BASELINE.json configs[1] - 4096 canonical 4-hop neighborhoods of the ENZYMES-shaped synthetic
pool x 29 atlas queries.  One "step" = canonical partition + SHMP typing -> SHMP forward (8 fused layers + readout) ->
query-conditioned count head, for the whole 4096-neighborhood batch.  `value` = neighborhoods/s with the target CSR and
centre list resident in HBM; `e2e` = the same through the public Python API from pinned HOST buffers (CSR + centres
copied H2D, counts copied D2H inside the timed region).  The line also carries the gossip half of the metric on a
1M-node power-law graph (`gossip`) and the 1-GPU base of the multi-GPU workload (`config5`).

N > 1 (the SCALE lines; one process per GPU under torchrun): BASELINE.json configs[4], the north-star multi-GPU split,
as STRONG scaling of ONE 10M-node / 100M-undirected-edge power-law target replicated on every GPU (0.84 GB of CSR):
  * `value` = gossip target-nodes/s: one step = the node-range-sharded GossipCountingModel forward over the WHOLE graph
    for all 29 queries, with the halo all-gathers of the layer-0 scalars (one per query group, pipelined under layer 1 of
    the previous group) and the all-gathers of the output rows INSIDE the timed region (NCCL over NVLink);
  * `config5.partition_count` = depth-2 canonical partition + SHMP typing + SHMP counting of a fixed seeded sample of
    centres, sharded by centre range (ranges balanced by estimated work), no collective;
  * parity of both against the CPU oracle on seeded samples is asserted outside the timed region and printed.

Timing: CUDA events on the launching stream around every step, max over ranks; config 2: L2 flushed (256 MiB write)
between steps and excluded; config 5: inputs (1.16 GB of counts, 4.6 GB of halo scalars) are far larger than L2.
Roofline: the dominant kernel timed live by the library's own CUDA-event hooks (include/desco_b200.h desco_profile_*).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "enzymes_shaped_4096nbh_depth4_x29queries"
DEPTH = 4
NUM_NBH = 4096
METRIC = "canonical_neighborhoods_per_sec"
UNIT = "neighborhoods/s"
GOSSIP_METRIC = "gossip_target_nodes_per_sec"
GOSSIP_UNIT = "target-nodes/s"
PROF_SLOTS = 6
TOL = 1e-4


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        for k in ("bf16_tflops_sustained", "bf16_tflops"):  # the chain is timed inside a long forward: sustained figure
            if k in d:
                return float(d[k]), f"measured (MEASURED_PEAKS.json {k})"
    return 1590.0, "fallback (B200_PROFILING.md)"


def gather_roofline(nodes, directed_edges, queries, world, gather_ms):
    """HBM roofline of gossip_gather_kernel on the bytes its own formulation has to move (per rank and forward): a
    16-byte record + a 4-byte adjacency entry per (edge, query) read, and the tile's operand images (u and x1 as bf16 hi /
    lo rows + c + d1 = 520 B per (node, query)) written for the chain kernel.  Pure function: tests/test_bench_contract_cpu.py."""
    peak, src = _peaks()
    alg = (20.0 * directed_edges * queries + 520.0 * nodes * queries) / max(world, 1)
    ach = alg / (gather_ms * 1e-3) / 1e9 if gather_ms > 0 else 0.0
    return {"kernel": "gossip_gather_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "peak_source": src, "algorithmic_bytes_per_step_per_rank": alg, "ms_per_step": gather_ms,
            "traffic": None,
            "note": "bytes of the kernel's OWN formulation (records are re-read across tiles through L2, so DRAM traffic is "
                    "lower: ncu 0.88 GB per 8192-tile launch, profiles/r2_ncu_gossip_gather_summary.txt); what bounds the "
                    "kernel is the warp-level tensor pipe + issue, profiles/r2_gossip_gather_ablation.txt"}


def _num_sms(torch):
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def build_workload(seed: int):
    from desco_b200.graph import first_nonempty_centres, gen_enzymes_shaped

    csr = gen_enzymes_shaped(seed=seed)
    centres = first_nonempty_centres(csr, NUM_NBH)
    assert len(centres) == NUM_NBH
    return csr, centres


# ---------------------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference): the cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------------------------
def cpu_step(csr, centres, om, qb):
    """Reference-style CPU step on `centres`: networkx partition loop (single process, like workload.py:250), literal
    sparse A*A@A+A typing, PyTorch CPU SHMP forward + per-query count head."""
    import torch

    from oracle import partition as P
    from oracle.shmp_types import type_batch

    b = P.partition_dataset(csr, DEPTH, mode="hetero", centres=centres, with_types=False)
    b["edge_tri"] = type_batch(b)
    with torch.no_grad():
        return b, om.graph_to_count(b, qb, pyg_batch_size=512)


def cpu_models():
    import torch

    from oracle import model as M

    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    return M.NeighborhoodCountingModel().eval(), M.query_batch()


def cpu_gossip_workload(nodes=5_000, edges=50_000):
    """Bounded CPU sample of the config-5 gossip workload: a power-law graph of the same recipe, 29 queries."""
    import torch

    from desco_b200.graph import gen_powerlaw
    from oracle import model as M

    csr = gen_powerlaw(nodes, edges, seed=0)
    torch.manual_seed(1)
    torch.set_num_threads(os.cpu_count() or 1)
    og = M.GossipCountingModel()
    g = torch.Generator().manual_seed(7)
    x = torch.floor(torch.exp(torch.randn(csr.num_nodes, 29, generator=g)))
    og.set_query_emb(torch.randn(29, 64, generator=g))
    return csr, og, x, torch.from_numpy(csr.edge_index())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import torch

    if world > 1 or args.gpus > 1:  # the multi-GPU arm's workload: gossip over the power-law target, CPU oracle on a bounded sample
        csr, og, x, ei = cpu_gossip_workload()
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            with torch.no_grad():
                og.graph_to_count(x, ei)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        total = sum(times)
        value = csr.num_nodes * len(times) / total
        sample = (f"one gossip forward (29 queries, literal per-edge lin_com) over a {csr.num_nodes}-node / "
                  f"{csr.num_directed_edges // 2}-undirected-edge power-law graph of the config-5 recipe per step")
        line = {
            "impl": "reference", "metric": GOSSIP_METRIC, "value": value, "unit": GOSSIP_UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": config5_name(args), "queries": 29, "sample_per_step": sample,
                       "note": "oracle port of the reference CPU path (gnn_model.py:231-359 restated in torch CPU); the "
                               "reference itself needs torch_geometric/pytorch_lightning which are not installable here"},
            "cpu_baseline": {"value": value, "unit": GOSSIP_UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": GOSSIP_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return
    csr, centres = build_workload(0)
    om, qb = cpu_models()
    sample = 512
    times = []
    for i in range(args.warmup + args.steps):
        lo = (i * sample) % NUM_NBH
        t0 = time.perf_counter()
        cpu_step(csr, centres[lo:lo + sample], om, qb)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "depth": DEPTH, "queries": 29, "sample_per_step": sample,
                   "note": "oracle port of the reference CPU path (networkx partition + sparse typing + torch CPU SHMP); "
                           "the reference itself needs torch_geometric/pytorch_lightning which are not installable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{sample} of the {NUM_NBH} neighborhoods per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# GPU path
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        busy = [s for s in sm if mx and s >= 0.5 * max(mx)]  # samples taken while the GPU was clocked up (under load)
        return {"sm_mhz": float(np.median(busy or sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Ctx:
    """Per-process bench context: device, ranks, library, timing helpers."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        from desco_b200 import _lib

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = _lib.load()
        self.flush = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, v, op="max"):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            ops = {"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}
            self.dist.all_reduce(t, op=ops[op])
        return float(t.item())

    def timed(self, fn, steps, warmup, profile=False, flush_l2=True):
        """W untimed warm-ups, then exactly K steps, each bracketed by a CUDA-event pair on the launching stream, the
        whole loop by barrier + synchronize; returns (sum of step ms, max over ranks), launches, profile slots."""
        torch = self.torch
        if flush_l2 and self.flush is None:
            self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        for _ in range(warmup):
            if flush_l2:
                self.flush.fill_(1)
            fn()
        self.barrier()
        if profile:
            self.lib.desco_profile_enable(1)
        launches0 = self.lib.desco_kernel_launches()
        evs = []
        for _ in range(steps):
            if flush_l2:
                self.flush.fill_(1)  # L2 flush, outside the timed events
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        self.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = self.lib.desco_kernel_launches() - launches0
        prof = None
        if profile:
            pm = (ctypes.c_double * PROF_SLOTS)()
            pl = (ctypes.c_int64 * PROF_SLOTS)()
            self.lib.desco_profile_read(pm, pl)
            self.lib.desco_profile_enable(0)
            prof = (list(pm), list(pl))
        return self.reduce(ms, "max"), launches, prof


def _rel_err(a, b):
    """max |a - b| / max(1, |b|); equal entries (including matching infinities of 2^pred) count as 0."""
    a, b = a.double(), b.double()
    d = (a - b).abs() / b.abs().clamp(min=1.0)
    d[a == b] = 0.0
    return float(d.max().item()) if d.numel() else 0.0


def run_gossip_leg(ctx, args, model):
    """Second half of BASELINE.json's metric at N = 1: gossip target-nodes/s.  One step = GossipCountingModel.graph_to_count
    over a power-law target graph (config-5 recipe, default 1M nodes / 10M undirected edges) for all 29 queries: both
    GossipConv layers + post_mp.  Inputs (CSR, per-node counts x[N,29], query embeddings) resident in HBM."""
    torch, dev, lib = ctx.torch, ctx.dev, ctx.lib
    from desco_b200.data import gen_powerlaw_device
    from desco_b200.lightning_model import GossipCountingModel

    g = gen_powerlaw_device(args.gossip_nodes, args.gossip_edges, seed=ctx.rank, device=dev)
    N, M = g.num_nodes, g.num_directed_edges
    torch.manual_seed(1)
    gm = GossipCountingModel().eval().to(dev)
    qe = model.get_query_emb()
    gm.set_query_emb(qe)
    Q = qe.shape[0]
    gen = torch.Generator(device=dev)
    gen.manual_seed(7 + ctx.rank)
    x = torch.floor(torch.exp(torch.randn((N, Q), device=dev, generator=gen)))  # floor(exp(N(0,1)))

    def step():
        with torch.no_grad():
            return gm.emb_model.forward_all_queries(g.rowptr, g.col, x, qe)

    steps = max(3, min(args.steps, 10))
    ms, launches, _ = ctx.timed(step, steps, 3, flush_l2=False)
    _, _, prof = ctx.timed(step, steps, 1, profile=True, flush_l2=False)
    if ctx.rank != 0:
        return None
    tpeak, tsrc = _tensor_peak()
    l0_ms, gather_ms, chain_ms = (prof[0][i] / steps for i in (3, 4, 5))
    mm_flops = 2.0 * (128 * 64 + 128 * 64 + 64 * 64 + 64 * 256) * N * Q  # the four GEMMs of the chain, per (node, query)
    alg = Q * (512 * M + 1280 * N) + 8 * Q * N + 8 * M  # reference formulation, fp32 [.,64] rows
    return {
        "metric": GOSSIP_METRIC, "value": ctx.world * N * steps / (ms * 1e-3), "unit": GOSSIP_UNIT,
        "ms_per_step": ms / steps, "steps": steps,
        "workload": f"powerlaw_chunglu_{N}nodes_{M // 2}undirected_edges_x{Q}queries_per_gpu", "nodes": N,
        "directed_edges": M, "queries": Q, "gpu_launches": int(launches), "input": "resident in HBM, far larger than L2",
        "stage_ms": {"layer0_scalar_sweep": l0_ms, "layer1_gated_sweep_gather": gather_ms, "layer1_postmp_tcgen05_chain": chain_ms,
                     "timed_in": "second pass of the same steps, CUDA events around every launch"},
        "precision": "bf16x3 tcgen05 GEMM chain, fp32 accumulate (1e-4 parity path)",
        "roofline": {"kernel": "gossip_chain_kernel", "bound": "tensor", "unit": "TFLOP/s",
                     "achieved": mm_flops / (chain_ms * 1e-3) / 1e12, "peak": tpeak, "frac": mm_flops / (chain_ms * 1e-3) / 1e12 / tpeak,
                     "peak_source": tsrc, "traffic": None,
                     "note": "USEFUL fp32-equivalent flops of the four GEMMs (the kernel issues 3x as many bf16 flops for the "
                             "hi/lo split) over the chain kernel's own CUDA-event time"},
        "gather": {"kernel": "gossip_gather_kernel",
                   "bound": "warp-level tensor pipe + issue: the neighbours' rows are recomputed as tf32 mma.sync blocks (8 "
                            "neighbour records x 64 features = 8 HMMA.1688 at 8.6 cycles each per sub-partition) with relu + add "
                            "on the CUDA cores; profiles/r2_gossip_gather_ablation.txt",
                   "ms_per_step": gather_ms, "edge_query_pairs_per_s": M * Q / (gather_ms * 1e-3),
                   "roofline": gather_roofline(N, M, Q, 1, gather_ms),
                   "mma_blocks_floor_ms": (M * Q / 8.0) * 8 * 8.6 / (4 * _num_sms(torch) * 1.965e9) * 1e3},
        "reference_formulation_bytes_per_step": alg,
        "note": "the reference formulation moves a 64-wide fp32 row per edge and query; these kernels move 16 B per edge and "
                "query and recompute the row, so bytes / time against that formulation is not a roofline and is not reported",
    }


def replicate_inputs(ctx, tensors, what):
    """The target CSR and the counts are generated on every rank from the same seeds.  Check that the copies really are
    bit-identical (sizes + two digests per tensor, all-gathered) and, if any rank's differs, take rank 0's: a
    strong-scaling job shards ONE input.  Returns (tensors, what was found)."""
    torch, dist, dev = ctx.torch, ctx.dist, ctx.dev
    if ctx.world == 1:
        return tensors, {"what": what, "checked": False}

    def digest(t):
        v = t.contiguous().view(-1).view(torch.int32).to(torch.int64)
        w = (torch.arange(v.numel(), device=dev, dtype=torch.int64) % 65521) + 1
        return [int(t.numel()), int(v.sum().item()), int((v * w).sum().item())]

    sig = torch.tensor([d for t in tensors for d in digest(t)], dtype=torch.int64, device=dev)
    allsig = [torch.zeros_like(sig) for _ in range(ctx.world)]
    dist.all_gather(allsig, sig)
    differing = [r for r in range(ctx.world) if not torch.equal(allsig[r], allsig[0])]
    if differing:
        out = []
        for i, t in enumerate(tensors):
            n0 = int(allsig[0][3 * i].item())
            if t.dim() == 1 and t.numel() != n0:
                t = torch.empty(n0, dtype=t.dtype, device=dev)
            dist.broadcast(t, 0)
            out.append(t)
        tensors = out
    return tensors, {"what": what, "checked": True, "ranks_whose_generated_copy_differed_from_rank0": differing,
                     "action": "rank 0's copy broadcast to all ranks" if differing else "none"}


def config5_name(args):
    return f"powerlaw_chunglu_{args.c5_nodes}nodes_{args.c5_edges}undirected_edges_depth2_x29queries"


def run_config5(ctx, args, nm, steps, warmup):
    """BASELINE.json configs[4] as STRONG scaling (see the module docstring).  Returns the per-rank-0 result dict."""
    torch, dist, dev, lib, rank, world = ctx.torch, ctx.dist, ctx.dev, ctx.lib, ctx.rank, ctx.world
    from desco_b200.data import gen_powerlaw_device
    from desco_b200.distributed import ShardedPipeline
    from desco_b200.lightning_model import GossipCountingModel

    depth = 2
    g = gen_powerlaw_device(args.c5_nodes, args.c5_edges, seed=0, device=dev)  # same seed: the CSR is replicated ...
    (g.rowptr, g.col), rep_graph = replicate_inputs(ctx, [g.rowptr, g.col], "target CSR")  # ... and checked to be
    N, M = g.num_nodes, g.num_directed_edges
    torch.manual_seed(1)
    gm = GossipCountingModel().eval().to(dev)
    qe = nm.get_query_emb()
    gm.set_query_emb(qe)
    Q = qe.shape[0]
    nm.set_pyg_batch_size(512)
    pipe = ShardedPipeline(g, nm, gm, None, depth=depth)

    # ---------------- partition + SHMP count: fixed seeded centre sample, sharded by centre range ----------------
    rng = np.random.default_rng(11)
    sample = np.sort(rng.choice(N, size=min(args.c5_centres, N), replace=False))
    lo, hi = pipe.centre_shards[rank]
    mine = torch.as_tensor(pipe.deal_centres(sample), dtype=torch.int32, device=dev)  # dealt by estimated work
    pipe.count_neighborhoods(torch.arange(lo, min(lo + 64, hi), dtype=torch.int32, device=dev))  # warm-up
    ctx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.desco_kernel_launches()
    lib.desco_profile_enable(1)
    a.record()
    kept, counts = pipe.count_neighborhoods(mine, max_centres=args.c5_chunk)
    b.record()
    ctx.barrier()
    pm, pl = (ctypes.c_double * PROF_SLOTS)(), (ctypes.c_int64 * PROF_SLOTS)()
    lib.desco_profile_read(pm, pl)
    lib.desco_profile_enable(0)
    pc_launches = lib.desco_kernel_launches() - l0
    my_ms = a.elapsed_time(b)
    pc_ms, pc_ms_min = ctx.reduce(my_ms, "max"), ctx.reduce(my_ms, "min")
    part_kernel_ms = ctx.reduce(pm[0], "max")
    shmp_kernel_ms = ctx.reduce(pm[1] + pm[2], "max")
    shmp_layer_kernel_ms = ctx.reduce(pm[1], "max")
    st = pipe.last_stats
    G, V, E = (ctx.reduce(st[k], "sum") for k in ("neighborhoods", "rows", "directed_edges"))
    max_rows = int(ctx.reduce(st["max_rows"], "max"))
    count_checksum = ctx.reduce(float(counts.double().sum().item()) if counts.numel() else 0.0, "sum")

    # ---------------- gossip over the whole graph: node-range sharded, halo exchange pipelined per query group ---------
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    x = torch.floor(torch.exp(torch.randn((N, Q), device=dev, generator=gen)))  # replicated hand-off of the counting stage
    (x,), rep_x = replicate_inputs(ctx, [x], "counts x[N, Q]")
    replicated = [rep_graph, rep_x]

    def step():
        with torch.no_grad():
            return pipe.gossip(x, qe, query_group=args.query_group)

    ms, launches, _ = ctx.timed(step, steps, warmup, flush_l2=False)
    _, _, prof = ctx.timed(step, min(steps, 3), 0, profile=True, flush_l2=False)
    psteps = min(steps, 3)
    l0_ms, gather_ms, chain_ms = (ctx.reduce(prof[0][i] / psteps, "max") for i in (3, 4, 5))
    out = step()
    from desco_b200.distributed import gossip_shard_plan

    plan = gossip_shard_plan(N, Q, world, args.query_group)

    # e2e: the same forward through the public API from pinned HOST buffers: this rank's rows of x copied H2D, the
    # hand-off all-gather of x, the sharded forward, this rank's output rows copied D2H
    nlo, nhi = plan.ranges[rank]
    h_x = x[nlo:nhi].cpu().pin_memory()
    h_out = torch.empty((nhi - nlo, Q), dtype=torch.float32).pin_memory()
    x_buf = torch.empty((plan.n_rows, Q), dtype=torch.float32, device=dev)

    def step_e2e():
        with torch.no_grad():
            x_buf[nlo:nhi].copy_(h_x, non_blocking=True)
            if world > 1:
                flat = x_buf.view(-1)
                per = flat.numel() // world
                dist.all_gather_into_tensor(flat, flat[rank * per:(rank + 1) * per])
            own = pipe.gossip(x_buf[:N], qe, query_group=args.query_group, gather_output=False)
            h_out.copy_(own, non_blocking=True)

    e2e_steps = max(3, min(steps, 5))
    ms_e2e, _, _ = ctx.timed(step_e2e, e2e_steps, 2, flush_l2=False)
    torch.cuda.synchronize()
    e2e_ok = bool(torch.equal(h_out, out[nlo:nhi].cpu()))

    # single-GPU forward of the same graph in the same job (rank 0 alone; the other ranks wait): the strong-scaling base
    single_ms, sharded_equals_single, sharded_vs_single = None, None, None
    if rank == 0:
        def single():
            with torch.no_grad():
                return gm.emb_model.forward_all_queries(g.rowptr, g.col, x, qe)
        ref_out = single()
        if world > 1:
            torch.cuda.synchronize()
            evs = []
            for _ in range(3):
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
                single()
                eb.record()
                evs.append((ea, eb))
            torch.cuda.synchronize()
            single_ms = sum(p.elapsed_time(q) for p, q in evs) / len(evs)
        sharded_equals_single = bool(torch.equal(ref_out, out))
        sharded_vs_single = float(((ref_out - out).abs() / ref_out.abs().clamp(min=1.0)).max().item())
        if not sharded_equals_single:  # say where: which rows, whose, how many queries
            dd = (ref_out - out).abs() / ref_out.abs().clamp(min=1.0)
            bad = (dd > 0).nonzero()
            rows = torch.unique(bad[:, 0])
            deg = g.rowptr[1:] - g.rowptr[:-1]
            print(f"[bench] sharded != single: {bad.shape[0]} entries in {rows.numel()} rows", file=sys.stderr)
            for r in rows[:16].tolist():
                qs = bad[bad[:, 0] == r][:, 1].tolist()
                print(f"[bench]   row {r} deg {int(deg[r])} owner {r // plan.n_loc} off {r % plan.n_loc} queries {qs[:12]} "
                      f"err {dd[r].max().item():.3e} sharded {out[r, qs[0]].item():.6f} single {ref_out[r, qs[0]].item():.6f}", file=sys.stderr)
            gm.emb_model.precision = "fp32"
            ref32 = single()
            gm.emb_model.precision = "bf16x3"
            for r in rows[:25].tolist():
                nb = g.col[int(g.rowptr[r]):int(g.rowptr[r + 1])]
                print(f"[bench]   row {r}: |sharded - fp32 path| {(out[r] - ref32[r]).abs().max().item():.3e}  |single - fp32 path| "
                      f"{(ref_out[r] - ref32[r]).abs().max().item():.3e}  neighbour degrees {sorted(deg[nb.long()].tolist())[-4:]} "
                      f"neighbours among the bad rows {int(torch.isin(nb.long(), rows).sum())}", file=sys.stderr)
            again = single()
            print(f"[bench]   single forward repeated equals itself: {bool(torch.equal(again, ref_out))}", file=sys.stderr)
        del ref_out
    ctx.barrier()

    # ---------------- parity against the CPU oracle on seeded samples (outside every timed region) ----------------
    parity = None
    if rank == 0 and not args.no_parity:
        parity = config5_parity(ctx, args, g, nm, gm, pipe, x, qe, out, depth)
        parity["sharded_forward_equals_single_gpu_forward_bitwise"] = sharded_equals_single
        parity["sharded_vs_single_gpu_forward_max_err_floor1"] = sharded_vs_single
        parity["e2e_output_equals_resident_output_bitwise"] = e2e_ok
    ctx.barrier()
    if rank != 0:
        return None
    tpeak, tsrc = _tensor_peak()
    mm_flops = 2.0 * (128 * 64 + 128 * 64 + 64 * 64 + 64 * 256) * N * Q / world  # per rank
    return {
        "workload": config5_name(args), "nodes": N, "directed_edges": M, "queries": Q, "depth": depth, "n_gpus": world,
        "scaling": "strong", "csr": "replicated on every GPU", "inputs_replicated_check": replicated,
        "partition_count": {
            "what": "canonical partition + SHMP typing + SHMP counting (29 queries) of a fixed seeded centre sample, dealt "
                    "over the ranks by estimated work (ShardedPipeline.deal_centres; a full sweep uses the contiguous "
                    "work-balanced centre ranges), no collective; int32-safe chunks",
            "centres": int(len(sample)), "chunk_centres": args.c5_chunk, "neighborhoods": int(G), "rows": int(V),
            "directed_edges": int(E), "max_rows": max_rows, "ms": pc_ms, "ms_fastest_rank": pc_ms_min,
            "partition_kernels_ms": part_kernel_ms, "shmp_kernels_ms": shmp_kernel_ms,
            "of_which_shmp_layer_kernels_ms": shmp_layer_kernel_ms, "gpu_launches_rank0": int(pc_launches),
            "centres_per_s": len(sample) / (pc_ms * 1e-3), "neighborhoods_per_s": G / (pc_ms * 1e-3),
            "count_checksum": count_checksum,
        },
        "gossip": {
            "metric": GOSSIP_METRIC, "value": N * steps / (ms * 1e-3), "unit": GOSSIP_UNIT, "ms_per_step": ms / steps,
            "steps": steps, "warmup": warmup, "gpu_launches_rank0": int(launches),
            "exchange": {"collective": "ncclAllGather (in place, all_gather_into_tensor), async on the NCCL stream" if world > 1 else "none",
                         "query_group": plan.query_group, "halo_all_gathers_per_step": len(plan.groups) if world > 1 else 0,
                         "halo_bytes_received_per_rank_per_step": plan.halo_bytes() if world > 1 else 0,
                         "output_bytes_received_per_rank_per_step": plan.output_bytes() if world > 1 else 0,
                         "inside_timed_region": True},
            "stage_ms_per_step": {"layer0_scalar_sweep": l0_ms, "layer1_gated_sweep_gather": gather_ms,
                                  "layer1_postmp_tcgen05_chain": chain_ms, "timed_in": "extra profiled steps, max over ranks"},
            "single_gpu_ms_same_job": single_ms,
            "roofline_gather": gather_roofline(N, M, Q, world, gather_ms),
            "roofline": {"kernel": "gossip_chain_kernel", "bound": "tensor", "unit": "TFLOP/s",
                         "achieved": mm_flops / (chain_ms * 1e-3) / 1e12, "peak": tpeak,
                         "frac": mm_flops / (chain_ms * 1e-3) / 1e12 / tpeak, "peak_source": tsrc, "traffic": None,
                         "note": "useful fp32-equivalent flops per rank (3x as many bf16 flops are issued) over the chain kernel's time"},
            "e2e": {"value": N * e2e_steps / (ms_e2e * 1e-3), "unit": GOSSIP_UNIT, "ms_per_step": ms_e2e / e2e_steps,
                    "h2d_bytes_per_step": int(h_x.numel() * 4) * world, "d2h_bytes_per_step": int(h_out.numel() * 4) * world,
                    "h2d_bytes_per_step_per_rank": int(h_x.numel() * 4), "d2h_bytes_per_step_per_rank": int(h_out.numel() * 4),
                    "what": "rank-local rows of the counts from pinned host memory, hand-off all-gather of x, sharded forward, "
                            "rank-local output rows back to pinned host memory"},
        },
        "parity": parity,
    }


def config5_parity(ctx, args, g, nm, gm, pipe, x, qe, out, depth):
    """Seeded samples of the 10M-node workload against the CPU oracle (oracle/large.py): canonical neighborhoods +
    SHMP types bit-exact, counts and gossip outputs within 1e-4 (floored at 1)."""
    torch = ctx.torch
    from desco_b200.data import partition_batch
    from oracle import model as M
    from oracle import partition as P
    from oracle.large import BallView, ball_sizes, gossip_closure

    rowptr, col = g.rowptr.cpu().numpy(), g.col.cpu().numpy()
    N = g.num_nodes
    rng = np.random.default_rng(5)
    cand = np.sort(rng.choice(N, size=64, replace=False))
    centres = cand[ball_sizes(rowptr, col, cand, depth) < 5000][:12]
    got = partition_batch(g, torch.as_tensor(centres, dtype=torch.int32, device=ctx.dev), depth, "hetero")
    ref = P.partition_dataset(BallView(rowptr, col, centres, depth), depth, mode="hetero", centres=centres)
    gn = got.to_numpy()
    part_ok = all(np.array_equal(gn[k], ref[k]) for k in ("nbh_ptr", "node_gid", "edge_ptr", "edge_col", "edge_tri", "centre", "indicator"))
    torch.manual_seed(0)
    om = M.NeighborhoodCountingModel().eval()
    om.load_state_dict({k: v.cpu() for k, v in nm.state_dict().items()})
    with torch.no_grad():
        want_pred = om.pre_exponent(ref, M.query_batch(), pyg_batch_size=512)
        want = 2 ** want_pred - 1
        have = nm.graph_to_count(got).cpu()
        have_pred = nm.graph_to_pred(got).cpu()
    # 2^pred magnifies a relative pre-exponent error by ln2 * |pred| (random-init weights push pred of these 10^3-row
    # neighborhoods to ~20): the bar is on the pre-exponent, and on the counts where |pred| <= 1 (there it follows from it)
    sane = want_pred.abs() <= 1.0
    pred_err = _rel_err(have_pred, want_pred) if want.numel() else 0.0
    count_err = _rel_err(have[sane], want[sane]) if bool(sane.any()) else 0.0
    # the gossip oracle runs in float64: its literal per-edge index_add accumulates a hub's 10^4 neighbour rows
    # sequentially, which in fp32 is itself off by more than the tolerance
    og = M.GossipCountingModel()
    og.emb_model.load_state_dict({k: v.cpu() for k, v in gm.emb_model.state_dict().items()})
    og = og.double()
    og.set_query_emb(qe.cpu().double())
    nodes_s = rng.choice(N, size=12, replace=False)
    nodes, ei, pos = gossip_closure(rowptr, col, nodes_s)
    with torch.no_grad():
        gref = og.graph_to_count(x[torch.as_tensor(nodes, device=ctx.dev)].cpu().double(), torch.from_numpy(ei))[torch.as_tensor(pos)]
    gerr = _rel_err(out[torch.as_tensor(nodes_s, device=ctx.dev)].cpu(), gref)
    res = {
        "oracle": "CPU restatement (oracle/) on the k-hop balls / 2-hop closure of the samples (oracle/large.py); gossip oracle in float64",
        "partition_sample_centres": int(len(centres)), "partition_and_types_bit_exact": bool(part_ok),
        "count_sample_neighborhoods": int(want.shape[0]), "pre_exponent_max_err_floor1": pred_err,
        "count_max_err_floor1_where_abs_pred_le_1": count_err,
        "gossip_sample_nodes": int(len(nodes_s)), "gossip_closure_nodes": int(len(nodes)), "gossip_max_err_floor1": gerr,
        "tolerance": TOL,
    }
    assert part_ok, "config-5 partition sample differs from the oracle"
    assert pred_err <= TOL and count_err <= TOL and gerr <= TOL, res
    return res


def run_ours(args):
    ctx = Ctx()
    torch, dev, lib, rank, world = ctx.torch, ctx.dev, ctx.lib, ctx.rank, ctx.world
    from desco_b200.data import DeviceCSR, partition_batch
    from desco_b200.lightning_model import STANDARD_QUERY_IDS, NeighborhoodCountingModel

    torch.manual_seed(0)
    model = NeighborhoodCountingModel().eval().to(dev)
    model.set_pyg_batch_size(512)
    model.set_queries(STANDARD_QUERY_IDS)
    model.get_query_emb()  # query embeddings are input-independent: computed once, like a cached set_queries

    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()

    if world > 1:
        c5 = run_config5(ctx, args, model, args.steps, args.warmup)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            gs = c5["gossip"]
            line = {
                "metric": GOSSIP_METRIC, "value": gs["value"], "unit": GOSSIP_UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": gs["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": c5["workload"], "nodes": c5["nodes"], "directed_edges": c5["directed_edges"], "queries": 29,
                           "sharding": "node ranges (gossip) / centre ranges (partition + SHMP count), CSR replicated",
                           "collectives_in_timed_region": "halo all-gathers of s4 (one per query group) + output-row all-gathers",
                           "l2": "inputs (1.16 GB of counts, 4.6 GB of halo scalars) far larger than L2",
                           "precision": "bf16x3 tcgen05 GEMM chain, fp32 accumulate (1e-4 parity path)",
                           "n1_base": "the N = 1 line of bench.py is BASELINE configs[1] (config 2); the 1-GPU time of THIS workload "
                                      "is config5.gossip.single_gpu_ms_same_job here and config5.gossip.ms_per_step of the N = 1 line"},
                "value_n1_same_workload": (c5["nodes"] / (gs["single_gpu_ms_same_job"] * 1e-3)
                                           if gs.get("single_gpu_ms_same_job") else None),
                "value_n1_note": "this workload on ONE GPU, measured in this job on rank 0 (the N = 1 bench line is config 2, another "
                                 "metric): strong-scaling efficiency = value / (n_gpus * value_n1_same_workload)",
                "e2e": gs["e2e"], "gpu_launches": gs["gpu_launches_rank0"], "clocks": clocks, "roofline": gs["roofline"],
                "cpu_baseline": None, "config5": c5, "parity": c5["parity"],
            }
            print(json.dumps(line))
        ctx.dist.destroy_process_group()
        return

    # ------------------------------------------------ N = 1: config 2 ------------------------------------------------
    csr, centres_np = build_workload(seed=0)
    # CSR + centres live in ONE device block mirrored by one pinned host block (pipeline.pack_int32_block): the e2e leg
    # refreshes all of a step's inputs with a single H2D copy
    from desco_b200.pipeline import NeighborhoodCountStep, pack_int32_block

    h_block, parts = pack_int32_block([csr.rowptr, csr.col, csr.graph_ptr, np.asarray(centres_np, dtype=np.int32)])
    d_block = h_block.to(dev)
    dv = [d_block[o:o + n] for o, n in parts]
    graph = DeviceCSR(dv[0], dv[1], dv[2], max(int(np.diff(csr.graph_ptr).max()), 1), csr)
    centres = dv[3]

    # pinned host copies for the e2e leg
    h_rowptr, h_col, h_gptr, h_centres = (h_block[o:o + n] for o, n in parts)  # views of the pinned block
    h_out = torch.empty((NUM_NBH, 29), dtype=torch.float32).pin_memory()
    h2d_bytes = (h_block.numel() * 4 if not args.eager_step else
                 sum(t.numel() * t.element_size() for t in (h_rowptr, h_col, h_gptr, h_centres)))
    d2h_bytes = h_out.numel() * 4 + 16

    # The step through the public API: desco_b200.pipeline.NeighborhoodCountStep = canonical partition + SHMP typing ->
    # fused SHMP layers -> readout -> count head, stream-ordered (no host round trip) and replayed as ONE CUDA graph.
    # (--eager-step: the same work as partition_batch + graph_to_count, one host sync inside the partition.)
    step_obj = None if args.eager_step else NeighborhoodCountStep(model, graph, centres, DEPTH, centres_storage=centres)
    h_sizes = torch.empty(4, dtype=torch.int32).pin_memory()

    def step_resident():
        with torch.no_grad():
            if step_obj is None:
                return model.graph_to_count(partition_batch(graph, centres, DEPTH, "hetero"))
            return step_obj()[0]

    def step_e2e():
        with torch.no_grad():
            if step_obj is None:
                g = DeviceCSR(h_rowptr.to(dev, non_blocking=True), h_col.to(dev, non_blocking=True),
                              h_gptr.to(dev, non_blocking=True), graph.max_graph_nodes)
                c = h_centres.to(dev, non_blocking=True)
                out = model.graph_to_count(partition_batch(g, c, DEPTH, "hetero"))
                h_out.copy_(out, non_blocking=True)
                return out
            d_block.copy_(h_block, non_blocking=True)  # the step's CSR + centres are refilled from the host every step: ONE copy
            out, sizes = step_obj()
            h_out.copy_(out, non_blocking=True)      # [C, Q]: rows >= G (sizes[0]) are padding
            h_sizes.copy_(sizes, non_blocking=True)
        return out

    # the batch shape (for the algorithmic-byte model)
    b0 = partition_batch(graph, centres, DEPTH, "hetero")
    G, V, E = b0.num_neighborhoods, b0.num_rows, b0.num_edges

    ms, launches, _ = ctx.timed(step_resident, args.steps, args.warmup)
    ms_e2e, _, _ = ctx.timed(step_e2e, args.steps, max(3, args.warmup))
    # the same K steps again with a CUDA-event pair around every kernel launch (the library's desco_profile_* hooks):
    # per-stage times and the dominant kernel's launch duration.  Kept out of the headline pass because the event
    # records themselves cost a few microseconds per launch.
    def step_eager():
        with torch.no_grad():
            return model.graph_to_count(partition_batch(graph, centres, DEPTH, "hetero"))

    # (the event hooks fire at launch time, so this pass runs the same kernels eagerly, not as a graph replay)
    ms_prof, _, prof = ctx.timed(step_eager, args.steps, 1, profile=True)
    gossip = run_gossip_leg(ctx, args, model) if not args.no_gossip else None
    c5 = run_config5(ctx, args, model, max(3, min(args.steps, 5)), 3) if not args.no_config5 else None
    clocks = sampler.stop()

    value = NUM_NBH * args.steps / (ms * 1e-3)
    e2e_value = NUM_NBH * args.steps / (ms_e2e * 1e-3)
    peak, peak_src = _peaks()
    layer_ms, layer_launches = prof[0][1], prof[1][1]
    # one fused launch runs all 8 layers; per layer the reference formulation gathers E rows, reads V self rows and
    # writes V rows (fp32 x 64): B_shmp = L * 4F * (E + 2V)   
    alg_bytes = 8 * 4 * 64 * (E + 2 * V)
    achieved = alg_bytes / (layer_ms / max(layer_launches, 1) * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get("shmp_fused_kernel_dram_bytes_per_launch")
        traffic_src = tj.get("source")
    # bounded CPU baseline (oracle port) on this box's host cores: 3 warm-up passes, then timed passes; the last pass
    # doubles as the in-run parity check of the exact bench workload (outside every timed region)
    om, qb = cpu_models()
    for w in range(3):
        cpu_step(csr, centres_np[w * 128:(w + 1) * 128], om, qb)
    sample, passes = 1024, 3
    t0 = time.perf_counter()
    for p in range(passes):
        ref_b, ref_counts = cpu_step(csr, centres_np[p * sample:(p + 1) * sample], om, qb)
    cpu_dt = time.perf_counter() - t0
    om.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    ref_b, ref_counts = cpu_step(csr, centres_np[(passes - 1) * sample:passes * sample], om, qb)
    with torch.no_grad():
        got_b = partition_batch(graph, centres[(passes - 1) * sample:passes * sample], DEPTH, "hetero")
        got_counts = model.graph_to_count(got_b).cpu()
    gn = got_b.to_numpy()
    part_ok = all(np.array_equal(gn[k], ref_b[k]) for k in ("nbh_ptr", "node_gid", "edge_ptr", "edge_col", "edge_tri", "centre"))
    count_err = _rel_err(got_counts, ref_counts)
    assert part_ok and count_err <= TOL, ("in-run parity failed", part_ok, count_err)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "depth": DEPTH, "queries": 29, "neighborhoods_per_gpu": G, "rows": V,
                   "directed_edges": E, "pyg_batch_size": 512, "l2": "flushed between steps (256 MiB write)",
                   "step": "partition_batch + graph_to_count, one host sync per step" if args.eager_step else
                           "pipeline.NeighborhoodCountStep: stream-ordered, one CUDA-graph replay per step",
                   "precision": "bf16x3 tcgen05 layers + bf16x6 readout, fp32 accumulate (1e-4 parity path)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches) if step_obj is None else int(step_obj.launches_per_step * args.steps),
        "gpu_launches_note": None if step_obj is None else f"{step_obj.launches_per_step} kernel nodes per CUDA-graph replay x {args.steps} steps",
        "clocks": clocks,
        "roofline": {"kernel": "shmp_fused_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": layer_ms / max(layer_launches, 1),
                     "launches_timed": int(layer_launches),
                     "timed_in": "second pass of the same K steps, CUDA events around every launch on the launching stream",
                     "note": "algorithmic bytes are the reference formulation's (SURVEY 8d); the kernel keeps the features on "
                             "chip for all 8 layers, so `traffic` (DRAM bytes per launch, ncu) is far below them and the "
                             "fraction measures reference-formulation bytes per second, not achieved HBM bandwidth"},
        "stage_ms_per_step": {"partition": prof[0][0] / args.steps, "shmp_layers": prof[0][1] / args.steps,
                              "shmp_other": prof[0][2] / args.steps, "step_with_profiling_events": ms_prof / args.steps},
        "parity": {"workload": f"neighborhoods [{(passes - 1) * sample}, {passes * sample}) of the bench batch, pyg_batch_size 512",
                   "oracle": "CPU restatement (oracle/), same weights", "partition_and_types_bit_exact": bool(part_ok),
                   "count_max_err_floor1": count_err, "tolerance": TOL,
                   "count_tolerance_definition": "|d| <= 1e-4 * max(1, |ref|) on 2^pred - 1 (random-init counts are ~ -0.03, so a "
                                                 "pure relative test is ill-conditioned; tests/test_shmp_gpu.py also bounds the pre-exponent)"},
        "gossip": gossip,
        "config5": c5,
        "cpu_baseline": {"value": sample * passes / cpu_dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{passes} passes over {sample} of the {NUM_NBH} neighborhoods each after 3 warm-up passes, "
                                   "networkx partition single-process + torch CPU forward on all cores"},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--eager-step", action="store_true", help="N = 1: partition_batch + graph_to_count instead of the CUDA-graph step")
    ap.add_argument("--no-gossip", action="store_true", help="N = 1: skip the 1M-node gossip leg")
    ap.add_argument("--no-config5", action="store_true", help="N = 1: skip the 1-GPU base of the config-5 workload")
    ap.add_argument("--no-parity", action="store_true", help="config 5: skip the CPU-oracle sample checks")
    ap.add_argument("--gossip-nodes", type=int, default=1_000_000)
    ap.add_argument("--gossip-edges", type=int, default=10_000_000)
    ap.add_argument("--c5-nodes", type=int, default=10_000_000)
    ap.add_argument("--c5-edges", type=int, default=100_000_000)
    ap.add_argument("--c5-centres", type=int, default=16384, help="config 5: size of the seeded centre sample (whole job)")
    ap.add_argument("--c5-chunk", type=int, default=2048, help="config 5: centres per packed batch")
    ap.add_argument("--query-group", type=int, default=4, help="queries per halo all-gather of the sharded gossip forward")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
