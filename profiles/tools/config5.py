"""Config 5 (BASELINE.json configs[4]): 10M-node / 100M-undirected-edge power-law target, sharded over the GPUs of one
box.  Run as  python profiles/tools/config5.py  or under torchrun (one rank per GPU).

  * canonical partition + SHMP typing, depth 2 (depth 4 is not materialisable on this graph), on
    `--chunks` chunks of `--chunk` consecutive centres per rank (centre ranges are interleaved over the ranks; no
    collective), then SHMP counting of the same chunks;
  * gossip over the WHOLE graph for all 29 queries, node-range sharded, with the all-gather of the layer-0 scalars
    (the halo) and of the result rows inside the timed region.
All times: CUDA events, max over ranks.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--edges", type=int, default=100_000_000)
    ap.add_argument("--depth", type=int, default=2)
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--gossip-steps", type=int, default=3)
    ap.add_argument("--no-shmp", action="store_true")
    args = ap.parse_args()

    from desco_b200.data import gen_powerlaw_device, partition_batch
    from desco_b200.distributed import ShardedPipeline
    from desco_b200.lightning_model import STANDARD_QUERY_IDS, GossipCountingModel, NeighborhoodCountingModel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    g = gen_powerlaw_device(args.nodes, args.edges, seed=0, device=dev)  # same seed: the CSR is replicated
    N, M = g.num_nodes, g.num_directed_edges
    deg = (g.rowptr[1:] - g.rowptr[:-1]).to(torch.int64)
    torch.manual_seed(0)
    nm = NeighborhoodCountingModel().eval().to(dev)
    nm.set_queries(STANDARD_QUERY_IDS)
    qe = nm.get_query_emb()
    torch.manual_seed(1)
    gm = GossipCountingModel().eval().to(dev)
    gm.set_query_emb(qe)

    def tmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def tsum(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- partition (+ SHMP counting) on interleaved centre chunks ----------------
    rng = np.random.default_rng(11)
    starts = rng.integers(0, N - args.chunk, size=args.chunks * world)
    mine = starts[rank::world]
    part_ms = shmp_ms = types_ms = 0.0
    G = V = E = X = P = 0
    tiers = np.zeros(2, dtype=np.int64)
    max_rows = 0
    warm = torch.arange(0, min(args.chunk, 256), dtype=torch.int32, device=dev)
    b = partition_batch(g, warm, args.depth, "hetero")
    if not args.no_shmp and b.num_neighborhoods:
        with torch.no_grad():
            nm.graph_to_count(b)
    barrier()
    for s0 in mine:
        centres = torch.arange(int(s0), int(s0) + args.chunk, dtype=torch.int32, device=dev)
        a, bb, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        batch = partition_batch(g, centres, args.depth, "hetero")
        bb.record()
        if not args.no_shmp and batch.num_neighborhoods:
            with torch.no_grad():
                counts = nm.graph_to_count(batch)
        c.record()
        from desco_b200.data import shmp_edge_types
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        if batch.num_rows:
            shmp_edge_types(batch.edge_ptr, batch.edge_col)  # the typing pass alone, to report its share of `ms`
        d1.record()
        torch.cuda.synchronize()
        types_ms += d0.elapsed_time(d1)
        part_ms += a.elapsed_time(bb)
        shmp_ms += bb.elapsed_time(c)
        G += batch.num_neighborhoods
        V += batch.num_rows
        E += batch.num_edges
        max_rows = max(max_rows, batch.max_rows)
        if "tier" in batch._cache:
            tiers += np.bincount(batch._cache["tier"].cpu().numpy(), minlength=2)[:2]
        # reference-formulation traffic of the k-hop BFS at depth 2: P = nodes expanded (distance <= 1),
        # X = adjacency entries of those nodes
        cl = centres.long()
        P += int((1 + deg[cl]).sum())
        lo, hi = g.rowptr[cl].long(), g.rowptr[cl + 1].long()
        idx = torch.repeat_interleave(lo, (hi - lo)) + (torch.arange(int((hi - lo).sum()), device=dev)
                                                         - torch.repeat_interleave(torch.cumsum(hi - lo, 0) - (hi - lo), hi - lo))
        X += int(deg[cl].sum() + deg[g.col[idx].long()].sum())
    barrier()
    import ctypes
    from desco_b200 import _lib
    ph = (ctypes.c_uint64 * 6)()
    _lib.load().desco_partition_large_phase_cycles(ph, 1)
    ph_tot = float(sum(ph)) or 1.0
    phases = {k: round(ph[i] / ph_tot, 4) for i, k in enumerate(("expand", "component", "sort", "degree", "emit", "wipe"))}
    part_ms, shmp_ms, types_ms = tmax(part_ms), tmax(shmp_ms), tmax(types_ms)
    G, V, E, X, P = (tsum(v) for v in (G, V, E, X, P))
    tiers = [tsum(t) for t in tiers]
    max_rows = int(tmax(max_rows))

    # ---------------- gossip over the whole graph, node-range sharded ----------------
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    Q = qe.shape[0]
    x = torch.floor(torch.exp(torch.randn((N, Q), device=dev, generator=gen)))
    pipe = ShardedPipeline(g, nm, gm, None, depth=args.depth)
    with torch.no_grad():
        out = pipe.gossip(x, qe)  # warm-up
    barrier()
    evs = []
    for _ in range(args.gossip_steps):
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        with torch.no_grad():
            out = pipe.gossip(x, qe)
        b2.record()
        evs.append((a, b2))
    barrier()
    gossip_ms = tmax(sum(a.elapsed_time(b2) for a, b2 in evs) / len(evs))
    checksum = float(out.double().sum().item())

    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        part_bytes = 4 * X + 8 * P + 4 * V + 9 * E
        gos_bytes = Q * (512 * M + 1280 * N) + 8 * Q * N + 8 * M
        print(json.dumps({
            "workload": f"powerlaw_chunglu_{N}nodes_{M // 2}undirected_edges", "n_gpus": world, "depth": args.depth,
            "partition": {
                "centres": args.chunk * args.chunks * world, "neighborhoods": G, "rows": V, "directed_edges": E,
                "max_rows": max_rows, "tier_counts_sharedhash_teambitmap": tiers, "sharedhash_tier_phase_share": phases,
                "ms": part_ms, "of_which_shmp_typing_ms": types_ms, "centres_per_s": args.chunk * args.chunks * world / (part_ms * 1e-3),
                "neighborhoods_per_s": G / (part_ms * 1e-3),
                "algorithmic_bytes": part_bytes, "achieved_gbs": part_bytes / (part_ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": part_bytes / (part_ms * 1e-3) / 1e9 / (hbm * world),
            },
            "shmp_count": None if args.no_shmp else {"ms": shmp_ms, "neighborhoods_per_s": G / (shmp_ms * 1e-3)},
            "gossip": {
                "nodes": N, "directed_edges": M, "queries": Q, "ms_per_forward": gossip_ms,
                "target_nodes_per_s": N / (gossip_ms * 1e-3), "algorithmic_bytes": gos_bytes,
                "achieved_gbs": gos_bytes / (gossip_ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": gos_bytes / (gossip_ms * 1e-3) / 1e9 / (hbm * world),
                "exchange": "all-gather of s4[N,Q,4] (halo) and of out[N,Q] inside the timed region" if world > 1 else "none",
                "checksum": checksum,
            },
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
