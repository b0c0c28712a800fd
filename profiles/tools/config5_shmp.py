"""Config-5 SHMP counting in isolation: one chunk of seeded centres of the 10M-node power-law target through
partition_batch + NeighborhoodCountingModel.graph_to_count (multi-tile tcgen05 path), timed with CUDA events.
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--edges", type=int, default=100_000_000)
    ap.add_argument("--centres", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    from desco_b200.data import gen_powerlaw_device, partition_batch
    from desco_b200.lightning_model import STANDARD_QUERY_IDS, NeighborhoodCountingModel

    dev = torch.device("cuda", 0)
    g = gen_powerlaw_device(args.nodes, args.edges, seed=0, device=dev)
    torch.manual_seed(0)
    nm = NeighborhoodCountingModel().eval().to(dev)
    nm.set_queries(STANDARD_QUERY_IDS)
    nm.get_query_emb()
    rng = np.random.default_rng(11)
    centres = torch.as_tensor(np.sort(rng.choice(g.num_nodes, size=args.centres, replace=False)), dtype=torch.int32, device=dev)
    batch = partition_batch(g, centres, 2, "hetero")
    with torch.no_grad():
        nm.graph_to_count(batch)
    torch.cuda.synchronize()
    evs = []
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        with torch.no_grad():
            nm.graph_to_count(batch)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    V, E, G = batch.num_rows, batch.num_edges, batch.num_neighborhoods
    alg = 8 * 256 * (E + 2 * V)
    print(json.dumps({"neighborhoods": G, "rows": V, "directed_edges": E, "max_rows": batch.max_rows, "shmp_count_ms": ms,
                      "algorithmic_bytes": alg, "algorithmic_gbs": alg / (min(ms) * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
