"""Config 5 (10M-node / 100M-undirected-edge power-law target): the streaming count-only partition (data.partition_sizes)
on seeded centre samples - depth 2 on a large sample, depths 3 and 4 on small ones - with the sizes extrapolated to a full
sweep.  Says what SURVEY 8(d).5's "depth-2 full sweep + depth-4 on 4096 centres" would emit, and what it would cost."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np
import torch
from desco_b200.data import gen_powerlaw_device, partition_sizes

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--edges", type=int, default=100_000_000)
ap.add_argument("--d2", type=int, default=65536)
ap.add_argument("--d3", type=int, default=512)
ap.add_argument("--d4", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = gen_powerlaw_device(a.nodes, a.edges, seed=0, device=dev)
rng = np.random.default_rng(5)
out = {"workload": f"powerlaw_chunglu_{a.nodes}nodes_{a.edges}undirected_edges", "directed_edges": int(g.col.numel())}
for depth, n in ((2, a.d2), (3, a.d3), (4, a.d4)):
    cen = torch.as_tensor(np.sort(rng.choice(a.nodes, size=n, replace=False)).astype(np.int32), device=dev)
    partition_sizes(g, cen[:256], depth)  # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nv, ne = partition_sizes(g, cen, depth, max_centres=2048 if depth == 2 else 64)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rows, edges = int(nv.sum(dtype=torch.int64)), int(ne.sum(dtype=torch.int64))
    scale = a.nodes / n
    out[f"depth{depth}"] = {
        "centres": n, "count_pass_ms": ms, "centres_per_s": n / ms * 1e3, "rows": rows, "directed_edges": edges,
        "rows_per_centre_mean": rows / n, "rows_per_centre_max": int(nv.max()), "kept": int((nv > 0).sum()),
        "full_sweep_extrapolated": {"rows": rows * scale, "directed_edges": edges * scale,
                                    "packed_bytes_4V_plus_9E": (4 * rows + 9 * edges) * scale,
                                    "count_pass_seconds_one_gpu": ms * scale / 1e3}}
print(json.dumps(out))
