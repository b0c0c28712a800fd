"""Gossip forward on the bench's power-law target (default 1M nodes / 10M undirected edges, 29 queries): per-kernel-group
device time through the library's desco_profile_* hooks, and the tensor-core path against the library's own fp32 FFMA path."""
import argparse, ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from types import SimpleNamespace
from desco_b200 import _lib
from desco_b200.data import gen_powerlaw_device
from desco_b200.lightning_model import STANDARD_QUERY_IDS, GossipCountingModel, NeighborhoodCountingModel

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=1_000_000)
ap.add_argument("--edges", type=int, default=10_000_000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-fp32", action="store_true")
args = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda", 0)
g = gen_powerlaw_device(args.nodes, args.edges, seed=0, device=dev)
torch.manual_seed(0)
nm = NeighborhoodCountingModel().eval().to(dev); nm.set_queries(STANDARD_QUERY_IDS); qe = nm.get_query_emb()
gm = GossipCountingModel().eval().to(dev); gm.set_query_emb(qe)
x = torch.floor(torch.exp(torch.randn((g.num_nodes, qe.shape[0]), device=dev)))
batch = SimpleNamespace(graph=g, x=x)
with torch.no_grad():
    for _ in range(3):
        out = gm.graph_to_count(batch)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
with torch.no_grad():
    for _ in range(args.steps):
        out = gm.graph_to_count(batch)
b.record(); torch.cuda.synchronize()
fwd = a.elapsed_time(b) / args.steps
lib.desco_profile_enable(1)
with torch.no_grad():
    for _ in range(args.steps):
        gm.graph_to_count(batch)
pm, pl = (ctypes.c_double * 6)(), (ctypes.c_int64 * 6)()
lib.desco_profile_read(pm, pl)
lib.desco_profile_enable(0)
rec = {"nodes": g.num_nodes, "directed_edges": int(g.col.numel()), "queries": int(qe.shape[0]), "forward_ms": fwd,
       "layer0_ms": pm[3] / args.steps, "gather_ms": pm[4] / args.steps, "chain_ms": pm[5] / args.steps}
if not args.no_fp32:
    gm.emb_model.precision = "fp32"
    with torch.no_grad():
        ref = gm.graph_to_count(batch)
    gm.emb_model.precision = "bf16x3"
    d = (out - ref).abs() / ref.abs().clamp(min=1.0)
    rec["max_err_floor1_vs_fp32_ffma_path"] = float(d.max())
print(json.dumps(rec))
