"""Per-phase cycles of the tensor-core gossip layer-1 kernel on the bench's power-law target (1M nodes, 29 queries)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from desco_b200 import _lib
from desco_b200.data import gen_powerlaw_device
from desco_b200.lightning_model import STANDARD_QUERY_IDS, GossipCountingModel, NeighborhoodCountingModel

lib = _lib.load()
dev = torch.device("cuda", 0)
g = gen_powerlaw_device(1_000_000, 10_000_000, seed=0, device=dev)
torch.manual_seed(0)
nm = NeighborhoodCountingModel().eval().to(dev); nm.set_queries(STANDARD_QUERY_IDS); qe = nm.get_query_emb()
gm = GossipCountingModel().eval().to(dev); gm.set_query_emb(qe)
x = torch.floor(torch.exp(torch.randn((g.num_nodes, qe.shape[0]), device=dev)))
from types import SimpleNamespace
with torch.no_grad():
    gm.graph_to_count(SimpleNamespace(graph=g, x=x))
torch.cuda.synchronize()
out = (ctypes.c_uint64 * 6)()
lib.desco_gossip_tc_phase_cycles(out, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
with torch.no_grad():
    gm.graph_to_count(SimpleNamespace(graph=g, x=x))
b.record(); torch.cuda.synchronize()
lib.desco_gossip_tc_phase_cycles(out, 1)
tot = float(sum(out))
print("forward ms", a.elapsed_time(b), "cycles per CTA", tot / 148)
for n, v in zip(("load", "x2", "y1", "y2", "y4", "spare"), out):
    print(f"{n:8s} {v / 148:12.0f} cyc/CTA  {v / tot:.3f}")
