import sys, os
sys.path.insert(0, '/root/repo')
import torch, numpy as np
from types import SimpleNamespace
from desco_b200.data import gen_powerlaw_device
from desco_b200.distributed import LocalComm
from desco_b200.gnn_model import GossipShardedRun
from desco_b200.lightning_model import STANDARD_QUERY_IDS, GossipCountingModel, NeighborhoodCountingModel
dev = torch.device('cuda', 0)
N, E = int(sys.argv[1]), int(sys.argv[2])
W = int(sys.argv[3])
g = gen_powerlaw_device(N, E, seed=0, device=dev)
torch.manual_seed(0)
nm = NeighborhoodCountingModel().eval().to(dev); nm.set_queries(STANDARD_QUERY_IDS); qe = nm.get_query_emb()
torch.manual_seed(1)
gm = GossipCountingModel().eval().to(dev); gm.set_query_emb(qe)
gen = torch.Generator(device=dev); gen.manual_seed(7)
x = torch.floor(torch.exp(torch.randn((g.num_nodes, qe.shape[0]), device=dev, generator=gen)))
with torch.no_grad():
    single = gm.emb_model.forward_all_queries(g.rowptr, g.col, x, qe)
    comm = LocalComm(W)
    runs = [GossipShardedRun(gm.emb_model, g.rowptr, g.col, x, qe, comm.for_rank(r), query_group=4, gather_output=False).start() for r in range(W)]
    for r in runs: r.finish()
    out = torch.cat([r.result() for r in runs], dim=0)[:N]
torch.cuda.synchronize()
d = (out - single).abs() / single.abs().clamp(min=1.0)
bad = (d > 0).nonzero()
print('mismatch entries', bad.shape[0], 'max', d.max().item())
deg = (g.rowptr[1:] - g.rowptr[:-1])
n_loc = runs[0].plan.n_loc
rows = torch.unique(bad[:, 0])
print('rows', rows.numel(), 'n_loc', n_loc)
for r in rows[:20].tolist():
    qs = bad[bad[:, 0] == r][:, 1].tolist()
    print(r, 'deg', int(deg[r]), 'rank', r // n_loc, 'off in rank', r % n_loc, 'tile row', r % 128, 'queries', qs[:8], 'err', d[r].max().item())
