"""Debug tool (torchrun): sharded gossip forward over NCCL vs the single-GPU forward on rank 0; lists the rows that differ."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
from desco_b200.data import gen_powerlaw_device
from desco_b200.distributed import ShardedPipeline
from desco_b200.lightning_model import STANDARD_QUERY_IDS, GossipCountingModel, NeighborhoodCountingModel
N, E = int(sys.argv[1]), int(sys.argv[2])
g = gen_powerlaw_device(N, E, seed=0, device=dev)
torch.manual_seed(0)
nm = NeighborhoodCountingModel().eval().to(dev); nm.set_queries(STANDARD_QUERY_IDS); qe = nm.get_query_emb()
torch.manual_seed(1)
gm = GossipCountingModel().eval().to(dev); gm.set_query_emb(qe)
pipe = ShardedPipeline(g, nm, gm, None, depth=2)
gen = torch.Generator(device=dev); gen.manual_seed(7)
x = torch.floor(torch.exp(torch.randn((g.num_nodes, qe.shape[0]), device=dev, generator=gen)))
# are the inputs really replicated?
chk = torch.stack([g.col.double().sum(), g.rowptr.double().sum(), x.double().sum(), qe.double().sum()])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
from desco_b200.distributed import gossip_shard_plan
plan = gossip_shard_plan(g.num_nodes, qe.shape[0], world, 4)
lo, hi = plan.ranges[rank]
deg = g.rowptr[1:] - g.rowptr[:-1]
with torch.no_grad():
    single = gm.emb_model.forward_all_queries(g.rowptr, g.col, x, qe)
    torch.cuda.synchronize()
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    for it in range(reps):
        out = pipe.gossip(x, qe, query_group=4)
        if it % 3 == 2:  # back-to-back forwards without a host sync in between, like the bench loop
            out = pipe.gossip(x, qe, query_group=4)
            out = pipe.gossip(x, qe, query_group=4)
        eq = torch.equal(out, single)
        flag = torch.tensor([0 if eq else 1], device=dev)
        dist.all_reduce(flag)
        if flag.item() and rank == 0:
            d = (out - single).abs() / single.abs().clamp(min=1.0)
            bad = (d > 0).nonzero()
            rows = torch.unique(bad[:, 0])
            print(f"iter {it}: {bad.shape[0]} entries in {rows.numel()} rows differ, max {d.max().item():.3e}", flush=True)
            common = None
            for r in rows.tolist():
                nb = g.col[int(g.rowptr[r]):int(g.rowptr[r + 1])].long()
                inbad = int(torch.isin(nb, rows).sum())
                nq = int((bad[:, 0] == r).sum())
                print(f"   row {r} deg {int(deg[r])} owner {r // plan.n_loc} queries differing {nq} err {d[r].max().item():.3e} "
                      f"neighbours among bad rows {inbad}", flush=True)
                sset = set(nb.tolist()) | {r}
                common = sset if common is None else (common & sset)
            print("   nodes that are in every bad row's closed neighbourhood:", sorted(common)[:10],
                  [int(deg[c]) for c in sorted(common)[:10]], flush=True)
        elif rank == 0:
            print(f"iter {it}: equal", flush=True)
dist.barrier()
dist.destroy_process_group()
