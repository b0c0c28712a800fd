"""Sharded gossip forward only (config 5 target), under torchrun: step time (CUDA events, max over ranks) and the
per-kernel-group times for one setting of the exchange (NCCL_* environment, --query-group).  Experiment tool."""
import argparse, ctypes, json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--edges", type=int, default=100_000_000)
ap.add_argument("--query-group", type=int, default=4)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--tag", default="")
args = ap.parse_args()
rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from desco_b200 import _lib
from desco_b200.data import gen_powerlaw_device
from desco_b200.distributed import ShardedPipeline
from desco_b200.lightning_model import STANDARD_QUERY_IDS, GossipCountingModel, NeighborhoodCountingModel

lib = _lib.load()
g = gen_powerlaw_device(args.nodes, args.edges, seed=0, device=dev)
torch.manual_seed(0)
nm = NeighborhoodCountingModel().eval().to(dev); nm.set_queries(STANDARD_QUERY_IDS); qe = nm.get_query_emb()
torch.manual_seed(1)
gm = GossipCountingModel().eval().to(dev); gm.set_query_emb(qe)
pipe = ShardedPipeline(g, nm, gm, None, depth=2)
gen = torch.Generator(device=dev); gen.manual_seed(7)
x = torch.floor(torch.exp(torch.randn((g.num_nodes, qe.shape[0]), device=dev, generator=gen)))


def step():
    with torch.no_grad():
        return pipe.gossip(x, qe, query_group=args.query_group)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    step()
barrier()
evs = []
for _ in range(args.steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record(); evs.append((a, b))
barrier()
ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
lib.desco_profile_enable(1)
for _ in range(3):
    step()
barrier()
pm, pl = (ctypes.c_double * 6)(), (ctypes.c_int64 * 6)()
lib.desco_profile_read(pm, pl)
lib.desco_profile_enable(0)
t = torch.tensor([ms, pm[3] / 3, pm[4] / 3, pm[5] / 3], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"tag": args.tag, "n_gpus": world, "query_group": args.query_group, "ms_per_step": t[0].item(),
                      "layer0_ms": t[1].item(), "gather_ms": t[2].item(), "chain_ms": t[3].item(),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}))
if world > 1:
    dist.destroy_process_group()
