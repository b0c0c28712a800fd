"""Where does the host time of one step go?  (wall clock of each Python-level call WITHOUT a trailing sync = pure
launch/bookkeeping overhead, then the same with a sync = overhead + device time)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from bench import build_workload, DEPTH
from desco_b200.data import DeviceCSR, partition_batch
from desco_b200.lightning_model import STANDARD_QUERY_IDS, NeighborhoodCountingModel

csr, cen = build_workload(0)
torch.manual_seed(0)
m = NeighborhoodCountingModel().eval().cuda(); m.set_pyg_batch_size(512); m.set_queries(STANDARD_QUERY_IDS); m.get_query_emb()
g = DeviceCSR.from_host(csr); c = torch.as_tensor(cen, dtype=torch.int32, device="cuda")

def timeit(fn, n=50, sync=False):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        if sync: torch.cuda.synchronize()
    if not sync: t1 = time.perf_counter(); torch.cuda.synchronize()
    else: t1 = time.perf_counter()
    return (t1 - t0) / n * 1e6

with torch.no_grad():
    b = partition_batch(g, c, DEPTH)
    print("partition_batch (has its own sync)      us:", timeit(lambda: partition_batch(g, c, DEPTH)))
    print("emb_model(b) launch only                us:", timeit(lambda: m.emb_model(b)))
    print("emb_model(b) + sync                     us:", timeit(lambda: m.emb_model(b), sync=True))
    e = m.emb_model(b); q = m.get_query_emb()
    print("get_query_emb (cached)                  us:", timeit(lambda: m.get_query_emb()))
    print("embed_to_count launch only              us:", timeit(lambda: m.embed_to_count((e, q))))
    print("embed_to_count + sync                   us:", timeit(lambda: m.embed_to_count((e, q)), sync=True))
    print("full step + sync                        us:", timeit(lambda: m.graph_to_count(partition_batch(g, c, DEPTH)), sync=True))
    print("packed_weights() check                  us:", timeit(lambda: m.emb_model.packed_weights()))
