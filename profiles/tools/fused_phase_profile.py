import os; os.environ["DESCO_FUSED_PHASE_TIMING"] = "1"  # launch the instantiation with the clock64 phase counters
import sys, ctypes, torch, numpy as np
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from bench import build_workload, DEPTH
from desco_b200 import _lib
from desco_b200.data import DeviceCSR, partition_batch
from desco_b200.lightning_model import STANDARD_QUERY_IDS, NeighborhoodCountingModel
lib=_lib.load()
csr, cen = build_workload(0)
torch.manual_seed(0)
m = NeighborhoodCountingModel().eval().cuda(); m.set_pyg_batch_size(512); m.set_queries(STANDARD_QUERY_IDS); m.get_query_emb()
g = DeviceCSR.from_host(csr); c = torch.as_tensor(cen, dtype=torch.int32, device='cuda')
b = partition_batch(g, c, DEPTH)
for _ in range(3): m.graph_to_count(b)
torch.cuda.synchronize()
out=(ctypes.c_uint64*10)()
lib.desco_shmp_fused_phase_cycles(out,1)
N=10
for _ in range(N): m.emb_model(b)
torch.cuda.synchronize()
lib.desco_shmp_fused_phase_cycles(out,1)
names=['setup','poolA(+barrier)','issue(warp0)','canon','wait_mma','t2s','gather','poolB(+barrier)','issuer: weight wait','issuer: mma issue']
tot=sum(out[:8])
print("cycles per CTA per launch:", tot/148/N, "= us @1.965GHz", tot/148/N/1965)
for n,v in zip(names,out): print(f"{n:9s} {v/148/N:10.0f} cyc/CTA/launch  {v/tot:.3f}")
