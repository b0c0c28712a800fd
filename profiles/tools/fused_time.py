"""Average launch time of the fused SHMP layer kernel on the bench batch (CUDA events around emb_model, minus nothing:
the readout launches are included; compare variants, not absolutes).  DESCO_FUSED_SCHED selects the fragment schedule."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from bench import build_workload, DEPTH
from desco_b200 import _lib
from desco_b200.data import DeviceCSR, partition_batch
from desco_b200.lightning_model import STANDARD_QUERY_IDS, NeighborhoodCountingModel
import ctypes
lib = _lib.load()
csr, cen = build_workload(0)
torch.manual_seed(0)
m = NeighborhoodCountingModel().eval().cuda(); m.set_pyg_batch_size(512); m.set_queries(STANDARD_QUERY_IDS); m.get_query_emb()
g = DeviceCSR.from_host(csr); c = torch.as_tensor(cen, dtype=torch.int32, device='cuda')
b = partition_batch(g, c, DEPTH)
for _ in range(5): m.emb_model(b)
torch.cuda.synchronize()
lib.desco_profile_enable(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ts = []
for _ in range(20):
    flush.fill_(1)
    m.emb_model(b)
torch.cuda.synchronize()
ms = (ctypes.c_double * 16)(); n = (ctypes.c_int64 * 16)()
lib.desco_profile_read(ms, n)
print("sched", os.environ.get("DESCO_FUSED_SCHED", "0"), "fused layer kernel ms/launch", ms[1] / max(n[1], 1), "launches", n[1], "other", ms[2] / 20)
