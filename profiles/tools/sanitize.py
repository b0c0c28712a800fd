"""Small end-to-end invocations of every hand-written kernel family, for compute-sanitizer (memcheck / racecheck /
synccheck / initcheck):  compute-sanitizer --tool <tool> python profiles/tools/sanitize.py [which]
which: partition | shmp | gossip | train | truth | all (default).  Sizes are tiny: the sanitizer slows kernels 10-100x."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    from types import SimpleNamespace

    from desco_b200 import _lib
    from desco_b200.data import DeviceCSR, partition_batch
    from desco_b200.graph import gen_enzymes_shaped, gen_mutag_shaped, gen_powerlaw, gen_syn1827_shaped
    from desco_b200.lightning_model import STANDARD_QUERY_IDS, GossipCountingModel, NeighborhoodCountingModel

    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    nm = NeighborhoodCountingModel().eval().to(dev)
    nm.set_queries(STANDARD_QUERY_IDS[:6])
    done = []
    if which in ("partition", "all"):
        g = DeviceCSR.from_host(gen_mutag_shaped(seed=0, num_graphs=12))
        partition_batch(g, None, 4, "hetero")                      # partition_small_kernel + scan
        partition_batch(g, None, 3, "canonical")
        g2 = DeviceCSR.from_host(gen_syn1827_shaped(seed=0, stride=400))
        partition_batch(g2, None, 2, "hetero")                     # partition_kernel (bitsets in shared memory)
        pl = DeviceCSR.from_host(gen_powerlaw(3000, 12000, seed=1))
        lib.desco_partition_large_set_caps(6, 16, 16, 0, 0, 0)      # tiny shared-memory tier: most centres go to the team tier
        c = torch.arange(2900, 3000, dtype=torch.int32, device=dev)
        partition_batch(pl, c, 2, "hetero", large=True)            # partition_sparse_kernel + partition_team_kernel + typing
        lib.desco_partition_large_set_caps(13, 5120, 4096, 0, 0, 0)
        partition_batch(pl, c, 2, "hetero", large=True)
        done.append("partition")
    if which in ("shmp", "all"):
        g = DeviceCSR.from_host(gen_enzymes_shaped(seed=0, num_graphs=6))
        b = partition_batch(g, None, 4, "hetero")
        with torch.no_grad():
            nm.graph_to_count(b)                                   # shmp_fused_kernel + dense_tc + readout + head
            nm.emb_model.force_multi_tile = True
            nm.graph_to_count(b)                                   # shmp_mt_* kernels
            nm.emb_model.force_multi_tile = False
            nm.set_precision("fp32")
            nm.graph_to_count(b)                                   # FFMA layer kernels
            nm.set_precision("bf16x3")
        done.append("shmp")
    if which in ("gossip", "all"):
        csr = gen_powerlaw(2000, 9000, seed=2)
        g = DeviceCSR.from_host(csr)
        torch.manual_seed(1)
        gm = GossipCountingModel().eval().to(dev)
        qe = nm.get_query_emb()
        gm.set_query_emb(qe)
        x = torch.floor(torch.exp(torch.randn((csr.num_nodes, qe.shape[0]), device=dev)))
        with torch.no_grad():
            gm.graph_to_count(SimpleNamespace(graph=g, x=x))       # layer0 + gather + tcgen05 chain
            gm.emb_model.precision = "fp32"
            gm.graph_to_count(SimpleNamespace(graph=g, x=x))       # FFMA layer-1 kernel
            gm.emb_model.precision = "bf16x3"
            from desco_b200.distributed import LocalComm
            from desco_b200.gnn_model import GossipShardedRun

            comm = LocalComm(2)
            runs = [GossipShardedRun(gm.emb_model, g.rowptr, g.col, x, qe, comm.for_rank(r), query_group=4).start() for r in range(2)]
            for r in runs:
                r.finish()
            for r in runs:
                r.result()
        done.append("gossip")
    if which in ("train", "all"):
        from desco_b200.data import NeighborhoodBatch

        g = DeviceCSR.from_host(gen_mutag_shaped(seed=3, num_graphs=6))
        b = partition_batch(g, None, 4, "hetero")
        b.y = torch.floor(torch.exp(torch.randn((b.num_neighborhoods, 6), device=dev)))
        nm.train()
        opt = nm.configure_optimizers()["optimizer"]
        opt.zero_grad()
        nm.training_step(b, 0).backward()
        opt.step()
        nm.eval()
        done.append("train")
    if which in ("truth", "all"):
        from desco_b200.groundtruth import canonical_count_truth

        canonical_count_truth(DeviceCSR.from_host(gen_mutag_shaped(seed=4, num_graphs=8)), query_ids=STANDARD_QUERY_IDS)
        done.append("truth")
    torch.cuda.synchronize()
    print("sanitize.py ran:", ", ".join(done))


if __name__ == "__main__":
    main()
