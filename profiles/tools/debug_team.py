import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from desco_b200 import _lib
from desco_b200.data import DeviceCSR, partition_batch
from desco_b200.graph import gen_enzymes_shaped, gen_imdb_shaped
from oracle import partition as P
lib = _lib.load()
_lib.check(lib.desco_partition_large_set_caps(9, 16, 8, 12, 512, 256), "caps")
KEYS = ("nbh_ptr", "node_gid", "edge_ptr", "edge_col", "edge_tri", "centre", "indicator")
for gen, kw in ((gen_enzymes_shaped, dict(num_graphs=40)), (gen_imdb_shaped, dict(num_graphs=25))):
    csr = gen(seed=3, **kw)
    d = DeviceCSR.from_host(csr)
    for mode in ("hetero", "canonical"):
        ref = P.partition_dataset(csr, 4, mode=mode)
        for rep in range(3):
            b = partition_batch(d, None, 4, mode, large=True)
            tier = b._cache["tier"].cpu().numpy()
            b = b.to_numpy()
            for k in KEYS:
                if not np.array_equal(b[k], ref[k]):
                    bad = np.nonzero(b[k] != ref[k])[0]
                    print(gen.__name__, mode, rep, k, "mismatches", len(bad), "first", bad[:5], b[k][bad[:5]], ref[k][bad[:5]])
                    if k == "edge_col":
                        rows = np.searchsorted(ref["edge_ptr"], bad[:5], side="right") - 1
                        g = np.searchsorted(ref["nbh_ptr"], rows, side="right") - 1
                        print("  rows", rows, "nbh", g, "nbh sizes", np.diff(ref["nbh_ptr"])[g], "centres", ref["centre"][g])
                    break
            else:
                print(gen.__name__, mode, rep, "ok; team centres", int((tier == 1).sum()), "of", len(tier))
