"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (shard ranges, variable-length row all-gather,
assembly of sharded results in centre / node order).  The kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from desco_b200.distributed import all_gather_rows, balanced_shards, node_ranges, shard_range


def test_shard_ranges_cover_without_overlap():
    for n in (0, 1, 7, 100, 4097):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert node_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]


def test_balanced_shards_balance_weight():
    rng = np.random.default_rng(0)
    w = 1.0 + rng.pareto(1.5, size=20000)  # power-law degrees
    for world in (2, 4, 8):
        sh = balanced_shards(w, world)
        assert sh[0][0] == 0 and sh[-1][1] == len(w)
        assert all(a[1] == b[0] for a, b in zip(sh[:-1], sh[1:]))
        tot = np.array([w[a:b].sum() for a, b in sh])
        assert tot.max() / tot.mean() < 1.05 + w.max() / tot.mean()
    assert balanced_shards(np.zeros(0), 4) == [(0, 0)] * 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # variable-length row blocks, 2-D and 3-D payloads (counts [n,Q] and halo scalars [n,Q,4])
        N, Q = 11, 3
        ranges = node_ranges(N, world)
        lo, hi = ranges[rank]
        full = torch.arange(N * Q, dtype=torch.float32).view(N, Q)
        got = all_gather_rows(full[lo:hi].clone())
        assert torch.equal(got, full)
        full3 = torch.arange(N * Q * 4, dtype=torch.float32).view(N, Q, 4)
        got3 = all_gather_rows(full3[lo:hi].clone(), sizes=[b - a for a, b in ranges])
        assert torch.equal(got3, full3)
        # sharded "counting": every rank scores its own centres; the assembled x[N,Q] must be in node order
        deg = np.array([1, 9, 1, 1, 5, 1, 1, 1, 7, 1, 1], dtype=np.float64)
        shards = balanced_shards(1.0 + deg, world)
        a, b = shards[rank]
        centres = torch.arange(a, b)
        keep = centres % 2 == 1  # "neighborhoods without edges are dropped"
        counts = (centres[keep].float() * 10).view(-1, 1).repeat(1, Q)
        x_local = torch.zeros(b - a, Q)
        x_local[centres[keep] - a] = counts
        x = all_gather_rows(x_local, sizes=[d - c for c, d in shards])
        ref = torch.zeros(N, Q)
        ref[1::2] = (torch.arange(1, N, 2).float() * 10).view(-1, 1)
        assert torch.equal(x, ref)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_row_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
