"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (shard ranges, variable-length row all-gather,
assembly of sharded results in centre / node order).  The kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from desco_b200.distributed import (LocalComm, ProcessGroupComm, all_gather_rows, balanced_shards, centre_work_estimate,
                                    gossip_shard_plan, node_ranges, shard_range)


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


def test_gossip_shard_plan_geometry():
    for n, q, world, qg in ((1000, 29, 2, 4), (1_000_000, 29, 8, 4), (5, 3, 4, 8), (128 * 8, 29, 8, 5)):
        plan = gossip_shard_plan(n, q, world, qg)
        assert plan.n_loc % 128 == 0 and plan.n_rows == plan.n_loc * world >= n
        assert plan.ranges[0][0] == 0 and plan.ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(plan.ranges[:-1], plan.ranges[1:]))
        assert all(0 <= hi - lo <= plan.n_loc for lo, hi in plan.ranges)
        assert plan.groups[0][0] == 0 and plan.groups[-1][1] == q
        assert all(a[1] == b[0] for a, b in zip(plan.groups[:-1], plan.groups[1:]))
        assert max(b - a for a, b in plan.groups) == min(qg, q)
        assert plan.halo_bytes() == 16 * q * plan.n_loc * (world - 1)


def test_local_comm_emulates_the_block_all_gather():
    comm = LocalComm(3)
    n_loc = 4
    bufs = []
    for r in range(3):
        b = torch.zeros(3 * n_loc, 2)
        b[r * n_loc:(r + 1) * n_loc] = r + 1
        bufs.append(b)
    works = [comm.for_rank(r).all_gather_block(bufs[r], n_loc, "t") for r in range(3)]
    for w in works:
        w.wait()
    want = torch.arange(1, 4).repeat_interleave(n_loc).view(-1, 1).repeat(1, 2).float()
    assert all(torch.equal(b, want) for b in bufs)


def test_centre_work_estimate_tracks_ball_and_position():
    from types import SimpleNamespace

    # star: node 9 is the hub of 0..8; path 10-11
    edges = [(9, i) for i in range(9)] + [(10, 11)]
    n = 12
    adj = [[] for _ in range(n)]
    for a, b in edges:
        adj[a].append(b)
        adj[b].append(a)
    rowptr = torch.tensor(np.cumsum([0] + [len(a) for a in adj]), dtype=torch.int32)
    col = torch.tensor([v for a in adj for v in sorted(a)], dtype=torch.int32)
    g = SimpleNamespace(rowptr=rowptr, col=col, num_graphs=1, graph_ptr=torch.tensor([0, n], dtype=torch.int32))
    w1, w2 = centre_work_estimate(g, 1), centre_work_estimate(g, 2)
    assert w1[9] > w1[8] and w2[0] > w1[0]  # leaves see the hub's adjacency only from depth 2 on
    assert np.isclose(w2[0] / (0.25 + 0.0), 1 + 1 + 9) and np.isclose(w2[9] / (0.25 + 0.75 * 9 / 12), 1 + 9 + 9)


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
        # the in-place block all-gather of the sharded gossip forward (one per query group), issued asynchronously
        plan = gossip_shard_plan(300, 7, world, 3)
        comm = ProcessGroupComm()
        assert (comm.rank, comm.world) == (rank, world)
        blocks = []
        for gi, (q0, q1) in enumerate(plan.groups):
            full = torch.zeros(plan.n_rows, q1 - q0, 4)
            full[rank * plan.n_loc:(rank + 1) * plan.n_loc] = 10 * (rank + 1) + gi
            blocks.append((full, comm.all_gather_block(full, plan.n_loc, ("s4", gi))))
        for gi, (full, work) in enumerate(blocks):
            work.wait()
            for r in range(world):
                assert torch.all(full[r * plan.n_loc:(r + 1) * plan.n_loc] == 10 * (r + 1) + gi)
        # bench.py's replicated-input check: a rank whose generated copy differs (here: another size, another value)
        # gets rank 0's; identical copies are left alone
        import sys
        import types

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench

        bctx = types.SimpleNamespace(torch=torch, dist=dist, dev=torch.device("cpu"), world=world, rank=rank)
        col = torch.arange(20 if rank == 0 else 22, dtype=torch.int32)
        (rp, c2), info = bench.replicate_inputs(bctx, [torch.arange(11, dtype=torch.int32), col], "csr")
        assert info["ranks_whose_generated_copy_differed_from_rank0"] == [1] and torch.equal(c2, torch.arange(20, dtype=torch.int32))
        xr = torch.ones(10, 29) + (0.0 if rank == 0 else 1e-3)
        (x2,), info = bench.replicate_inputs(bctx, [xr], "x")
        assert info["ranks_whose_generated_copy_differed_from_rank0"] == [1] and torch.equal(x2, torch.ones(10, 29))
        (y2,), info = bench.replicate_inputs(bctx, [torch.ones(5, 3)], "y")
        assert info["ranks_whose_generated_copy_differed_from_rank0"] == [] and info["action"] == "none"
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
