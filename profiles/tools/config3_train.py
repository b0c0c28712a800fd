"""Config 3 (BASELINE.json configs[2]): Syn_1827-shaped synthetic set, SHMP neighborhood-counting training step
(forward + backward + Adam, `lightning_model.py:228-254,160-173`) on batches of 512 neighborhoods (`config.py:255`).

Prints one JSON line: training steps/s and neighborhoods/s on the GPU (CUDA events, L2 flushed between steps), with the
oracle's autograd step timed on the host cores beside it (`--cpu-steps`, a bounded sample).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stride", type=int, default=12, help="every stride-th graph id of the 1827 (SURVEY App. C)")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-steps", type=int, default=3)
    args = ap.parse_args()

    from desco_b200 import _lib
    from desco_b200.data import DeviceCSR, partition_batch
    from desco_b200.graph import gen_syn1827_shaped
    from desco_b200.lightning_model import STANDARD_QUERY_IDS, NeighborhoodCountingModel
    from desco_b200.training import FusedAdam

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    csr = gen_syn1827_shaped(seed=0, stride=args.stride)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    full = partition_batch(DeviceCSR.from_host(csr), None, 4, "hetero")
    b.record()
    torch.cuda.synchronize()
    part_ms = a.elapsed_time(b)
    G = full.num_neighborhoods
    nb = G // args.batch
    rng = np.random.default_rng(3)
    y_all = torch.from_numpy(np.floor(np.exp(rng.normal(0.0, 1.5, size=(G, 29)))).astype(np.float32)).to(dev)
    batches = []
    for i in range(min(nb, args.steps)):
        bt = full.slice(i * args.batch, (i + 1) * args.batch)
        sizes = (bt.nbh_ptr[1:] - bt.nbh_ptr[:-1])
        bt.max_rows = int(sizes.max())
        bt.y = y_all[i * args.batch:(i + 1) * args.batch].contiguous()
        batches.append(bt)

    torch.manual_seed(0)
    model = NeighborhoodCountingModel().to(dev).train()
    model.set_queries(STANDARD_QUERY_IDS)
    model.set_pyg_batch_size(args.batch)
    opt = FusedAdam(model.parameters(), lr=1e-4)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(bt):
        opt.zero_grad()
        loss = model.training_step(bt, 0)
        loss.backward()
        opt.step()
        return loss

    # warm-up = one untimed pass over the same batches (an epoch: the batches differ in size by 30x, and the first visit of
    # each size pays cudaMalloc inside the caching allocator; training runs hundreds of epochs over the same loader)
    for i in range(max(args.warmup, 1)):
        for bt in batches:
            step(bt)
    torch.cuda.synchronize()
    launches0 = int(lib.desco_kernel_launches())
    evs, losses = [], []
    for i in range(args.steps):
        bt = batches[i % len(batches)]
        flush.fill_(i & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        losses.append(step(bt))
        e1.record()
        evs.append((e0, e1, bt))
    torch.cuda.synchronize()
    launches = int(lib.desco_kernel_launches()) - launches0
    ms = [e0.elapsed_time(e1) for e0, e1, _ in evs]
    ms_step = float(np.mean(ms))
    rows = float(np.mean([bt.num_rows for _, _, bt in evs]))
    edges = float(np.mean([bt.num_edges for _, _, bt in evs]))

    cpu = None
    if args.cpu_steps > 0:
        from oracle import model as M

        torch.set_num_threads(os.cpu_count())
        torch.manual_seed(0)
        om = M.NeighborhoodCountingModel().train()
        oopt = torch.optim.Adam(om.parameters(), lr=1e-4)
        qb = M.query_batch()
        t = 0.0
        pick = np.linspace(0, len(batches) - 1, args.cpu_steps).round().astype(int) if args.cpu_steps > 1 else [len(batches) // 2]
        for i in pick:
            bt = batches[int(i)]
            b_np = bt.to_numpy()
            y = bt.y.cpu()
            t0 = time.perf_counter()
            oopt.zero_grad()
            loss = om.train_forward(b_np, qb, y, pyg_batch_size=args.batch)
            loss.backward()
            oopt.step()
            t += time.perf_counter() - t0
        cpu = {"value": args.batch * args.cpu_steps / t, "unit": "neighborhoods/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{args.cpu_steps} training step(s) of {args.batch} neighborhoods (batches spread evenly over the timed set), torch autograd on the oracle, all host cores",
               "ms_per_step": 1e3 * t / args.cpu_steps}

    print(json.dumps({
        "workload": f"syn1827_shaped_stride{args.stride}_train_step_batch{args.batch}", "metric": "training_neighborhoods_per_sec",
        "value": args.batch / (ms_step * 1e-3), "unit": "neighborhoods/s", "steps_per_s": 1e3 / ms_step, "ms_per_step": ms_step,
        "ms_min": float(np.min(ms)), "ms_max": float(np.max(ms)), "steps": args.steps, "warmup": args.warmup,
        "graphs": int(len(csr.graph_ptr) - 1), "target_nodes": int(csr.num_nodes), "neighborhoods_total": G,
        "max_rows": int(full.max_rows), "partition_ms_whole_set": part_ms,
        "rows_per_step": rows, "directed_edges_per_step": edges, "gpu_launches_per_step": launches / args.steps,
        "loss_first": float(losses[0].item()), "loss_last": float(losses[-1].item()),
        "l2": "flushed between steps (256 MiB write)", "cpu_baseline": cpu,
    }))


if __name__ == "__main__":
    main()
