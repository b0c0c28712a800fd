"""CPU: the arithmetic the tensor-path gossip gather (csrc/gossip.cu gossip_gather_kernel) relies on, emulated in numpy.
  (1) both gates are positive, so g1 * sum_{j<i} relu(z_j) + (1-g1) * sum_{j>i} relu(z_j) = sum_j relu(w_j z_j);
  (2) the 4-deep dot products as tf32 tensor-core blocks: weights as tf32 hi (round to nearest) + lo, the neighbour scalars
      as tf32 hi (truncation) + lo (truncated by the hardware), all four partial products accumulated in fp32 - within
      2e-6 of the float64 value relative to the size of the terms, i.e. inside the 1e-4 budget with room to spare;
  (3) a neighbour slot that is absent contributes exactly zero (the octet of eight rows pads the shorter rows)."""
import numpy as np


def _tf32_trunc(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _tf32_rna(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x1000)) & np.uint64(0xFFFFE000)
    return u.astype(np.uint32).view(np.float32)


def _block(vals, W):
    """vals [n, 4] gated neighbour records, W [4, 64] -> what the two chained MMAs produce, [n, 64] (fp32 accumulate)."""
    w_hi = _tf32_rna(W)
    w_lo = _tf32_trunc((W - w_hi).astype(np.float32))  # the tensor core reads the top 19 bits of the lo operand
    b_hi = _tf32_trunc(vals)
    b_lo = _tf32_trunc((vals - b_hi).astype(np.float32))
    acc = np.zeros((vals.shape[0], W.shape[1]), dtype=np.float32)
    for b in (b_hi, b_lo):  # first MMA: hi parts of the records, second: lo parts; K slots: weights hi | weights lo
        for w in (w_hi, w_lo):
            acc = (acc.astype(np.float64) + b.astype(np.float64) @ w.astype(np.float64)).astype(np.float32)
    return acc


def test_gate_moves_inside_the_relu():
    rng = np.random.default_rng(0)
    z = rng.normal(size=(50, 64))
    lt = rng.random(50) < 0.4
    for g1 in (0.0, 1e-6, 0.31, 0.5, 0.999, 1.0):
        split = g1 * np.maximum(z[lt], 0).sum(0) + (1 - g1) * np.maximum(z[~lt], 0).sum(0)
        w = np.where(lt, g1, 1 - g1)[:, None]
        one = np.maximum(w * z, 0).sum(0)
        assert np.allclose(split, one, rtol=1e-12, atol=1e-12)


def test_tf32_hi_lo_blocks_are_fp32_grade():
    rng = np.random.default_rng(1)
    W = rng.normal(size=(4, 64)).astype(np.float32)
    for scale in (1.0, 1e-3, 3e4):  # degree mixes reach 1e4-1e5 on a power-law target, counts are O(1)
        a = (rng.normal(size=(4096, 4)) * scale).astype(np.float32)
        a[:, 3] = 1.0
        gate = np.where(rng.random(4096) < 0.5, 0.37, 0.63).astype(np.float32)
        vals = (a * gate[:, None]).astype(np.float32)
        got = _block(vals, W)
        ref = vals.astype(np.float64) @ W.astype(np.float64)
        size = np.abs(vals).astype(np.float64) @ np.abs(W).astype(np.float64)  # sum of |terms|
        assert np.max(np.abs(got - ref) / size) < 2e-6


def test_absent_neighbour_slots_add_exactly_zero():
    W = np.random.default_rng(2).normal(size=(4, 64)).astype(np.float32)
    z = _block(np.zeros((8, 4), dtype=np.float32), W)
    assert not z.any()
    acc = np.float32(1.2345678) + np.maximum(z, 0)
    assert (acc == np.float32(1.2345678)).all()


def test_tile_decomposition_covers_every_neighbour_once():
    """The index arithmetic of gossip_gather_kernel's work decomposition, restated: a 128-row tile is split into hub rows
    (> 512 neighbours, slabs of 256 dealt over the CTA's warps), wide rows (33..512, eight interleaved sub-rows, slabs of
    256) and octets of <= 32-neighbour rows in descending degree; every adjacency entry of every row must be visited
    exactly once and the per-sub-row counts the kernel derives must add up."""
    TR, OCT, HUB, NW = 128, 32, 512, 4
    rng = np.random.default_rng(3)
    for trial in range(20):
        deg = np.minimum((rng.pareto(1.2, TR) * 6).astype(np.int64), 3000)
        deg[rng.integers(0, TR, 5)] = [0, 32, 33, 512, 513]
        bucket = np.where(deg > HUB, 0, np.where(deg > OCT, 1, 2 + OCT - deg))
        order = np.argsort(bucket, kind="stable")  # any order inside a bucket (the kernel's is atomic-slot order)
        n_hub, n_wide = int((bucket == 0).sum()), int((bucket == 1).sum())
        n_items = n_wide + (TR - n_hub - n_wide + 7) // 8
        seen = [np.zeros(d, dtype=np.int64) for d in deg]
        rows_done = np.zeros(TR, dtype=np.int64)
        for item in range(n_items):
            if item < n_wide:
                r = order[n_hub + item]
                d = int(deg[r])
                e0 = 0
                while e0 < d:  # walk_wide(first = 0, step = 1)
                    left = d - e0
                    nblk = min(OCT, (left + 7) >> 3)
                    for g in range(8):
                        n_own = min(OCT, max(0, (left - g + 7) >> 3))
                        assert n_own <= nblk
                        for k in range(n_own):
                            seen[r][e0 + g + 8 * k] += 1
                    e0 += 8 * OCT
                rows_done[r] += 1
            else:
                base = n_hub + n_wide + 8 * (item - n_wide)
                rows = [order[base + g] for g in range(8) if base + g < TR]
                nblk = max(int(deg[r]) for r in rows)
                assert nblk <= OCT
                for r in rows:
                    seen[r][:int(deg[r])] += 1
                    rows_done[r] += 1
        for h in range(n_hub):
            r = order[h]
            d = int(deg[r])
            for warp in range(NW):  # walk_wide(first = warp, step = NW)
                e0 = warp * 8 * OCT
                while e0 < d:
                    left = d - e0
                    for g in range(8):
                        for k in range(min(OCT, max(0, (left - g + 7) >> 3))):
                            seen[r][e0 + g + 8 * k] += 1
                    e0 += NW * 8 * OCT
            rows_done[r] += 1
        assert (rows_done == 1).all()
        assert all((s == 1).all() for s in seen)
