"""Host-side packing of tensor-core operands for the tcgen05 kernels (csrc/tc05.cuh conventions).

A "B operand" is a weight matrix ``W[n, k]`` with k = 64 contiguous (nn.Linear's own [out, in] layout).  It is split
into bf16 ``hi = bf16(W)`` and ``lo = bf16(W - hi)`` and each part is written as the exact shared-memory image the
kernel needs: 128 bytes per row, rows in groups of 8, the 16-byte chunk c of row r stored at chunk position
``c ^ (r & 7)`` (SWIZZLE_128B).  The kernel then fetches an image with one bulk async copy - no tensor map, no
in-kernel shuffling.
"""
from __future__ import annotations

import torch


def split_bf16(w: torch.Tensor):
    """fp32/fp64 -> (hi, lo) bf16 with hi + lo = w up to 2^-17 relative."""
    w32 = w.to(torch.float32)
    hi = w32.to(torch.bfloat16)
    lo = (w32 - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def swizzle_rows(x: torch.Tensor) -> torch.Tensor:
    """x: [n, 64] bf16 -> uint8 [n * 128] SWIZZLE_128B image (n % 8 == 0)."""
    n, k = x.shape
    assert k == 64 and n % 8 == 0, (n, k)
    chunks = x.contiguous().view(torch.int16).view(n, 8, 8)  # [row, chunk, 8 bf16]
    r = torch.arange(n).view(n, 1)
    pos = torch.arange(8).view(1, 8) ^ (r & 7)  # destination chunk position of source chunk c
    out = torch.empty_like(chunks)
    out[r.expand(n, 8), pos] = chunks
    return out.view(torch.uint8).reshape(-1)


def pack_b_operand(w: torch.Tensor) -> torch.Tensor:
    """W[n, 64] -> uint8 image: hi rows (n*128 bytes) then lo rows (n*128 bytes)."""
    hi, lo = split_bf16(w.detach().cpu())
    return torch.cat([swizzle_rows(hi), swizzle_rows(lo)])


def split_bf16_3(w: torch.Tensor):
    """fp32 -> (hi, mid, lo) bf16 with hi + mid + lo == w exactly for normal fp32 values (3 x 8 significant bits)."""
    w32 = w.to(torch.float32)
    hi = w32.to(torch.bfloat16)
    r1 = w32 - hi.to(torch.float32)
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.to(torch.float32)).to(torch.bfloat16)
    return hi, mid, lo


def pack_dense_tc(w: torch.Tensor, nblk: int) -> torch.Tensor:
    """nn.Linear weight W[N, K] -> images for csrc/dense_tc.cu: [column block of nblk][64-wide K atom][hi | mid | lo],
    each image nblk * 128 bytes (3-way bf16 split: the dense kernel runs up to 6 passes = fp32-grade products)."""
    w = w.detach().cpu()
    n, k = w.shape
    assert n % nblk == 0 and k % 64 == 0, (n, k, nblk)
    parts = []
    for nb in range(n // nblk):
        for a in range(k // 64):
            blk = w[nb * nblk:(nb + 1) * nblk, a * 64:(a + 1) * 64]
            parts += [swizzle_rows(x) for x in split_bf16_3(blk)]
    return torch.cat(parts)


def pack_mma_b_frags(w: torch.Tensor) -> torch.Tensor:
    """W[K, N] (K-major weight, K % 16 == 0, N % 8 == 0) -> uint8 image of warp-level ``mma.sync.m16n8k16`` B fragments,
    bf16 hi/lo split: for column tile nt, k-step ks and lane (g = lane // 4, t = lane % 4) one 16-byte record
    ``{hi[k0:k0+2], hi[k0+8:k0+10], lo[k0:k0+2], lo[k0+8:k0+10]}`` of column ``8 nt + g`` with ``k0 = 16 ks + 2 t``
    (csrc/shmp_fused.cu: the canonical rows of a tile).  Same byte count as the fp32 matrix."""
    w = w.detach().cpu()
    k, n = w.shape
    assert k % 16 == 0 and n % 8 == 0, (k, n)
    hi, lo = split_bf16(w)
    hi, lo = hi.view(torch.int16), lo.view(torch.int16)
    lane = torch.arange(32)
    g, t = lane // 4, lane % 4
    k0 = (16 * torch.arange(k // 16).view(1, -1, 1) + 2 * t.view(1, 1, 32)).expand(n // 8, k // 16, 32)
    col = (8 * torch.arange(n // 8).view(-1, 1, 1) + g.view(1, 1, 32)).expand(n // 8, k // 16, 32)
    parts = []
    for src in (hi, lo):
        for off in (0, 8):
            parts += [src[k0 + off, col], src[k0 + off + 1, col]]
    out = torch.stack(parts, dim=-1)  # [nt, ks, lane, 8 bf16]
    return out.contiguous().view(torch.uint8).reshape(-1)
