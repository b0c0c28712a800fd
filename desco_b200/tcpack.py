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
