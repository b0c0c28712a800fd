"""Host-side operand packing (desco_b200/tcpack.py): the byte layouts the tcgen05 / mma.sync kernels expect.  CPU only."""
import torch

from desco_b200.tcpack import pack_b_operand, pack_dense_tc, pack_mma_b_frags, split_bf16, split_bf16_3, swizzle_rows


def test_split_bf16_reconstructs_to_2pow_minus17():
    torch.manual_seed(0)
    w = torch.randn(64, 64) * 3
    hi, lo = split_bf16(w)
    err = (hi.float() + lo.float() - w).abs().max().item()
    assert err <= 2.0 ** -16 * w.abs().max().item()
    h, m, l = split_bf16_3(w)
    assert torch.equal(h.float() + m.float() + l.float(), w)  # three parts are exact for normal fp32 values


def test_swizzle_128b_places_chunk_c_of_row_r_at_c_xor_r():
    x = torch.arange(16 * 64, dtype=torch.float32).view(16, 64).to(torch.bfloat16)
    img = swizzle_rows(x).view(torch.int16).view(16, 8, 8)  # [row][chunk position][8 bf16]
    src = x.view(torch.int16).view(16, 8, 8)
    for r in (0, 1, 5, 7, 8, 13):
        for c in range(8):
            assert torch.equal(img[r, c ^ (r & 7)], src[r, c]), (r, c)


def test_pack_b_operand_is_hi_image_then_lo_image():
    torch.manual_seed(1)
    w = torch.randn(64, 64)
    img = pack_b_operand(w)
    hi, lo = split_bf16(w)
    assert img.numel() == 2 * 64 * 128
    assert torch.equal(img[: 64 * 128], swizzle_rows(hi)) and torch.equal(img[64 * 128:], swizzle_rows(lo))


def test_pack_dense_tc_block_order():
    torch.manual_seed(2)
    w = torch.randn(288, 128)  # N = 288 (two column blocks of 144), K = 128 (two atoms)
    img = pack_dense_tc(w, 144)
    assert img.numel() == 288 * 128 * 6
    per_image = 144 * 128
    h, m, l = split_bf16_3(w[144:288, 64:128])  # column block 1, K atom 1 -> images 9, 10, 11
    base = (1 * 2 + 1) * 3 * per_image
    assert torch.equal(img[base: base + per_image], swizzle_rows(h))
    assert torch.equal(img[base + per_image: base + 2 * per_image], swizzle_rows(m))
    assert torch.equal(img[base + 2 * per_image: base + 3 * per_image], swizzle_rows(l))


def test_pack_mma_b_frags_layout():
    torch.manual_seed(3)
    w = torch.randn(192, 64)  # [K, N]
    img = pack_mma_b_frags(w)
    assert img.numel() == w.numel() * 4  # same bytes as fp32
    hi, lo = split_bf16(w)
    v = img.view(torch.int16).view(8, 12, 32, 8)  # [column tile][k step][lane][hi k0,k0+1,k0+8,k0+9 | lo ...]
    for nt, ks, lane in ((0, 0, 0), (3, 5, 13), (7, 11, 31)):
        g, t = lane // 4, lane % 4
        k0, n = 16 * ks + 2 * t, 8 * nt + g
        exp = [hi[k0, n], hi[k0 + 1, n], hi[k0 + 8, n], hi[k0 + 9, n], lo[k0, n], lo[k0 + 1, n], lo[k0 + 8, n], lo[k0 + 9, n]]
        for i, e in enumerate(exp):
            assert v[nt, ks, lane, i] == e.view(torch.int16), (nt, ks, lane, i)
