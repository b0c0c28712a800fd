"""GPU: the tcgen05 plumbing (UMMA descriptors, SWIZZLE_128B operand images, TMEM, bulk copy) against a float64 GEMM."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(n, passes, seed):
    from desco_b200 import _lib
    from desco_b200.data import _ptr, _stream
    from desco_b200.tcpack import pack_b_operand

    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(128, 64, generator=g) * 3.0
    w = torch.randn(n, 64, generator=g)
    img = pack_b_operand(w).cuda()
    d = torch.full((128, n), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.desco_tc_selftest(_ptr(a.cuda()), _ptr(img), n, passes, _ptr(d), _ptr(status), _stream()), "selftest")
    torch.cuda.synchronize()
    assert int(status.item()) == 0, "tcgen05 self test timed out on an mbarrier"
    ref = a.double() @ w.double().t()
    return d.cpu().double(), ref


@pytest.mark.parametrize("n", [32, 64, 192, 256])
def test_bf16x3_matches_fp64(cuda_device, n):
    d, ref = _run(n, 3, seed=n)
    err = (d - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


def test_single_pass_is_bf16_accurate_only(cuda_device):
    d, ref = _run(192, 1, seed=7)
    err = (d - ref).abs().max().item() / ref.abs().max().item()
    assert 1e-5 < err < 2e-2, err  # one bf16 pass: ~2^-9 per operand - confirms the extra passes do the work
