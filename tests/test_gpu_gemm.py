"""tcgen05 GEMM (trs_gemm_bf16_tn) against torch on the same bf16 operands (``-m gpu``).

The kernel multiplies bf16 inputs exactly and accumulates in fp32, so against an fp32 matmul of the
SAME bf16-rounded operands only the summation order differs: tolerance rtol 1e-4 / atol 1e-4 *
sqrt(k) for fp32 output; bf16 output adds one bf16 rounding (rtol 2^-8).  Integer-valued fixtures
(every product and partial sum exact in fp32) must match bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _L():
    from torchrecsys_b200 import _lib
    return _lib


SHAPES = [(128, 128, 64), (256, 512, 128), (384, 256, 512), (100, 72, 40), (1000, 136, 200),
          (32768, 512, 128), (128, 64, 64), (4096, 32, 256), (130, 8, 8)]


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_gemm_exact_on_integer_fixture(dev, m, n, k):
    g = torch.Generator(device="cpu").manual_seed(m * 7 + n * 3 + k)
    a = torch.randint(-4, 5, (m, k), generator=g).float()
    b = torch.randint(-4, 5, (n, k), generator=g).float()
    bias = torch.randint(-8, 9, (n,), generator=g).float()
    want = a @ b.t() + bias
    out = torch.empty((m, n), dtype=torch.float32, device=dev)
    _L().gemm_bf16_tn(a.to(dev).bfloat16(), b.to(dev).bfloat16(), out, bias=bias.to(dev))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), want)


@pytest.mark.parametrize("m,n,k", SHAPES[:6])
def test_gemm_random_fp32_and_bf16_out(dev, m, n, k):
    g = torch.Generator(device="cpu").manual_seed(k)
    a = torch.randn((m, k), generator=g).to(dev).bfloat16()
    b = torch.randn((n, k), generator=g).to(dev).bfloat16()
    want = a.float() @ b.float().t()
    out = torch.empty((m, n), dtype=torch.float32, device=dev)
    _L().gemm_bf16_tn(a, b, out)
    torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-4 * k ** 0.5)
    outb = torch.empty((m, n), dtype=torch.bfloat16, device=dev)
    _L().gemm_bf16_tn(a, b, outb)
    torch.testing.assert_close(outb.float(), want, rtol=2 ** -7, atol=1e-2 * k ** 0.5)


def test_gemm_split_k_partials_sum_to_the_product(dev):
    m, n, k, splits = 512, 128, 32768, 37
    g = torch.Generator(device="cpu").manual_seed(1)
    a = torch.randint(-2, 3, (m, k), generator=g).float()
    b = torch.randint(-2, 3, (n, k), generator=g).float()
    out = torch.empty((splits, m, n), dtype=torch.float32, device=dev)
    _L().gemm_bf16_tn(a.to(dev).bfloat16(), b.to(dev).bfloat16(), out, splits=splits)
    assert torch.equal(out.sum(0).cpu(), a @ b.t())


def test_gemm_strided_operands_and_column_statistics(dev):
    """Operands that are column slices of wider buffers (lda/ldb != k), bf16 output with bias and the
    BatchNorm column partials over the valid rows of two stacked, padded halves."""
    B, Bpad, n, k = 300, 384, 256, 128
    g = torch.Generator(device="cpu").manual_seed(3)
    abuf = torch.randn((2 * Bpad, k + 64), generator=g).to(dev).bfloat16()
    bbuf = torch.randn((n, k + 8), generator=g).to(dev).bfloat16()
    a, b = abuf[:, 64:], bbuf[:, :k]
    bias = torch.randn((n,), generator=g).to(dev)
    out = torch.zeros((2 * Bpad, n), dtype=torch.bfloat16, device=dev)
    tiles = 2 * Bpad // 128
    cs = torch.zeros((tiles, n), device=dev)
    css = torch.zeros((tiles, n), device=dev)
    _L().gemm_bf16_tn(a, b, out, bias=bias, col_sum=cs, col_sumsq=css, rows_per_half=Bpad, rows_valid=B)
    want = (a.float() @ b.float().t() + bias)
    torch.testing.assert_close(out.float(), want, rtol=2 ** -7, atol=0.1)
    z = out.float().view(2, Bpad, n)[:, :B]  # what the next kernel reads back
    got_sum = cs.view(2, Bpad // 128, n).sum(1)
    got_sq = css.view(2, Bpad // 128, n).sum(1)
    torch.testing.assert_close(got_sum, z.sum(1), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(got_sq, (z * z).sum(1), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("m,n,k,a_mn,b_mn", [
    (256, 128, 128, False, True), (1000, 136, 200, False, True), (4096, 128, 512, False, True),
    (4096, 32, 64, False, True), (128, 128, 128, True, True), (512, 128, 4096, True, True),
    (256, 136, 1000, True, True), (32, 16, 640, True, True), (512, 512, 2048, True, True)])
def test_gemm_mn_major_operands(dev, m, n, k, a_mn, b_mn):
    """dgrad / wgrad operand layouts: the operand is handed over as stored ([k, m] / [k, n])."""
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    a = torch.randint(-4, 5, (m, k), generator=g).float()
    b = torch.randint(-4, 5, (n, k), generator=g).float()
    want = a @ b.t()
    ad = (a.t().contiguous() if a_mn else a).to(dev).bfloat16()
    bd = (b.t().contiguous() if b_mn else b).to(dev).bfloat16()
    splits = 4 if k >= 1000 else 1
    out = torch.empty((splits, m, n) if splits > 1 else (m, n), dtype=torch.float32, device=dev)
    _L().gemm_bf16_tn(ad, bd, out, splits=splits, a_mn=a_mn, b_mn=b_mn)
    got = out.sum(0) if splits > 1 else out
    assert torch.equal(got.cpu(), want)
