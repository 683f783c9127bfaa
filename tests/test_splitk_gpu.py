"""Split-K partial GEMM (cvb_op_gemm_bf16 epilogue 6) + cvb_op_rmsnorm_reduce against torch.

The pair replaces `o_proj -> + residual -> post_attention_layernorm` and `down_proj -> + residual -> next input_layernorm`
of the action expert (paligemma_with_expert.py:327-355) during PI0FlowMatching.denoise_step (modeling_pi0.py:717-752).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemma_rmsnorm(h, w, eps=1e-6):
    x = h.float()
    x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return (x * (1.0 + w.float())).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K,S", [(200, 1024, 2048, 8), (200, 1024, 2048, 16), (200, 1024, 4096, 16),
                                     (200, 1024, 4096, 5), (40, 1024, 4096, 12), (5, 128, 64, 4), (256, 1152, 4304, 9),
                                     (20, 64, 200, 2), (200, 1024, 2048, 1)])
def test_partials_sum_to_the_product(M, N, K, S):
    from cover_vla_b200 import ops
    torch.manual_seed(M + N + K + S)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    p = ops.gemm_splitk_partial(a, w, S)
    torch.cuda.synchronize()
    kb = (K + 63) // 64
    S_eff = min(S, kb)
    assert p.shape == (S_eff, M, N)
    # every partial is the product over its own balanced run of 64-wide k-blocks (fp32 accumulation of bf16 products)
    for s in range(S_eff):
        k0, k1 = (s * kb // S_eff) * 64, min(((s + 1) * kb // S_eff) * 64, K)
        ref = a[:, k0:k1].float() @ w[:, k0:k1].float().t()
        err = (p[s] - ref).abs().max().item()
        assert err < 2e-4 * max(1.0, ref.abs().max().item()), (s, err)
    ref = a.float() @ w.float().t()
    assert ((p.sum(0) - ref).norm() / ref.norm()).item() < 1e-5


@pytest.mark.parametrize("resid_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("w_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,S", [(200, 1024, 16), (200, 1024, 7), (3, 64, 1), (40, 2048, 4)])
def test_rmsnorm_reduce_matches_the_ledger(M, N, S, resid_dtype, w_dtype):
    from cover_vla_b200 import ops
    torch.manual_seed(M + N + S)
    p = torch.randn(S, M, N, device="cuda", dtype=torch.float32)
    resid = torch.randn(M, N, device="cuda").to(resid_dtype)
    w = (0.1 * torch.randn(N, device="cuda")).to(w_dtype)
    h, y = ops.rmsnorm_reduce(p, resid, w)
    torch.cuda.synchronize()
    acc = torch.zeros(M, N, device="cuda")
    for s in range(S):  # split order, fp32
        acc = acc + p[s]
    h_ref = (acc.to(torch.bfloat16).float() + resid.float()).to(torch.bfloat16)
    assert torch.equal(h, h_ref)
    y_ref = _gemma_rmsnorm(h_ref, w)
    # statistics are fp32 in both; the block reduction order differs from torch's -> allow one bf16 ulp
    diff = (y.float() - y_ref.float()).abs()
    assert (diff <= 2.0 ** -7 * y_ref.float().abs() + 1e-6).all()
    assert (diff > 0).float().mean().item() < 0.02


def test_pair_matches_fused_epilogue_gemm():
    """o_proj/down_proj through partials + reduce == the EPI_RESID GEMM followed by rmsnorm, up to accumulation order."""
    from cover_vla_b200 import ops
    torch.manual_seed(3)
    M, N, K = 200, 1024, 4096
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    resid = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    g = (0.1 * torch.randn(N, device="cuda")).to(torch.bfloat16)
    h_fused = ops.gemm_bf16(a, w, epilogue=ops.EPI_RESID, resid=resid, force_bn=64)
    h, y = ops.rmsnorm_reduce(ops.gemm_splitk_partial(a, w, 16), resid, g)
    torch.cuda.synchronize()
    d = (h.float() - h_fused.float()).abs()
    # <= 1 bf16 ulp of bf16(acc) (accumulation order differs), then <= 1 ulp of the rounded sum
    mag = (h_fused.float() - resid.float()).abs() + h_fused.float().abs()
    assert (d <= 2.0 ** -6 * mag + 1e-6).all()
    assert (d > 0).float().mean().item() < 0.05
    assert ((y.float() - _gemma_rmsnorm(h, g).float()).abs() <= 2.0 ** -7 * y.float().abs() + 1e-6).all()


@pytest.mark.parametrize("M,N,S,with_bias", [(256, 1152, 8, True), (256, 1152, 6, True), (16, 64, 1, False), (100, 288, 3, True)])
def test_layernorm_reduce_matches_the_ledger(M, N, S, with_bias):
    from cover_vla_b200 import ops
    torch.manual_seed(M + N + S)
    p = torch.randn(S, M, N, device="cuda", dtype=torch.float32)
    bias = torch.randn(N, device="cuda").to(torch.bfloat16) if with_bias else None
    resid = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    w = (1.0 + 0.1 * torch.randn(N, device="cuda")).to(torch.bfloat16)
    b = (0.1 * torch.randn(N, device="cuda")).to(torch.bfloat16)
    h, y = ops.layernorm_reduce(p, bias, resid, w, b)
    torch.cuda.synchronize()
    acc = torch.zeros(M, N, device="cuda")
    for s in range(S):
        acc = acc + p[s]
    if with_bias:
        acc = acc + bias.float()
    h_ref = (acc.to(torch.bfloat16).float() + resid.float()).to(torch.bfloat16)
    assert torch.equal(h, h_ref)
    y_ref = torch.nn.functional.layer_norm(h_ref.float(), (N,), w.float(), b.float(), 1e-6).to(torch.bfloat16)
    diff = (y.float() - y_ref.float()).abs()
    assert (diff <= 2.0 ** -7 * y_ref.float().abs() + 2e-3).all()
    assert (diff > 0).float().mean().item() < 0.05


def test_siglip_block_tail_matches_fused_epilogue_gemm():
    """fc2 (K = 4304, ragged last k-block) through partials + LayerNorm-reduce vs the EPI_RESID GEMM."""
    from cover_vla_b200 import ops
    torch.manual_seed(5)
    M, N, K = 256, 1152, 4304
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda").to(torch.bfloat16)
    resid = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    g = torch.ones(N, device="cuda", dtype=torch.bfloat16)
    z = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    h_fused = ops.gemm_bf16(a, w, bias=bias, epilogue=ops.EPI_RESID, resid=resid, force_bn=64)
    h, _ = ops.layernorm_reduce(ops.gemm_splitk_partial(a, w, 8), bias, resid, g, z)
    torch.cuda.synchronize()
    d = (h.float() - h_fused.float()).abs()
    mag = (h_fused.float() - resid.float()).abs() + h_fused.float().abs()
    assert (d <= 2.0 ** -6 * mag + 1e-6).all()
    assert (d > 0).float().mean().item() < 0.05


@pytest.mark.parametrize("M,N,xdt,wdt", [(2240, 2048, torch.bfloat16, torch.bfloat16), (17920, 2048, torch.bfloat16, torch.bfloat16),
                                         (600, 1024, torch.bfloat16, torch.bfloat16), (513, 256, torch.bfloat16, torch.bfloat16),
                                         (200, 1024, torch.bfloat16, torch.bfloat16), (300, 2048, torch.float32, torch.float32),
                                         (1000, 1792, torch.bfloat16, torch.bfloat16), (700, 640, torch.bfloat16, torch.float32)])
def test_rmsnorm_row_kernels(M, N, xdt, wdt):
    """GemmaRMSNorm alone: the warp-per-row kernel (bf16, width a multiple of 256 up to 2048, >= 512 rows) and the
    CTA-per-row kernel give the ledger value bf16((x * r) * (1 + w)) with r from an fp32 sum of squares."""
    from cover_vla_b200 import ops
    torch.manual_seed(M + N)
    x = (torch.randn(M, N, device="cuda") * 3).to(xdt)
    w = (torch.randn(N, device="cuda") * 0.2).to(wdt)
    y = ops.rmsnorm(x, w)
    torch.cuda.synchronize()
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)
    ref = ((xf * r) * (1.0 + w.float())).to(torch.bfloat16)
    # r differs from torch's in the last bit (summation order): at most one bf16 ulp on a few elements
    diff = (y.float() - ref.float()).abs()
    assert (diff <= 2.0 ** -7 * ref.float().abs() + 1e-6).all()
    assert (y != ref).float().mean().item() < 0.02
    # strided input rows (a column slice of a wider buffer)
    big = torch.zeros(M, N + 64, device="cuda", dtype=xdt)
    big[:, :N] = x
    assert torch.equal(ops.rmsnorm(big[:, :N], w), y)
