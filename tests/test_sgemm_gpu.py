"""fp32 linear (cvb_op_sgemm_f32) against torch float64: the path's fp32 islands (modeling_pi0.py:598-609 suffix MLP,
efficient_ensemble_merged.py:194-247 verifier heads) must stay true fp32 (score tolerance 1e-3 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# (M, N, K): denoise-loop MLP, verifier trajectory encoder, ragged tiles, the small-K fallback kernel
SHAPES = [(160, 1024, 1024), (400, 1536, 512), (400, 2048, 512), (400, 512, 2048), (64, 1024, 1024),
          (37, 70, 132), (1, 1024, 32), (400, 512, 7), (33, 65, 64), (129, 131, 260)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_sgemm_matches_fp64(M, N, K, act):
    from cover_vla_b200 import ops
    torch.manual_seed(M + 3 * N + 7 * K + act)
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    bias, row_bias = torch.randn(N, device="cuda"), torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    y = ops.sgemm_f32(a, w, bias=bias, row_bias=row_bias, resid=resid, act=act)
    torch.cuda.synchronize()
    z = a.double() @ w.double().t() + bias.double() + row_bias.double()
    if act == 1:
        z = torch.relu(z)
    elif act == 2:
        z = torch.nn.functional.gelu(z)
    elif act == 3:
        z = torch.nn.functional.silu(z)
    z = z + resid.double()
    err = (y.double() - z).abs().max().item()
    assert err < 2e-5 * max(1.0, K ** 0.5 / 8), (M, N, K, act, err)


def test_sgemm_strided_and_plain():
    from cover_vla_b200 import ops
    torch.manual_seed(5)
    big_a = torch.randn(160, 2048, device="cuda")
    big_w = torch.randn(1024, 2048, device="cuda") / 32
    a, w = big_a[:, :1024], big_w[:, 1024:]  # the folded action_time_mlp_in halves use exactly such views
    y = ops.sgemm_f32(a, w)
    torch.cuda.synchronize()
    z = a.double() @ w.double().t()
    assert (y.double() - z).abs().max().item() < 5e-5


@pytest.mark.parametrize("M,N,K", [(400, 1536, 512), (400, 512, 1024), (160, 1024, 1024), (40, 512, 512)])
def test_three_term_bf16_split_gemm_is_fp32_accurate(M, N, K):
    """The tensor-core replacement of the fp32 SIMT GEMMs (verifier trajectory encoder, action_time_mlp_out): a_hi w_hi +
    a_hi w_lo + a_lo w_hi accumulated in fp32 by one bf16 tcgen05 GEMM over the 3K axis.  Error budget: measured against a
    float64 product it must stay within 4x the fp32 SIMT kernel's own error and below 3e-5 of the row/column norms -
    two orders of magnitude under the 1e-3 score tolerance."""
    from cover_vla_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda")
    a3, w3 = ops.split3_f32(a), ops.split3_f32(w, weight_layout=True)
    assert a3.shape == (M, 3 * K) and torch.equal(a3[:, :K], a.to(torch.bfloat16)) and torch.equal(a3[:, :K], a3[:, K:2 * K])
    assert torch.equal(w3[:, :K], w3[:, 2 * K:]) and torch.equal(w3[:, K:2 * K], (w - w3[:, :K].float()).to(torch.bfloat16))
    out = ops.gemm_bf16(a3, w3, epilogue=ops.EPI_F32, bias=bias)
    ref = (a.double() @ w.double().T + bias.double())
    scale = (a.double().norm(dim=1)[:, None] * w.double().norm(dim=1)[None, :])
    err_tc = ((out.double() - ref).abs() / scale).max().item()
    simt = ops.sgemm_f32(a, w, bias=bias)
    err_simt = ((simt.double() - ref).abs() / scale).max().item()
    print(f"{M}x{N}x{K}: 3-term bf16 split {err_tc:.2e} vs fp32 SIMT {err_simt:.2e} (relative to |a||w|)")
    assert err_tc < 3e-5 and err_tc < max(4 * err_simt, 2e-6)
    relu = ops.split3_f32(a, relu=True)
    assert torch.equal(relu[:, :K], a.clamp_min(0).to(torch.bfloat16))
