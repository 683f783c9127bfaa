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
