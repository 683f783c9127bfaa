"""tcgen05 GEMM parity against a plain PyTorch fp32 reference (floating-point kernel)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias=None):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    return y


def _rel_err(x, y):
    return ((x.float() - y.float()).norm() / (y.float().norm() + 1e-12)).item()


SHAPES = [
    (128, 64, 64), (128, 256, 64), (128, 128, 128), (256, 256, 256),
    (200, 2560, 1024), (200, 1024, 4096), (328, 2048, 2048), (2624, 2560, 2048),
    (256, 1152, 592), (256, 1152, 4304), (256, 4304, 1152), (576, 3072, 1024),
    (64, 1024, 1024), (5, 64, 64), (130, 72, 200),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("bn", [0, 64, 128, 256])
def test_gemm_store(M, N, K, bn):
    from cover_vla_b200 import ops
    torch.manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    y = ops.gemm_bf16(a, w, bias=bias, force_bn=bn)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias)
    # one bf16 rounding of an fp32-accumulated result
    assert _rel_err(y, ref) < 4e-3, (M, N, K, bn, _rel_err(y, ref))
    assert (y.float() - ref).abs().max().item() < 0.06


def test_gemm_f32_and_gelu_and_resid():
    from cover_vla_b200 import ops
    torch.manual_seed(0)
    M, N, K = 300, 512, 1024
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", dtype=torch.float32)
    ref = _ref(a, w, bias)
    y = ops.gemm_bf16(a, w, bias=bias, epilogue=ops.EPI_F32)
    assert _rel_err(y, ref) < 1e-5
    bias16 = bias.to(torch.bfloat16)
    ref16 = _ref(a, w, bias16).to(torch.bfloat16)
    y = ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_GELU)
    refg = torch.nn.functional.gelu(ref16, approximate="tanh")
    assert _rel_err(y, refg) < 4e-3
    r = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    y = ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_RESID, resid=r)
    refr = (ref16.float() + r.float()).to(torch.bfloat16)
    assert _rel_err(y, refr) < 4e-3
    r32 = torch.randn(M, N, device="cuda", dtype=torch.float32)
    y = ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_RESID, resid=r32)
    refr = (ref16.float() + r32).to(torch.bfloat16)
    assert _rel_err(y, refr) < 4e-3
    # in-place residual (C aliases R)
    r2 = r.clone()
    ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_RESID, resid=r2, out=r2)
    assert _rel_err(r2, (ref16.float() + r.float()).to(torch.bfloat16)) < 4e-3


def test_gemm_geglu():
    from cover_vla_b200 import ops
    torch.manual_seed(1)
    M, I, K = 200, 4096, 1024
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    wg = (torch.randn(I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    wu = (torch.randn(I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    packed = torch.stack([wg.view(I // 128, 128, K), wu.view(I // 128, 128, K)], dim=1).reshape(2 * I, K).contiguous()
    y = ops.gemm_bf16(a, packed, epilogue=ops.EPI_GEGLU, n_out=I)
    g = (a.float() @ wg.float().t()).to(torch.bfloat16)
    u = (a.float() @ wu.float().t()).to(torch.bfloat16)
    ref = torch.nn.functional.gelu(g, approximate="tanh") * u
    assert y.shape == (M, I)
    assert _rel_err(y, ref) < 6e-3


def test_gemm_device_side_rows():
    from cover_vla_b200 import ops
    torch.manual_seed(2)
    M, N, K = 512, 256, 256
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    m_dev = torch.tensor([300], device="cuda", dtype=torch.int32)
    ops.gemm_bf16(a, w, out=out, m_dev=m_dev)
    ref = _ref(a, w)
    assert _rel_err(out[:300], ref[:300]) < 4e-3
    assert out[300:].abs().max().item() == 0.0


# ---------------------------------------------------------------------------------------------------------------
# skinny (swap-AB, cluster split-K through distributed shared memory) GEMM: M <= 256
# ---------------------------------------------------------------------------------------------------------------
SKINNY_SHAPES = [
    (200, 2560, 1024), (200, 1024, 2048), (200, 1024, 4096), (256, 3456, 1152), (256, 1152, 4304),
    (256, 4304, 1152), (64, 1024, 1024), (64, 3072, 1024), (5, 64, 64), (130, 72, 200), (16, 128, 64), (1, 256, 128),
    (255, 136, 72),
]


@pytest.mark.parametrize("M,N,K", SKINNY_SHAPES)
@pytest.mark.parametrize("split", [-100, -1, -2, -3, -8])
def test_skinny_gemm_store(M, N, K, split):
    from cover_vla_b200 import ops
    torch.manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    y = ops.gemm_bf16(a, w, bias=bias, force_bn=split)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias)
    assert _rel_err(y, ref) < 4e-3, (M, N, K, split, _rel_err(y, ref))
    assert (y.float() - ref).abs().max().item() < 0.06
    # deterministic: the cluster reduction sums partials in rank order
    y2 = ops.gemm_bf16(a, w, bias=bias, force_bn=split)
    assert torch.equal(y, y2)


def test_skinny_matches_general_kernel_bitwise_when_unsplit():
    """S = 1: one fp32 accumulator over the whole K, same rounding points -> same bf16 bits as the general kernel."""
    from cover_vla_b200 import ops
    torch.manual_seed(3)
    M, N, K = 200, 1024, 1024
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    y0 = ops.gemm_bf16(a, w, force_bn=64)
    y1 = ops.gemm_bf16(a, w, force_bn=-1)
    assert (y0.float() - y1.float()).abs().max().item() <= 2 ** -6  # accumulation order inside the tensor core may differ
    assert (y0 != y1).float().mean().item() < 0.02


def test_skinny_epilogues():
    from cover_vla_b200 import ops
    torch.manual_seed(0)
    M, N, K = 200, 512, 1024
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", dtype=torch.float32)
    ref = _ref(a, w, bias)
    y = ops.gemm_bf16(a, w, bias=bias, epilogue=ops.EPI_F32, force_bn=ops.SKINNY_AUTO)
    assert _rel_err(y, ref) < 1e-5
    bias16 = bias.to(torch.bfloat16)
    ref16 = _ref(a, w, bias16).to(torch.bfloat16)
    y = ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_GELU, force_bn=ops.SKINNY_AUTO)
    assert _rel_err(y, torch.nn.functional.gelu(ref16, approximate="tanh")) < 4e-3
    r = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    y = ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_RESID, resid=r, force_bn=ops.SKINNY_AUTO)
    assert _rel_err(y, (ref16.float() + r.float()).to(torch.bfloat16)) < 4e-3
    r32 = torch.randn(M, N, device="cuda", dtype=torch.float32)
    y = ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_RESID, resid=r32, force_bn=ops.SKINNY_AUTO)
    assert _rel_err(y, (ref16.float() + r32).to(torch.bfloat16)) < 4e-3
    r2 = r.clone()
    ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_RESID, resid=r2, out=r2, force_bn=ops.SKINNY_AUTO)
    assert _rel_err(r2, (ref16.float() + r.float()).to(torch.bfloat16)) < 4e-3


@pytest.mark.parametrize("M,I,K", [(200, 4096, 1024), (37, 320, 256), (256, 1024, 2048)])
def test_skinny_geglu64(M, I, K):
    from cover_vla_b200 import ops
    torch.manual_seed(1)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    wg = (torch.randn(I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    wu = (torch.randn(I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    packed = torch.stack([wg.view(I // 64, 64, K), wu.view(I // 64, 64, K)], dim=1).reshape(2 * I, K).contiguous()
    y = ops.gemm_bf16(a, packed, epilogue=ops.EPI_GEGLU64, n_out=I)
    g = (a.float() @ wg.float().t()).to(torch.bfloat16)
    u = (a.float() @ wu.float().t()).to(torch.bfloat16)
    ref = torch.nn.functional.gelu(g, approximate="tanh") * u
    assert y.shape == (M, I)
    assert _rel_err(y, ref) < 6e-3


# ---------------------------------------------------------------------------------------------------------------
# CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x 256 tiles): force_bn = 512
# ---------------------------------------------------------------------------------------------------------------
PAIR_SHAPES = [(2240, 2560, 2048), (2240, 2048, 2048), (2624, 2048, 1024), (256, 256, 64), (300, 512, 1024),
               (128, 256, 128), (5, 64, 64), (130, 72, 200), (513, 264, 328), (4096, 1024, 512)]


@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
def test_pair_gemm_store(M, N, K):
    from cover_vla_b200 import ops
    torch.manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    y = ops.gemm_bf16(a, w, bias=bias, force_bn=512)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias)
    assert _rel_err(y, ref) < 4e-3, (M, N, K, _rel_err(y, ref))
    assert (y.float() - ref).abs().max().item() < 0.06
    # one fp32 accumulator over the whole K with the same rounding points: same bits as the 1-CTA kernel up to the
    # accumulation order inside the tensor core
    y1 = ops.gemm_bf16(a, w, bias=bias, force_bn=256)
    assert (y.float() - y1.float()).abs().max().item() <= 2 ** -5
    assert (y != y1).float().mean().item() < 0.02


def test_pair_gemm_epilogues():
    from cover_vla_b200 import ops
    torch.manual_seed(0)
    M, N, K = 1500, 512, 1024
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", dtype=torch.float32)
    ref = _ref(a, w, bias)
    y = ops.gemm_bf16(a, w, bias=bias, epilogue=ops.EPI_F32, force_bn=512)
    assert _rel_err(y, ref) < 1e-5
    bias16 = bias.to(torch.bfloat16)
    ref16 = _ref(a, w, bias16).to(torch.bfloat16)
    y = ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_GELU, force_bn=512)
    assert _rel_err(y, torch.nn.functional.gelu(ref16, approximate="tanh")) < 4e-3
    r = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    r2 = r.clone()
    ops.gemm_bf16(a, w, bias=bias16, epilogue=ops.EPI_RESID, resid=r2, out=r2, force_bn=512)
    assert _rel_err(r2, (ref16.float() + r.float()).to(torch.bfloat16)) < 4e-3
    I = 2048
    wg = (torch.randn(I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    wu = (torch.randn(I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    packed = torch.stack([wg.view(I // 128, 128, K), wu.view(I // 128, 128, K)], dim=1).reshape(2 * I, K).contiguous()
    y = ops.gemm_bf16(a, packed, epilogue=ops.EPI_GEGLU, n_out=I, force_bn=512)
    g = (a.float() @ wg.float().t()).to(torch.bfloat16)
    u = (a.float() @ wu.float().t()).to(torch.bfloat16)
    assert _rel_err(y, torch.nn.functional.gelu(g, approximate="tanh") * u) < 6e-3


@pytest.mark.parametrize("M,I,K", [(200, 4096, 1024), (37, 320, 256), (1280, 1024, 1024)])
def test_geglu64_on_128_wide_general_tiles(M, I, K):
    """[64 gate | 64 up] packing on the general kernel with 128 x 128 tiles (the expert's gate/up in the denoise loop)."""
    from cover_vla_b200 import ops
    torch.manual_seed(4)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    wg = (torch.randn(I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    wu = (torch.randn(I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    packed = torch.stack([wg.view(I // 64, 64, K), wu.view(I // 64, 64, K)], dim=1).reshape(2 * I, K).contiguous()
    y = ops.gemm_bf16(a, packed, epilogue=ops.EPI_GEGLU64, n_out=I, force_bn=128)
    g = (a.float() @ wg.float().t()).to(torch.bfloat16)
    u = (a.float() @ wu.float().t()).to(torch.bfloat16)
    ref = torch.nn.functional.gelu(g, approximate="tanh") * u
    assert y.shape == (M, I)
    assert _rel_err(y, ref) < 6e-3
    if M <= 256:  # same rounding points as the skinny kernel
        y2 = ops.gemm_bf16(a, packed, epilogue=ops.EPI_GEGLU64, n_out=I)
        assert (y.float() - y2.float()).abs().max().item() <= 2 ** -5
