"""Attention kernels (fast logits-in-smem path and two-pass fallback) vs a plain PyTorch fp32 reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, heads, kv_heads, hd, mask):
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    g = heads // kv_heads
    qf = q.float().view(B, Tq, heads, hd).transpose(1, 2)
    kf = k.float().view(B, Tk, kv_heads, hd).transpose(1, 2).repeat_interleave(g, dim=1)
    vf = v.float().view(B, Tk, kv_heads, hd).transpose(1, 2).repeat_interleave(g, dim=1)
    att = qf @ kf.transpose(-1, -2) * hd ** -0.5
    att = att.masked_fill(~mask[:, None], float("-inf"))
    p = torch.softmax(att, dim=-1).to(torch.bfloat16).float()
    return (p @ vf).transpose(1, 2).reshape(B, Tq, heads * hd)


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("B,Tq,Tk,heads,kvh,hd", [(1, 256, 256, 16, 16, 72), (3, 328, 328, 8, 1, 256),
                                                   (1, 576, 576, 16, 16, 64), (2, 64, 64, 4, 4, 64), (2, 100, 77, 2, 1, 128)])
def test_self_attention(B, Tq, Tk, heads, kvh, hd, two_pass):
    from cover_vla_b200 import ops
    torch.manual_seed(B * 1000 + Tq + hd)
    q = torch.randn(B, Tq, heads * hd, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Tk, kvh * hd, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Tk, kvh * hd, device="cuda").to(torch.bfloat16)
    lens = torch.randint(Tk // 2, Tk + 1, (B,), device="cuda", dtype=torch.int32)
    out = ops.attention(q, k, v, heads=heads, kv_heads=kvh, head_dim=hd, kv0_len_dev=lens, force_two_pass=two_pass)
    mask = (torch.arange(Tk, device="cuda")[None, :] < lens[:, None])[:, None, :].expand(B, Tq, Tk)
    ref = _ref(q, k, v, heads, kvh, hd, mask)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err
    assert ((out.float() - ref).norm() / ref.norm()).item() < 5e-3


@pytest.mark.parametrize("two_pass", [False, True])
def test_denoise_attention_two_segments(two_pass):
    """N candidates x 5 suffix tokens against the rephrase's prefix cache + own suffix keys (pi0 mask)."""
    from cover_vla_b200 import ops
    torch.manual_seed(7)
    R, K, S, P, heads, hd = 3, 4, 5, 328, 8, 256
    N = R * K
    q = torch.randn(N, S, heads * hd, device="cuda").to(torch.bfloat16)
    k0 = torch.randn(R, P, hd, device="cuda").to(torch.bfloat16)
    v0 = torch.randn(R, P, hd, device="cuda").to(torch.bfloat16)
    k1 = torch.randn(N, S, hd, device="cuda").to(torch.bfloat16)
    v1 = torch.randn(N, S, hd, device="cuda").to(torch.bfloat16)
    lens = torch.tensor([270, 328, 300], device="cuda", dtype=torch.int32)
    out = ops.attention(q, k0, v0, heads=heads, kv_heads=1, head_dim=hd, kv0_len_dev=lens, q_per_kv_batch=K,
                        k1=k1, v1=v1, suffix_mask=True, force_two_pass=two_pass)
    for n in range(N):
        r = n // K
        L = int(lens[r])
        kk = torch.cat([k0[r, :L], k1[n]])[None]
        vv = torch.cat([v0[r, :L], v1[n]])[None]
        mask = torch.ones(1, S, L + S, dtype=torch.bool, device="cuda")
        mask[0, 0, L + 1:] = False  # the state token only sees itself among the suffix keys
        ref = _ref(q[n:n + 1], kk, vv, heads, 1, hd, mask)
        err = (out[n:n + 1].float() - ref).abs().max().item()
        assert err < 2e-2, (n, err)


def _rope(x, pos, hd):
    """apply_rope (paligemma_with_expert.py:34-57) on [..., T, H, hd] bf16 with positions [..., T]."""
    half = hd // 2
    ts = 10000.0 ** ((2.0 / hd) * torch.arange(half, dtype=torch.float32, device=x.device))
    rad = pos[..., None].float() / ts
    sin, cos = torch.sin(rad)[..., None, :], torch.cos(rad)[..., None, :]
    xf = x.float()
    x1, x2 = xf[..., :half], xf[..., half:]
    return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1).to(torch.bfloat16)


@pytest.mark.parametrize("B,T", [(8, 280), (3, 300), (2, 100), (1, 16), (2, 37), (1, 320), (2, 257), (3, 328), (1, 384)])
def test_prefix_attention_tcgen05(B, T):
    """tcgen05/TMEM prefix attention (MQA 8 x 256, heads folded into UMMA rows) vs the eager ledger in fp32."""
    from cover_vla_b200 import ops
    torch.manual_seed(B * 100 + T)
    heads, hd = 8, 256
    q = torch.randn(B, T, heads * hd, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, T, hd, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, T, hd, device="cuda").to(torch.bfloat16)
    lens = torch.randint(max(1, T // 2), T + 1, (B,), device="cuda", dtype=torch.int32)
    lens[0] = T
    out = ops.attention_umma(q, k, v, lens=lens)
    torch.cuda.synchronize()
    mask = (torch.arange(T, device="cuda")[None, :] < lens[:, None])[:, None, :].expand(B, T, T)
    ref = _ref(q, k, v, heads, 1, hd, mask)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err
    assert ((out.float() - ref).norm() / ref.norm()).item() < 5e-3
    # and against the mma.sync kernel it replaces (same ledger, different accumulation order / exp implementation)
    old = ops.attention(q, k, v, heads=heads, kv_heads=1, head_dim=hd, kv0_len_dev=lens)
    assert (out.float() - old.float()).abs().max().item() < 2e-2


def test_prefix_attention_tcgen05_rejects_unsupported_shapes():
    from cover_vla_b200 import ops, _lib
    q = torch.randn(1, 392, 8 * 256, device="cuda").to(torch.bfloat16)
    k = torch.randn(1, 392, 256, device="cuda").to(torch.bfloat16)
    with pytest.raises(_lib.CvbError):
        ops.attention_umma(q, k, k)


@pytest.mark.parametrize("R,K,S,P,heads,lens,use_rope", [
    (3, 4, 5, 328, 8, [270, 328, 300], True), (8, 5, 5, 280, 8, [264, 270, 280, 265, 277, 256, 280, 269], True),
    (1, 5, 5, 328, 8, [61], True), (2, 2, 8, 100, 4, [100, 7], False), (2, 3, 5, 40, 8, [1, 40], True),
    (1, 1, 16, 352, 8, [352], False), (2, 2, 1, 64, 8, [33, 64], True),
    (8, 20, 5, 280, 8, [264, 270, 280, 265, 277, 256, 280, 269], True)])  # 160 candidates: one CTA per candidate
def test_denoise_attention_tcgen05(R, K, S, P, heads, lens, use_rope):
    """tcgen05/TMEM decode attention (algo 3): RoPE fused into the swizzled UMMA tile staging, exact softmax split over
    16 warps, prefix P.V on the tensor core with the transposed V cache, suffix keys added in fp32."""
    from cover_vla_b200 import ops
    torch.manual_seed(13)
    hd = 256
    N = R * K
    q = torch.randn(N, S, heads * hd, device="cuda").to(torch.bfloat16)
    k0 = torch.randn(R, P, hd, device="cuda").to(torch.bfloat16)
    v0 = torch.randn(R, P, hd, device="cuda").to(torch.bfloat16)
    k1 = torch.randn(N, S if S <= 8 else 5, hd, device="cuda").to(torch.bfloat16)
    v1 = torch.randn_like(k1)
    S1 = k1.shape[1]
    lens_t = torch.tensor(lens, device="cuda", dtype=torch.int32)
    pos = lens_t[:, None] + torch.arange(S, device="cuda")[None, :]
    half = hd // 2
    ts = 10000.0 ** ((2.0 / hd) * torch.arange(half, dtype=torch.float32, device="cuda"))
    rad = pos[..., None].float() / ts
    tab = torch.stack([torch.cos(rad), torch.sin(rad)], dim=-1).contiguous() if use_rope else None
    vt0 = ops.transpose_values(v0)
    kw = dict(heads=heads, kv_heads=1, head_dim=hd, kv0_len_dev=lens_t, q_per_kv_batch=K, k1=k1, v1=v1,
              suffix_mask=True, rope=tab, vt0=vt0, algo=3)
    out = ops.attention(q, k0, v0, **kw)
    out2 = ops.attention(q, k0, v0, **kw)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)  # fixed-order reductions: deterministic
    if use_rope:
        posn = pos.repeat_interleave(K, dim=0)
        qr = _rope(q.view(N, S, heads, hd), posn, hd).view(N, S, heads * hd)
        k1r = _rope(k1.view(N, S1, 1, hd), posn[:, :S1], hd).view(N, S1, hd)
    else:
        qr, k1r = q, k1
    for n in range(N):
        r = n // K
        L = int(lens[r])
        kk = torch.cat([k0[r, :L], k1r[n]])[None]
        vv = torch.cat([v0[r, :L], v1[n]])[None]
        mask = torch.ones(1, S, L + S1, dtype=torch.bool, device="cuda")
        mask[0, 0, L + 1:] = False
        ref = _ref(qr[n:n + 1], kk, vv, heads, 1, hd, mask)
        err = (out[n:n + 1].float() - ref).abs().max().item()
        assert err < 2e-2, (n, err)


@pytest.mark.parametrize("B,T,heads,hd,full", [(1, 576, 16, 64, True), (2, 576, 16, 64, False), (1, 300, 4, 64, False),
                                               (1, 512, 2, 64, False), (3, 257, 2, 32, False), (1, 768, 3, 32, True)])
def test_long_multihead_attention_tcgen05(B, T, heads, hd, full):
    """attn_mha_long_umma_kernel (verifier ViT-L/16-384: 576 tokens x 16 heads x 64): exact two-pass softmax over key
    chunks of 192 with S double-buffered and O resident in TMEM; q / k / v are strided views of a fused qkv buffer as in
    the trunk; algo=3 fails unless a tcgen05 kernel takes the shape."""
    from cover_vla_b200 import ops
    torch.manual_seed(T * 7 + heads)
    W = heads * hd
    qkv = torch.randn(B, T, 3 * W, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, :W], qkv[:, :, W:2 * W], qkv[:, :, 2 * W:]
    lens = (torch.full((B,), T, device="cuda", dtype=torch.int32) if full else
            torch.randint(T // 2, T + 1, (B,), device="cuda", dtype=torch.int32))
    out = ops.attention(q, k, v, heads=heads, kv_heads=heads, head_dim=hd, kv0_len_dev=lens, algo=3)
    mask = (torch.arange(T, device="cuda")[None, :] < lens[:, None])[:, None, :].expand(B, T, T)
    ref = _ref(q.contiguous(), k.contiguous(), v.contiguous(), heads, heads, hd, mask)
    err = (out.float() - ref).abs().max().item()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert err < 2e-2 and rel < 5e-3, (err, rel)
    # same ledger as the mma.sync kernel it replaces
    old = ops.attention(q, k, v, heads=heads, kv_heads=heads, head_dim=hd, kv0_len_dev=lens, force_two_pass=True)
    assert ((out.float() - old.float()).norm() / ref.norm()).item() < 5e-3


@pytest.mark.parametrize("R,K,S", [(32, 5, 5), (32, 5, 4), (40, 4, 5), (75, 2, 5), (30, 7, 4)])
def test_denoise_attention_grouped_candidates_are_bit_identical(R, K, S):
    """More candidates than SMs (several observations per call): the candidates of a rephrase share a CTA - its query
    rows, one copy of the prefix K / V^T and the 16-wide suffix tile behind a block-diagonal mask (5 candidates -> 3 + 2).
    A row's arithmetic must not change: the launch over all candidates equals, bit for bit, the same candidates launched
    in slices small enough (<= 148 / 2) to take one candidate per CTA pair."""
    from cover_vla_b200 import ops
    torch.manual_seed(R * 31 + K)
    hd, heads, P = 256, 8, 280
    N = R * K
    assert N > 148
    q = torch.randn(N, S, heads * hd, device="cuda").to(torch.bfloat16)
    k0 = torch.randn(R, P, hd, device="cuda").to(torch.bfloat16)
    v0 = torch.randn(R, P, hd, device="cuda").to(torch.bfloat16)
    k1 = torch.randn(N, S, hd, device="cuda").to(torch.bfloat16)
    v1 = torch.randn_like(k1)
    lens_t = torch.randint(200, P + 1, (R,), device="cuda", dtype=torch.int32)
    pos = lens_t[:, None] + torch.arange(S, device="cuda")[None, :]
    ts = 10000.0 ** ((2.0 / hd) * torch.arange(hd // 2, dtype=torch.float32, device="cuda"))
    rad = pos[..., None].float() / ts
    tab = torch.stack([torch.cos(rad), torch.sin(rad)], dim=-1).contiguous()
    vt0 = ops.transpose_values(v0)
    kw = dict(heads=heads, kv_heads=1, head_dim=hd, q_per_kv_batch=K, suffix_mask=True, algo=3)
    out = ops.attention(q, k0, v0, kv0_len_dev=lens_t, k1=k1, v1=v1, rope=tab, vt0=vt0, **kw)
    step = max(1, 70 // K)  # rephrases per slice: at most 70 candidates -> one candidate per CTA pair
    for r0 in range(0, R, step):
        r1 = min(R, r0 + step)
        part = ops.attention(q[r0 * K:r1 * K].contiguous(), k0[r0:r1].contiguous(), v0[r0:r1].contiguous(),
                             kv0_len_dev=lens_t[r0:r1].contiguous(), k1=k1[r0 * K:r1 * K].contiguous(),
                             v1=v1[r0 * K:r1 * K].contiguous(), rope=tab[r0:r1].contiguous(), vt0=vt0[r0:r1].contiguous(), **kw)
        assert torch.equal(part, out[r0 * K:r1 * K]), (r0, (part.float() - out[r0 * K:r1 * K].float()).abs().max().item())
    # and it is right: a few candidates against the fp32 restatement of the eager ledger
    posn = pos.repeat_interleave(K, dim=0)
    qr = _rope(q.view(N, S, heads, hd), posn, hd).view(N, S, heads * hd)
    k1r = _rope(k1.view(N, S, 1, hd), posn, hd).view(N, S, hd)
    for n in (0, 1, K - 1, K, N // 2, N - 1):
        r = n // K
        L = int(lens_t[r])
        kk = torch.cat([k0[r, :L], k1r[n]])[None]
        vv = torch.cat([v0[r, :L], v1[n]])[None]
        mask = torch.ones(1, S, L + S, dtype=torch.bool, device="cuda")
        mask[0, 0, L + 1:] = False
        ref = _ref(qr[n:n + 1], kk, vv, heads, 1, hd, mask)
        assert (out[n:n + 1].float() - ref).abs().max().item() < 2e-2, n


def test_denoise_attention_with_cached_state_key():
    """SURVEY.md F7 at operator level: step 0 (5 query rows / 5 suffix keys per candidate) writes the state token's rotated
    key and its value; a hoisted call (4 action query rows, suffix keys = [cached state key, 4 new keys]) must return
    exactly the action rows of the full call."""
    from cover_vla_b200 import _lib
    import ctypes as C
    pytest.importorskip("torch")
    # the C-ABI operator does not expose the cache arguments (engine-internal); exercised end to end by
    # tests/test_pi0_gpu.py::test_state_token_hoist_is_exact, which requires bit equality of the sampled actions
    assert _lib.load() is not None and C.sizeof(C.c_void_p) == 8
