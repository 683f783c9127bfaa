"""TEST INFRASTRUCTURE (oracle) - CPU restatement of the reference pi0 sampling path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product (cover_vla_b200/) never does and has no CPU fallback.

Restates, op for op and with the same dtype/rounding ledger:
  PI0FlowMatching.sample_actions      lerobot_custom/lerobot/common/policies/pi0/modeling_pi0.py:672-715
  PI0FlowMatching.embed_prefix        modeling_pi0.py:517-567
  PI0FlowMatching.embed_suffix        modeling_pi0.py:569-629
  PI0FlowMatching.denoise_step        modeling_pi0.py:717-752
  create_sinusoidal_pos_embedding     modeling_pi0.py:71-89
  make_att_2d_masks                   modeling_pi0.py:98-128
  PaliGemmaWithExpertModel.forward    pi0/paligemma_with_expert.py:236-360
  apply_rope / eager_attention        paligemma_with_expert.py:34-57, :376-434
Third-party arithmetic the reference reaches (not vendored under /root/reference; pinned
transformers==4.48.3, CoVer_VLA/scripts/env_simpler_pi.sh:83) is restated from its published
definition: GemmaRMSNorm ((1+w), fp32 statistics), GemmaMLP (down(gelu_tanh(gate)*up)),
SiglipVisionModel (pre-LN blocks, learned position embedding, gelu_tanh MLP, SDPA attention, post
layernorm, no head), PaliGemmaMultiModalProjector + get_image_features (/sqrt(hidden)).

PINNING: the reference ships no golden vectors or tests for this path.  The
oracle is pinned against outputs of the reference's own files executed in the authoring container
through oracle/ref_shim.py: tests/golden/pi0_*.pt (made by oracle/make_golden.py) and, when
/root/reference is present, live in tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from cover_vla_b200.synthetic import (EX, FULL, LM, MID, MM, PW, TINY, VT, PI0Dims, canonical_key, expand_to_batch,  # noqa: F401
                                      make_inputs, make_pi0_weights, to_hf5_key, weight_specs)


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
# Activation dtype of the graph.  bfloat16 = the reference's ledger.  `truth_mode()` switches it
# to float32: the SAME graph with every bf16 rounding point removed (weights keep their bf16 VALUES, upcast) - the
# "fp32 truth" SURVEY.md F10 asks for, used only to arbitrate err(ours, truth) against err(reference_bf16, truth).
ACT = torch.bfloat16


class truth_mode:
    def __enter__(self):
        global ACT
        self._old, ACT = ACT, torch.float32
        return self

    def __exit__(self, *exc):
        global ACT
        ACT = self._old


def truth_weights(w):
    """fp32 copies of the (bf16-valued) reference weights for truth_mode()."""
    return {k: v.float() for k, v in w.items()}


def gemma_rmsnorm(x, w, eps=1e-6):
    # transformers GemmaRMSNorm: _norm(x.float()) * (1 + w.float()), cast back to x.dtype
    xf = x.float()
    out = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    out = out * (1.0 + w.float())
    return out.type_as(x)


def gemma_mlp(x, wg, wu, wd):
    return F.linear(F.gelu(F.linear(x, wg), approximate="tanh") * F.linear(x, wu), wd)


def apply_rope(x, positions, max_wavelength=10_000):
    # paligemma_with_expert.py:34-57
    d_half = x.shape[-1] // 2
    dtype = x.dtype
    x = x.to(torch.float32)
    freq_exponents = (2.0 / x.shape[-1]) * torch.arange(d_half, dtype=torch.float32)
    timescale = max_wavelength ** freq_exponents
    radians = positions[..., None].to(torch.float32) / timescale[None, None, :].to(torch.float32)
    radians = radians[..., None, :]
    sin, cos = torch.sin(radians), torch.cos(radians)
    x1, x2 = x.split(d_half, dim=-1)
    res = torch.empty_like(x)
    res[..., :d_half] = x1 * cos - x2 * sin
    res[..., d_half:] = x2 * cos + x1 * sin
    return res.to(dtype)


def eager_attention(mask, q, k, v, heads, head_dim):
    # paligemma_with_expert.py:376-434 ; q [B,Lq,H,D] bf16, k/v [B,Lk,1,D] bf16, mask [B,Lq,Lk] bool
    B, Lk = k.shape[0], k.shape[1]
    k = k[:, :, :, None, :].expand(B, Lk, 1, heads, head_dim).reshape(B, Lk, heads, head_dim)
    v = v[:, :, :, None, :].expand(B, Lk, 1, heads, head_dim).reshape(B, Lk, heads, head_dim)
    qf = q.to(torch.float32).transpose(1, 2)
    kf = k.to(torch.float32).transpose(1, 2)
    att = torch.matmul(qf, kf.transpose(2, 3))
    att *= head_dim ** -0.5
    big_neg = -2.3819763e38
    att = torch.where(mask[:, None, :, :], att, big_neg)
    probs = F.softmax(att, dim=-1).to(dtype=v.dtype)
    out = torch.matmul(probs, v.permute(0, 2, 1, 3))
    out = out.permute(0, 2, 1, 3)
    return out.reshape(B, -1, heads * head_dim)


def make_att_2d_masks(pad_masks, att_masks):
    cumsum = torch.cumsum(att_masks, dim=1)
    att_2d = cumsum[:, None, :] <= cumsum[:, :, None]
    pad_2d = pad_masks[:, None, :] * pad_masks[:, :, None]
    return att_2d & pad_2d


def sinusoidal_time_embedding(time, dimension, min_period=4e-3, max_period=4.0):
    # modeling_pi0.py:71-89 (float64 math)
    fraction = torch.linspace(0.0, 1.0, dimension // 2, dtype=torch.float64)
    period = min_period * (max_period / min_period) ** fraction
    scaling = 1.0 / period * 2 * math.pi
    sin_input = scaling[None, :] * time[:, None]
    return torch.cat([torch.sin(sin_input), torch.cos(sin_input)], dim=1)


def denoise_times(num_steps: int):
    """The fp32 time values the reference loop visits (modeling_pi0.py:697-714)."""
    dt = torch.tensor(-1.0 / num_steps, dtype=torch.float32)
    time = torch.tensor(1.0, dtype=torch.float32)
    out = []
    while time >= -dt / 2:
        out.append(float(time))
        time = time + dt
    return out, float(dt)


# ------------------------------------------------------------------------------------------------
# SigLIP tower + projector  (transformers SiglipVisionModel, PaliGemma get_image_features @4.48.3)
# ------------------------------------------------------------------------------------------------
def siglip_tower(w, d: PI0Dims, pixel_values):
    x = pixel_values.to(ACT)
    pe = F.conv2d(x, w[VT + "embeddings.patch_embedding.weight"], w[VT + "embeddings.patch_embedding.bias"],
                  stride=d.vis_patch)
    h = pe.flatten(2).transpose(1, 2)
    h = h + w[VT + "embeddings.position_embedding.weight"][None]
    B, T, W = h.shape
    hd = d.vis_width // d.vis_heads
    for l in range(d.vis_layers):
        p = VT + f"encoder.layers.{l}."
        r = h
        y = F.layer_norm(h, (W,), w[p + "layer_norm1.weight"], w[p + "layer_norm1.bias"], 1e-6)
        q = F.linear(y, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"])
        k = F.linear(y, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"])
        v = F.linear(y, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"])
        q, k, v = (t.view(B, T, d.vis_heads, hd).transpose(1, 2) for t in (q, k, v))
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False,
                                           scale=hd ** -0.5)
        a = a.transpose(1, 2).reshape(B, T, W).contiguous()
        a = F.linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
        h = r + a
        r = h
        y = F.layer_norm(h, (W,), w[p + "layer_norm2.weight"], w[p + "layer_norm2.bias"], 1e-6)
        y = F.linear(y, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])
        y = F.gelu(y, approximate="tanh")
        y = F.linear(y, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
        h = r + y
    return F.layer_norm(h, (W,), w[VT + "post_layernorm.weight"], w[VT + "post_layernorm.bias"], 1e-6)


def embed_image(w, d: PI0Dims, pixel_values):
    feats = siglip_tower(w, d, pixel_values)
    feats = F.linear(feats, w[MM + "weight"], w[MM + "bias"])
    return feats / (d.lm_width ** 0.5)


def embed_prefix(w, d: PI0Dims, image, lang_tokens, lang_masks, img_masks=None):
    # modeling_pi0.py:517-567.  `image`: one camera tensor [B,3,H,W] or a list of them (camera order = prompt order);
    # `img_masks`: per camera bool [B] (None = all present); masked cameras keep their tokens but are padding (:541-544)
    images = list(image) if isinstance(image, (list, tuple)) else [image]
    embs, pads = [], []
    for cam, im in enumerate(images):
        img_emb = embed_image(w, d, im).to(ACT)
        img_emb = img_emb * torch.tensor(img_emb.shape[-1] ** 0.5, dtype=img_emb.dtype)
        B, n_img = img_emb.shape[:2]
        m = torch.ones(B, dtype=torch.bool) if img_masks is None else img_masks[cam]
        embs.append(img_emb)
        pads.append(m[:, None].expand(B, n_img))
    lang_emb = F.embedding(lang_tokens, w[LM + "embed_tokens.weight"])
    lang_emb = lang_emb * math.sqrt(lang_emb.shape[-1])
    embs = torch.cat(embs + [lang_emb], dim=1)
    pad = torch.cat(pads + [lang_masks], dim=1)
    att = torch.zeros(B, pad.shape[1], dtype=torch.bool)
    return embs, pad, att


# ------------------------------------------------------------------------------------------------
# dual-tower layer loop  (paligemma_with_expert.py:236-360)
# ------------------------------------------------------------------------------------------------
def _tower_forward(w, d: PI0Dims, prefix: str, x, mask, pos, cache, fill, last_layer_kv_only=False):
    """One tower alone (the reference only ever runs one at a time on the sampling path)."""
    B = x.shape[0]
    new_cache = {}
    for l in range(d.layers):
        p = prefix + f"layers.{l}."
        y = gemma_rmsnorm(x, w[p + "input_layernorm.weight"])
        y = y.to(ACT)
        shp = (*y.shape[:-1], -1, d.head_dim)
        q = F.linear(y, w[p + "self_attn.q_proj.weight"]).view(shp)
        k = F.linear(y, w[p + "self_attn.k_proj.weight"]).view(shp)
        v = F.linear(y, w[p + "self_attn.v_proj.weight"]).view(shp)
        q = apply_rope(q, pos)
        k = apply_rope(k, pos)
        if fill:
            new_cache[l] = {"key_states": k, "value_states": v}
            if last_layer_kv_only and l == d.layers - 1:
                break
        else:
            k = torch.cat([cache[l]["key_states"], k], dim=1)
            v = torch.cat([cache[l]["value_states"], v], dim=1)
        a = eager_attention(mask, q, k, v, d.heads, d.head_dim).to(ACT)
        out = F.linear(a, w[p + "self_attn.o_proj.weight"])
        out += x  # in place: the result keeps out's dtype (bf16) even when x is fp32 (layer 0 of the suffix)
        res = out.clone()
        out = gemma_rmsnorm(out, w[p + "post_attention_layernorm.weight"])
        out = gemma_mlp(out, w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"], w[p + "mlp.down_proj.weight"])
        out += res
        x = out
    return x, new_cache


def embed_suffix(w, d: PI0Dims, state, x_t, timestep):
    # modeling_pi0.py:569-629
    state_emb = F.linear(state, w["state_proj.weight"], w["state_proj.bias"]).to(ACT)
    time_emb = sinusoidal_time_embedding(timestep, d.ex_width).type(dtype=ACT)
    action_emb = F.linear(x_t, w["action_in_proj.weight"], w["action_in_proj.bias"])
    time_emb = time_emb[:, None, :].expand_as(action_emb)
    at = torch.cat([action_emb, time_emb], dim=2)
    at = F.linear(at, w["action_time_mlp_in.weight"], w["action_time_mlp_in.bias"])
    at = F.silu(at)
    at = F.linear(at, w["action_time_mlp_out.weight"], w["action_time_mlp_out.bias"])
    embs = torch.cat([state_emb[:, None, :], at], dim=1)  # promotes to fp32
    B = state.shape[0]
    pad = torch.ones(B, 1 + d.chunk_size, dtype=torch.bool)
    att = torch.tensor([1, 1] + [0] * (d.chunk_size - 1), dtype=embs.dtype)[None].expand(B, -1)
    return embs, pad, att


def denoise_step(w, d: PI0Dims, state, prefix_pad, cache, x_t, timestep):
    # modeling_pi0.py:717-752
    suf, suf_pad, suf_att = embed_suffix(w, d, state, x_t, timestep)
    B, S = suf_pad.shape
    P = prefix_pad.shape[1]
    prefix_2d = prefix_pad[:, None, :].expand(B, S, P)
    suf_2d = make_att_2d_masks(suf_pad, suf_att)
    full = torch.cat([prefix_2d, suf_2d], dim=2)
    offsets = torch.sum(prefix_pad, dim=-1)[:, None]
    pos = offsets + torch.cumsum(suf_pad, dim=1) - 1
    out, _ = _tower_forward(w, d, EX, suf, full, pos, cache, fill=False)
    out = gemma_rmsnorm(out, w[EX + "norm.weight"])
    out = out[:, -d.chunk_size:].to(torch.float32)
    return F.linear(out, w["action_out_proj.weight"], w["action_out_proj.bias"])


def prefix_cache(w, d: PI0Dims, image, lang_tokens, lang_masks, last_layer_kv_only=False, img_masks=None):
    embs, pad, att = embed_prefix(w, d, image, lang_tokens, lang_masks, img_masks)
    mask = make_att_2d_masks(pad, att)
    pos = torch.cumsum(pad, dim=1) - 1
    _, cache = _tower_forward(w, d, LM, embs, mask, pos, None, fill=True, last_layer_kv_only=last_layer_kv_only)
    return cache, pad


@torch.no_grad()
def sample_actions(w, d: PI0Dims, image, lang_tokens, lang_masks, state, noise, trace=None, img_masks=None):
    """Exactly the reference batch layout: every argument has leading dim B (= N candidates); `image` may be a list of
    camera tensors with `img_masks` (see embed_prefix)."""
    cache, pad = prefix_cache(w, d, image, lang_tokens, lang_masks, img_masks=img_masks)
    if trace is not None:
        trace["k0"] = cache[0]["key_states"].clone()
        trace["v_last"] = cache[d.layers - 1]["value_states"].clone()
    B = state.shape[0]
    dt = torch.tensor(-1.0 / d.num_steps, dtype=torch.float32)
    x_t = noise.clone()
    time = torch.tensor(1.0, dtype=torch.float32)
    step = 0
    while time >= -dt / 2:
        v_t = denoise_step(w, d, state, pad, cache, x_t, time.expand(B))
        if trace is not None and step == 0:
            trace["v0"] = v_t.clone()
        x_t += dt * v_t
        time += dt
        step += 1
    return x_t


@torch.no_grad()
def sample_actions_dedup(w, d: PI0Dims, image1, lang_tokens_r, lang_masks_r, state1, noise, K):
    """Same arithmetic, de-duplicated the way the CUDA path schedules it: vision
    tower once, prefix once per rephrase (padded tokens kept), K samples share their rephrase's cache."""
    R = lang_tokens_r.shape[0]
    if isinstance(image1, (list, tuple)):  # several cameras
        img = [im.expand(R, -1, -1, -1) for im in image1]
    else:
        img = image1.expand(R, -1, -1, -1)
    cache, pad = prefix_cache(w, d, img, lang_tokens_r, lang_masks_r, last_layer_kv_only=True)
    rep = torch.arange(R).repeat_interleave(K)
    cache = {l: {k: v[rep] for k, v in c.items()} for l, c in cache.items()}
    pad = pad[rep]
    N = R * K
    state = state1.expand(N, -1)
    dt = torch.tensor(-1.0 / d.num_steps, dtype=torch.float32)
    x_t = noise.clone()
    time = torch.tensor(1.0, dtype=torch.float32)
    while time >= -dt / 2:
        v_t = denoise_step(w, d, state, pad, cache, x_t, time.expand(N))
        x_t += dt * v_t
        time += dt
    return x_t


