"""TEST INFRASTRUCTURE (oracle) - CPU restatement of the CoVer bridge_verifier scoring path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product never does.

Heads, fusion and selection restate reference code that IS in /root/reference:
  TextAwareVisualExtraction.forward      bridge_verifier/ensemble_eval/model.py:58-73
  CrossAttentionBlock / AttentionPooling model.py:7-38, 76-112
  get_embeddings_from_model_batch        ensemble_eval/efficient_ensemble_merged.py:194-247
  compute_max_similarity_scores_batch    efficient_ensemble_merged.py:309-454 (padding :379-390, fuse
                                         :404-411, scores :414, group-mean / argmax :417-447)
  VLA_SigLIP2_Bridge.extract_features    ensemble_eval/finetune_trajectory_bridge_ddp.py:297-355
and are pinned against that code run through oracle/ref_shim.py (tests/golden/verifier_*.pt and, when
/root/reference is present, tests/test_oracle_vs_reference.py).

The SigLIP2 trunk is NOT in /root/reference: it lives in the un-vendored, unpinned third-party
packages open_clip_torch and timm (bridge_verifier/setup.py:22-23; model
"hf-hub:timm/ViT-L-16-SigLIP2-384", efficient_ensemble_merged.py:42).  `trunk_*` below restates their
published architecture (timm VisionTransformer: patch16, learned pos-emb, pre-norm blocks, fused qkv
with bias, LN eps 1e-6, GELU-tanh MLP, no class token; open_clip TextTransformer: token + positional
embedding, pre-norm ResidualAttentionBlocks around nn.MultiheadAttention, no attention mask, ln_final,
Linear text_projection with bias) up to the two hook points the reference reads
(finetune_trajectory_bridge_ddp.py:271-278).  TRUNK PARITY IS UNPINNED upstream: no reference test,
golden vector or source pins it, and with random-init weights this restatement DEFINES the trunk.
"""
from __future__ import annotations

import math

import numpy as np  # noqa: F401
import torch
import torch.nn.functional as F

from cover_vla_b200.synthetic import (TR, VFULL, VMID, VMID_MLP, VTINY, VTINY_MLP, VerifierDims, head_specs, make_verifier_weights,  # noqa: F401
                                      pad_histories, sincos_position_embedding, trunk_specs)
from cover_vla_b200.synthetic import make_verifier_inputs as make_inputs  # noqa: F401


# ------------------------------------------------------------------------------------------------
# trunk (restated third-party architecture; see module docstring)
# ------------------------------------------------------------------------------------------------
# Activation dtype of the trunk: bfloat16 = what the reference runs (efficient_ensemble_merged.py:66 casts the trunk).
# truth_mode() + truth_weights() evaluate the SAME trunk in fp32 (weights keep their bf16 values): the arbiter for the
# score gate, like pi0_oracle.truth_mode - the heads are fp32 on both sides already.
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
    return {k: (v.float() if torch.is_tensor(v) and v.dtype == torch.bfloat16 else v) for k, v in w.items()}


def _mha_self(x, w_in, b_in, w_out, b_out, heads):
    B, T, Wd = x.shape
    hd = Wd // heads
    qkv = F.linear(x, w_in, b_in)
    q, k, v = qkv.split(Wd, dim=-1)
    q, k, v = (t.view(B, T, heads, hd).transpose(1, 2) for t in (q, k, v))
    a = F.scaled_dot_product_attention(q, k, v)
    a = a.transpose(1, 2).reshape(B, T, Wd)
    return F.linear(a, w_out, b_out)


def trunk_image_patches(w, d: VerifierDims, image):
    """image [B,3,H,W] -> output of visual.trunk.blocks[-1].attn (the hook at ddp.py:272-274), bf16."""
    v = TR + "visual.trunk."
    x = image.to(ACT)
    x = F.conv2d(x, w[v + "patch_embed.proj.weight"], w[v + "patch_embed.proj.bias"], stride=d.patch)
    x = x.flatten(2).transpose(1, 2)
    x = x + w[v + "pos_embed"]
    for l in range(d.layers):
        p = v + f"blocks.{l}."
        y = F.layer_norm(x, (d.width,), w[p + "norm1.weight"], w[p + "norm1.bias"], 1e-6)
        a = _mha_self(y, w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"], w[p + "attn.proj.weight"],
                      w[p + "attn.proj.bias"], d.heads)
        if l == d.layers - 1:
            return a
        x = x + a
        y = F.layer_norm(x, (d.width,), w[p + "norm2.weight"], w[p + "norm2.bias"], 1e-6)
        y = F.linear(y, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])
        y = F.gelu(y, approximate="tanh")
        y = F.linear(y, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
        x = x + y


def trunk_text_tokens(w, d: VerifierDims, tokens):
    """tokens [B,ctx] -> text_projection(ln_final(text.transformer(x))) for every token (ddp.py:320-327), bf16."""
    t = TR + "text."
    x = F.embedding(tokens, w[t + "token_embedding.weight"]) + w[t + "positional_embedding"]
    for l in range(d.text_layers):
        p = t + f"transformer.resblocks.{l}."
        y = F.layer_norm(x, (d.width,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], 1e-6)
        x = x + _mha_self(y, w[p + "attn.in_proj_weight"], w[p + "attn.in_proj_bias"], w[p + "attn.out_proj.weight"],
                          w[p + "attn.out_proj.bias"], d.heads)
        y = F.layer_norm(x, (d.width,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], 1e-6)
        y = F.linear(y, w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"])
        y = F.gelu(y, approximate="tanh")
        y = F.linear(y, w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"])
        x = x + y
    x = F.layer_norm(x, (d.width,), w[t + "ln_final.weight"], w[t + "ln_final.bias"], 1e-6)
    B, T, Wd = x.shape
    x = F.linear(x.reshape(-1, Wd), w[t + "text_projection.weight"], w[t + "text_projection.bias"])
    return x.reshape(B, T, -1)


@torch.no_grad()
def extract_features(w, d: VerifierDims, image, tokens):
    # finetune_trajectory_bridge_ddp.py:297-355
    text = trunk_text_tokens(w, d, tokens).float()
    text = text / text.norm(dim=-1, keepdim=True)
    patch = trunk_image_patches(w, d, image).float()
    patch = patch / patch.norm(dim=-1, keepdim=True)
    return patch, text


# ------------------------------------------------------------------------------------------------
# heads (reference code restated)
# ------------------------------------------------------------------------------------------------
def _mha_cross_1q(q, kv, wq, wk, wv, b_in, wo, bo, heads):
    """nn.MultiheadAttention(kdim=vdim!=embed_dim, batch_first=True), need_weights=True math path."""
    B, Tq, E = q.shape
    hd = E // heads
    qp = F.linear(q, wq, b_in[:E])
    kp = F.linear(kv, wk, b_in[E:2 * E])
    vp = F.linear(kv, wv, b_in[2 * E:])
    qp = qp.view(B, Tq, heads, hd).transpose(1, 2)
    kp = kp.view(B, -1, heads, hd).transpose(1, 2)
    vp = vp.view(B, -1, heads, hd).transpose(1, 2)
    att = torch.matmul(qp * math.sqrt(1.0 / hd), kp.transpose(-2, -1))
    att = F.softmax(att, dim=-1)
    o = torch.matmul(att, vp).transpose(1, 2).reshape(B, Tq, E)
    return F.linear(o, wo, bo)


def attention_pooling(w, prefix, d: VerifierDims, x):
    # model.py:97-112 with CrossAttentionBlock model.py:25-38
    B = x.shape[0]
    E = d.embed
    q = w[prefix + "query"].expand(B, -1, -1)
    for i in range(d.pool_layers):
        p = prefix + f"blocks.{i}."
        q = F.layer_norm(q, (E,), w[p + "q_layer_norm.weight"], w[p + "q_layer_norm.bias"])
        a = _mha_cross_1q(q, x, w[p + "attention.q_proj_weight"], w[p + "attention.k_proj_weight"],
                          w[p + "attention.v_proj_weight"], w[p + "attention.in_proj_bias"],
                          w[p + "attention.out_proj.weight"], w[p + "attention.out_proj.bias"], d.pool_heads)
        q = q + a
        q = F.layer_norm(q, (E,), w[p + "layer_norm.weight"], w[p + "layer_norm.bias"])
        y = F.linear(F.gelu(F.linear(q, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])), w[p + "mlp.fc2.weight"],
                     w[p + "mlp.fc2.bias"])
        q = q + y
    q = F.layer_norm(q, (E,), w[prefix + "layer_norm.weight"], w[prefix + "layer_norm.bias"])
    return q.reshape(B, -1)


def image_text_embedding(w, m: int, d: VerifierDims, patch, text):
    # efficient_ensemble_merged.py:216-223 (+ model.py:58-73)
    b = f"verifier.{m}."
    sim = torch.einsum("bij,bkj->bik", text, patch)
    att = F.softmax(sim / w[b + "text_aware_visual_extraction.temperature"].clamp(0, 100), dim=-1)
    pe = patch + w[b + "text_aware_visual_extraction.pos_emb"]
    taf = torch.einsum("bik,bkj->bij", att, pe)
    vision_token = attention_pooling(w, b + "vision_poolings.", d, taf)
    text_token = attention_pooling(w, b + "text_pooling.", d, text)
    c = torch.cat([text_token, vision_token], dim=-1)
    c = F.linear(c, w[b + "input_projection.weight"], w[b + "input_projection.bias"])
    return c / c.norm(dim=-1, keepdim=True)


def trajectory_embedding(w, m: int, d: VerifierDims, traj, pad_value=-5.0):
    # efficient_ensemble_merged.py:226-245 ; nn.TransformerEncoderLayer defaults (post-norm, ReLU, eps 1e-5)
    b = f"verifier.{m}."
    E, H = d.embed, d.pool_heads
    hd = E // H
    a = traj.float()
    if d.traj_layers == 0:
        # use_transformer = False (efficient_ensemble_merged.py:161-171, 241-245): nn.Sequential(Linear, LayerNorm, ReLU,
        # Dropout, Linear) over the flattened history; Dropout is the identity in eval mode
        q = b + "complex_action_encoder."
        hid = F.linear(a.reshape(a.shape[0], -1), w[q + "0.weight"], w[q + "0.bias"])
        hid = F.relu(F.layer_norm(hid, (hid.shape[-1],), w[q + "1.weight"], w[q + "1.bias"]))
        t = F.linear(hid, w[q + "4.weight"], w[q + "4.bias"])
        return t / t.norm(dim=-1, keepdim=True)
    pad = a[:, :, 0] == pad_value  # [N, S]
    x = F.linear(a, w[b + "single_step_action_encoder.weight"], w[b + "single_step_action_encoder.bias"])
    N, S, _ = x.shape
    for i in range(d.traj_layers):
        p = b + f"trajectory_encoder.layers.{i}."
        qkv = F.linear(x, w[p + "self_attn.in_proj_weight"], w[p + "self_attn.in_proj_bias"])
        q, k, v = qkv.split(E, dim=-1)
        q, k, v = (t.view(N, S, H, hd).transpose(1, 2) for t in (q, k, v))
        mask = torch.zeros(N, 1, 1, S).masked_fill(pad[:, None, None, :], float("-inf"))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
        o = o.transpose(1, 2).reshape(N, S, E)
        o = F.linear(o, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
        x = F.layer_norm(x + o, (E,), w[p + "norm1.weight"], w[p + "norm1.bias"])
        f = F.linear(F.relu(F.linear(x, w[p + "linear1.weight"], w[p + "linear1.bias"])), w[p + "linear2.weight"],
                     w[p + "linear2.bias"])
        x = F.layer_norm(x + f, (E,), w[p + "norm2.weight"], w[p + "norm2.bias"])
    keep = (~pad).unsqueeze(-1).float()
    summed = (x * keep).sum(dim=1)
    cnt = torch.clamp(keep.sum(dim=1), min=1e-9)
    t = summed / cnt
    return t / t.norm(dim=-1, keepdim=True)


@torch.no_grad()
def scores_from_features(w, d: VerifierDims, patch, text, traj):
    """fused scores [N] for ONE (image, instruction) against N trajectories (row 0 of the matrix)."""
    its, acts = [], []
    for m in range(d.members):
        its.append(image_text_embedding(w, m, d, patch, text))
        acts.append(trajectory_embedding(w, m, d, traj))
    fit = torch.stack(its).mean(dim=0)
    fact = torch.stack(acts).mean(dim=0)
    fit = fit / fit.norm(dim=-1, keepdim=True)
    fact = fact / fact.norm(dim=-1, keepdim=True)
    return torch.matmul(fit, fact.T)[0]


def select(scores: torch.Tensor, group_size: int):
    """efficient_ensemble_merged.py:417-447 -> (max_score, global_idx, best_group, group_means)"""
    g = scores.view(-1, group_size)
    means = g.mean(dim=1)
    _, gi = means.max(dim=0)
    best, ai = g[gi].max(dim=0)
    return float(best), int(gi * group_size + ai), int(gi), means


@torch.no_grad()
def compute_max_similarity_scores(w, d: VerifierDims, image, tokens, histories, group_size: int):
    patch, text = extract_features(w, d, image, tokens)
    traj = pad_histories(histories, d.history)
    scores = scores_from_features(w, d, patch, text, traj)
    best, idx, gi, means = select(scores, group_size)
    return best, idx, scores, means


