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
from dataclasses import asdict, dataclass

import numpy as np
import torch
import torch.nn.functional as F

from oracle.pi0_oracle import _gen


@dataclass
class VerifierDims:
    image: int = 384
    patch: int = 16
    width: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    text_layers: int = 24
    text_ctx: int = 64
    vocab: int = 256000
    members: int = 3
    embed: int = 512
    pool_heads: int = 8
    pool_layers: int = 4
    traj_layers: int = 4
    traj_ff: int = 1024
    history: int = 10
    action_dim: int = 7

    @property
    def n_patches(self) -> int:
        return (self.image // self.patch) ** 2

    def as_dict(self):
        return asdict(self)


VFULL = VerifierDims()
VTINY = VerifierDims(image=64, patch=16, width=128, layers=2, heads=2, mlp=256, text_layers=2, text_ctx=16,
                     vocab=500, members=2, embed=64, pool_heads=2, pool_layers=2, traj_layers=2, traj_ff=128)
VMID = VerifierDims(image=192, patch=16, width=256, layers=3, heads=4, mlp=1024, text_layers=3, text_ctx=64,
                    vocab=2000, members=3, embed=512, pool_heads=8, pool_layers=4, traj_layers=4, traj_ff=1024)

TR = "verifier.trunk."


def trunk_specs(d: VerifierDims):
    bf = torch.bfloat16
    out = []

    def lin(key, o, i, wname="weight", bname="bias"):
        out.append((key + wname, (o, i), 1.0 / math.sqrt(i), 0.0, bf))
        out.append((key + bname, (o,), 0.05, 0.0, bf))

    def ln(key):
        out.append((key + "weight", (d.width,), 0.1, 1.0, bf))
        out.append((key + "bias", (d.width,), 0.1, 0.0, bf))

    v = TR + "visual.trunk."
    out.append((v + "patch_embed.proj.weight", (d.width, 3, d.patch, d.patch), 1.0 / math.sqrt(3 * d.patch ** 2), 0.0, bf))
    out.append((v + "patch_embed.proj.bias", (d.width,), 0.05, 0.0, bf))
    out.append((v + "pos_embed", (1, d.n_patches, d.width), 0.5, 0.0, bf))
    for l in range(d.layers):
        p = v + f"blocks.{l}."
        ln(p + "norm1.")
        lin(p + "attn.qkv.", 3 * d.width, d.width)
        lin(p + "attn.proj.", d.width, d.width)
        if l < d.layers - 1:  # the last block's MLP is computed by the reference but never read
            ln(p + "norm2.")
            lin(p + "mlp.fc1.", d.mlp, d.width)
            lin(p + "mlp.fc2.", d.width, d.mlp)
    t = TR + "text."
    out.append((t + "token_embedding.weight", (d.vocab, d.width), 0.7, 0.0, bf))
    out.append((t + "positional_embedding", (d.text_ctx, d.width), 0.5, 0.0, bf))
    for l in range(d.text_layers):
        p = t + f"transformer.resblocks.{l}."
        ln(p + "ln_1.")
        lin(p + "attn.", 3 * d.width, d.width, "in_proj_weight", "in_proj_bias")
        lin(p + "attn.out_proj.", d.width, d.width)
        ln(p + "ln_2.")
        lin(p + "mlp.c_fc.", d.mlp, d.width)
        lin(p + "mlp.c_proj.", d.width, d.mlp)
    ln(t + "ln_final.")
    lin(t + "text_projection.", d.width, d.width)
    return out


def head_specs(d: VerifierDims):
    """Names = 'verifier.<m>.<component>.<state-dict key>' with the component / key names of the merged
    checkpoint (efficient_ensemble_merged.py:94-160)."""
    f32 = torch.float32
    E, Wd = d.embed, d.width
    out = []
    for m in range(d.members):
        b = f"verifier.{m}."
        out.append((b + "text_aware_visual_extraction.temperature", (), 0.0, 0.07, f32))
        out.append((b + "text_aware_visual_extraction.pos_emb", (d.n_patches, Wd), None, None, f32))  # sincos buffer
        for pool in ("vision_poolings", "text_pooling"):
            p = b + pool + "."
            out.append((p + "query", (1, 1, E), 1.0, 0.0, f32))
            out.append((p + "layer_norm.weight", (E,), 0.1, 1.0, f32))
            out.append((p + "layer_norm.bias", (E,), 0.1, 0.0, f32))
            for i in range(d.pool_layers):
                q = p + f"blocks.{i}."
                out.append((q + "attention.q_proj_weight", (E, E), 1.0 / math.sqrt(E), 0.0, f32))
                out.append((q + "attention.k_proj_weight", (E, Wd), 4.0 / math.sqrt(Wd), 0.0, f32))
                out.append((q + "attention.v_proj_weight", (E, Wd), 4.0 / math.sqrt(Wd), 0.0, f32))
                out.append((q + "attention.in_proj_bias", (3 * E,), 0.05, 0.0, f32))
                out.append((q + "attention.out_proj.weight", (E, E), 1.0 / math.sqrt(E), 0.0, f32))
                out.append((q + "attention.out_proj.bias", (E,), 0.05, 0.0, f32))
                out.append((q + "mlp.fc1.weight", (E, E), 1.0 / math.sqrt(E), 0.0, f32))
                out.append((q + "mlp.fc1.bias", (E,), 0.05, 0.0, f32))
                out.append((q + "mlp.fc2.weight", (E, E), 1.0 / math.sqrt(E), 0.0, f32))
                out.append((q + "mlp.fc2.bias", (E,), 0.05, 0.0, f32))
                for nm in ("q_layer_norm", "layer_norm"):
                    out.append((q + nm + ".weight", (E,), 0.1, 1.0, f32))
                    out.append((q + nm + ".bias", (E,), 0.1, 0.0, f32))
        out.append((b + "input_projection.weight", (E, 2 * E), 1.0 / math.sqrt(2 * E), 0.0, f32))
        out.append((b + "input_projection.bias", (E,), 0.05, 0.0, f32))
        out.append((b + "single_step_action_encoder.weight", (E, d.action_dim), 1.0 / math.sqrt(d.action_dim), 0.0, f32))
        out.append((b + "single_step_action_encoder.bias", (E,), 0.05, 0.0, f32))
        for i in range(d.traj_layers):
            q = b + f"trajectory_encoder.layers.{i}."
            out.append((q + "self_attn.in_proj_weight", (3 * E, E), 1.0 / math.sqrt(E), 0.0, f32))
            out.append((q + "self_attn.in_proj_bias", (3 * E,), 0.05, 0.0, f32))
            out.append((q + "self_attn.out_proj.weight", (E, E), 1.0 / math.sqrt(E), 0.0, f32))
            out.append((q + "self_attn.out_proj.bias", (E,), 0.05, 0.0, f32))
            out.append((q + "linear1.weight", (d.traj_ff, E), 1.0 / math.sqrt(E), 0.0, f32))
            out.append((q + "linear1.bias", (d.traj_ff,), 0.05, 0.0, f32))
            out.append((q + "linear2.weight", (E, d.traj_ff), 1.0 / math.sqrt(d.traj_ff), 0.0, f32))
            out.append((q + "linear2.bias", (E,), 0.05, 0.0, f32))
            for nm in ("norm1", "norm2"):
                out.append((q + nm + ".weight", (E,), 0.1, 1.0, f32))
                out.append((q + nm + ".bias", (E,), 0.1, 0.0, f32))
    return out


def sincos_position_embedding(seq_len: int, dim: int) -> torch.Tensor:
    # model.py:40-47
    pos = torch.arange(seq_len).float()
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
    sinusoid_inp = torch.einsum("i,j->ij", pos, inv_freq)
    return torch.cat((sinusoid_inp.sin(), sinusoid_inp.cos()), dim=-1)


def make_verifier_weights(d: VerifierDims, seed: int = 0, trunk: bool = True) -> dict:
    w = {}
    specs = head_specs(d) + (trunk_specs(d) if trunk else [])
    for k, shape, std, mean, dtype in specs:
        if k.endswith("text_aware_visual_extraction.pos_emb"):
            w[k] = sincos_position_embedding(d.n_patches, d.width)
        elif shape == ():
            w[k] = torch.tensor(mean, dtype=dtype)
        else:
            w[k] = _gen(k, seed + 77, shape, std, dtype, mean)
    return w


# ------------------------------------------------------------------------------------------------
# trunk (restated third-party architecture; see module docstring)
# ------------------------------------------------------------------------------------------------
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
    x = image.to(torch.bfloat16)
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


def pad_histories(histories, history: int = 10):
    # efficient_ensemble_merged.py:379-390 (left-pad with -5 to 10 steps)
    out = []
    for ah in histories:
        ah = np.asarray(ah)
        if len(ah) < history:
            ah = np.vstack([np.ones((history - len(ah), ah.shape[1])) * -5, ah])
        out.append(ah)
    return torch.tensor(np.array(out), dtype=torch.float32)


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


# ------------------------------------------------------------------------------------------------
def make_inputs(d: VerifierDims, N: int, seed: int = 0):
    g = torch.Generator().manual_seed(2000 + seed)
    image = torch.rand(1, 3, d.image, d.image, generator=g) * 2 - 1
    tokens = torch.randint(1, d.vocab - 1, (1, d.text_ctx), generator=g)
    hist = []
    for n in range(N):
        T = int(torch.randint(4, d.history + 1, (1,), generator=g))
        a = torch.rand(T, d.action_dim, generator=g) * 2 - 1
        a[:, :6] *= 0.05
        a[:, 6] = (a[:, 6] > 0).float()
        hist.append(a.numpy())
    return dict(image=image, tokens=tokens, histories=hist)
