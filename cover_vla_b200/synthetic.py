"""Deterministic synthetic weights and inputs for the CoVer sample-and-verify path (no network: neither the
pi0 checkpoint nor the verifier checkpoint can be downloaded, so every benchmark / test configuration of
BASELINE.json runs on random-init weights of the right architecture, generated from a seed).

Shared by bench.py, __graft_entry__.py, the tests and the oracle; contains no model arithmetic.
Tensor names are the reference state-dict names.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import asdict, dataclass

import numpy as np
import torch


@dataclass
class PI0Dims:
    vis_layers: int = 27
    vis_width: int = 1152
    vis_heads: int = 16
    vis_mlp: int = 4304
    vis_patch: int = 14
    vis_image: int = 224
    layers: int = 18
    lm_width: int = 2048
    lm_mlp: int = 16384
    heads: int = 8
    head_dim: int = 256
    ex_width: int = 1024
    ex_mlp: int = 4096
    vocab: int = 257152
    max_state_dim: int = 32
    max_action_dim: int = 32
    chunk_size: int = 4
    max_lang_len: int = 72
    num_steps: int = 10

    @property
    def n_img_tokens(self) -> int:
        return (self.vis_image // self.vis_patch) ** 2

    def as_dict(self):
        return asdict(self)


FULL = PI0Dims()
TINY = PI0Dims(vis_layers=2, vis_width=128, vis_heads=2, vis_mlp=256, vis_patch=14, vis_image=56,
               layers=3, lm_width=128, lm_mlp=512, heads=2, head_dim=64, ex_width=64, ex_mlp=256,
               vocab=1000, max_lang_len=16)
# mid-size: every awkward property of the full model (head_dim 72 in the tower, K tails, 256-d heads)
# at a size the CPU oracle finishes in seconds
MID = PI0Dims(vis_layers=3, vis_width=288, vis_heads=4, vis_mlp=1072, vis_patch=14, vis_image=224,
              layers=4, lm_width=512, lm_mlp=2048, heads=8, head_dim=256, ex_width=256, ex_mlp=1024,
              vocab=4096, max_lang_len=72)

PW = "paligemma_with_expert."
VT = PW + "paligemma.vision_tower.vision_model."
MM = PW + "paligemma.multi_modal_projector.linear."
LM = PW + "paligemma.language_model.model."
EX = PW + "gemma_expert.model."


# ------------------------------------------------------------------------------------------------
# deterministic synthetic weights (canonical = transformers-4.48.3 key names, no "model." prefix)
# ------------------------------------------------------------------------------------------------
def _gen(name: str, seed: int, shape, std: float, dtype, mean: float = 0.0):
    h = int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:8], "little") % (2 ** 62)
    g = torch.Generator(device="cpu").manual_seed(h)
    n = 1
    for s in shape:
        n *= s
    if n > (1 << 26):  # huge tables (token embedding): tile a 4M-element random block
        blk = torch.empty(1 << 22, dtype=torch.float32).normal_(0.0, std, generator=g)
        t = blk.repeat((n + blk.numel() - 1) // blk.numel())[:n].reshape(shape)
    else:
        t = torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=g)
    if mean != 0.0:
        t = t + mean
    return t.to(dtype)


def weight_specs(d: PI0Dims):
    """(key, shape, std, mean, dtype) for every tensor the sampling path reads."""
    bf, f32 = torch.bfloat16, torch.float32
    out = []

    def lin(key, o, i, dtype=bf, bias=False):
        out.append((key + ".weight", (o, i), 1.0 / math.sqrt(i), 0.0, dtype))
        if bias:
            out.append((key + ".bias", (o,), 0.05, 0.0, dtype))

    out.append((VT + "embeddings.patch_embedding.weight", (d.vis_width, 3, d.vis_patch, d.vis_patch),
                1.0 / math.sqrt(3 * d.vis_patch ** 2), 0.0, bf))
    out.append((VT + "embeddings.patch_embedding.bias", (d.vis_width,), 0.05, 0.0, bf))
    out.append((VT + "embeddings.position_embedding.weight", (d.n_img_tokens, d.vis_width), 0.5, 0.0, bf))
    for l in range(d.vis_layers):
        p = VT + f"encoder.layers.{l}."
        for ln in ("layer_norm1", "layer_norm2"):
            out.append((p + ln + ".weight", (d.vis_width,), 0.1, 1.0, bf))
            out.append((p + ln + ".bias", (d.vis_width,), 0.1, 0.0, bf))
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lin(p + "self_attn." + nm, d.vis_width, d.vis_width, bias=True)
        lin(p + "mlp.fc1", d.vis_mlp, d.vis_width, bias=True)
        lin(p + "mlp.fc2", d.vis_width, d.vis_mlp, bias=True)
    out.append((VT + "post_layernorm.weight", (d.vis_width,), 0.1, 1.0, bf))
    out.append((VT + "post_layernorm.bias", (d.vis_width,), 0.1, 0.0, bf))
    lin(MM[:-1], d.lm_width, d.vis_width, bias=True)
    out.append((LM + "embed_tokens.weight", (d.vocab, d.lm_width), 1.0 / math.sqrt(d.lm_width), 0.0, bf))
    qd = d.heads * d.head_dim
    for l in range(d.layers):
        p = LM + f"layers.{l}."
        lin(p + "self_attn.q_proj", qd, d.lm_width)
        lin(p + "self_attn.k_proj", d.head_dim, d.lm_width)
        lin(p + "self_attn.v_proj", d.head_dim, d.lm_width)
        lin(p + "self_attn.o_proj", d.lm_width, qd)
        lin(p + "mlp.gate_proj", d.lm_mlp, d.lm_width)
        lin(p + "mlp.up_proj", d.lm_mlp, d.lm_width)
        lin(p + "mlp.down_proj", d.lm_width, d.lm_mlp)
        out.append((p + "input_layernorm.weight", (d.lm_width,), 0.1, 0.0, bf))
        out.append((p + "post_attention_layernorm.weight", (d.lm_width,), 0.1, 0.0, bf))
        p = EX + f"layers.{l}."
        lin(p + "self_attn.q_proj", qd, d.ex_width)
        lin(p + "self_attn.k_proj", d.head_dim, d.ex_width)
        lin(p + "self_attn.v_proj", d.head_dim, d.ex_width)
        lin(p + "self_attn.o_proj", d.ex_width, qd)
        lin(p + "mlp.gate_proj", d.ex_mlp, d.ex_width)
        lin(p + "mlp.up_proj", d.ex_mlp, d.ex_width)
        lin(p + "mlp.down_proj", d.ex_width, d.ex_mlp)
        out.append((p + "input_layernorm.weight", (d.ex_width,), 0.1, 0.0, bf))
        out.append((p + "post_attention_layernorm.weight", (d.ex_width,), 0.1, 0.0, bf))
    out.append((EX + "norm.weight", (d.ex_width,), 0.1, 0.0, f32))
    lin("state_proj", d.ex_width, d.max_state_dim, f32, bias=True)
    lin("action_in_proj", d.ex_width, d.max_action_dim, f32, bias=True)
    lin("action_out_proj", d.max_action_dim, d.ex_width, f32, bias=True)
    lin("action_time_mlp_in", d.ex_width, 2 * d.ex_width, f32, bias=True)
    lin("action_time_mlp_out", d.ex_width, d.ex_width, f32, bias=True)
    return out


def make_pi0_weights(d: PI0Dims, seed: int = 0) -> dict:
    return {k: _gen(k, seed, shape, std, dtype, mean) for k, shape, std, mean, dtype in weight_specs(d)}


def to_hf5_key(k: str) -> str:
    """canonical (4.48.3) -> transformers>=4.52 module layout (what the shim-built reference uses)."""
    k = k.replace("paligemma.vision_tower.", "paligemma.model.vision_tower.")
    k = k.replace("paligemma.multi_modal_projector.", "paligemma.model.multi_modal_projector.")
    k = k.replace("paligemma.language_model.model.", "paligemma.model.language_model.")
    return k


def canonical_key(k: str) -> str:
    """accept 'model.'-prefixed PI0Policy keys and either transformers layout."""
    if k.startswith("model."):
        k = k[len("model."):]
    k = k.replace("paligemma.model.vision_tower.", "paligemma.vision_tower.")
    k = k.replace("paligemma.model.multi_modal_projector.", "paligemma.multi_modal_projector.")
    k = k.replace("paligemma.model.language_model.", "paligemma.language_model.model.")
    return k


# ------------------------------------------------------------------------------------------------
# synthetic pi0 inputs
# ------------------------------------------------------------------------------------------------
def make_inputs(d: PI0Dims, R: int, K: int, seed: int = 0, noise_std: float = 1.0):
    g = torch.Generator().manual_seed(1000 + seed)
    image = torch.rand(1, 3, d.vis_image, d.vis_image, generator=g) * 2 - 1
    lens = torch.randint(min(8, d.max_lang_len), min(24, d.max_lang_len) + 1, (R,), generator=g)
    tokens = torch.randint(3, d.vocab - 1, (R, d.max_lang_len), generator=g)
    masks = torch.arange(d.max_lang_len)[None, :] < lens[:, None]
    tokens = torch.where(masks, tokens, torch.zeros_like(tokens))
    state = torch.zeros(1, d.max_state_dim)
    state[0, :7] = torch.randn(7, generator=g)
    noise = torch.randn(R * K, d.chunk_size, d.max_action_dim, generator=g) * noise_std
    return dict(image=image, tokens=tokens, masks=masks, state=state, noise=noise, lens=lens)


def expand_to_batch(inp, K):
    """What run_simpler_eval_with_openpi.py:305-319 hands to select_action (rephrase-major)."""
    R = inp["tokens"].shape[0]
    N = R * K
    rep = torch.arange(R).repeat_interleave(K)
    return dict(image=inp["image"].repeat(N, 1, 1, 1), tokens=inp["tokens"][rep], masks=inp["masks"][rep],
                state=inp["state"].repeat(N, 1), noise=inp["noise"])


# ================================================================================================
# verifier
# ================================================================================================
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

# use_transformer = False checkpoints: traj_layers = 0, traj_ff = hidden width of the MLP action encoder
VTINY_MLP = VerifierDims(**{**asdict(VTINY), "traj_layers": 0, "traj_ff": 48})
VMID_MLP = VerifierDims(**{**asdict(VMID), "traj_layers": 0, "traj_ff": 512})

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
        if d.traj_layers == 0:  # use_transformer = False: the MLP complex_action_encoder (hidden width = traj_ff)
            q, Hm, Kin = b + "complex_action_encoder.", d.traj_ff, d.history * d.action_dim
            out.append((q + "0.weight", (Hm, Kin), 1.0 / math.sqrt(Kin), 0.0, f32))
            out.append((q + "0.bias", (Hm,), 0.05, 0.0, f32))
            out.append((q + "1.weight", (Hm,), 0.1, 1.0, f32))
            out.append((q + "1.bias", (Hm,), 0.1, 0.0, f32))
            out.append((q + "4.weight", (E, Hm), 1.0 / math.sqrt(Hm), 0.0, f32))
            out.append((q + "4.bias", (E,), 0.05, 0.0, f32))
            continue
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


def pad_histories(histories, history: int = 10):
    # efficient_ensemble_merged.py:379-390 (left-pad with -5 to 10 steps)
    out = []
    for ah in histories:
        ah = np.asarray(ah)
        if len(ah) < history:
            ah = np.vstack([np.ones((history - len(ah), ah.shape[1])) * -5, ah])
        out.append(ah)
    return torch.tensor(np.array(out), dtype=torch.float32)




# ------------------------------------------------------------------------------------------------
def make_verifier_inputs(d: VerifierDims, N: int, seed: int = 0):
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


# ================================================================================================
# engine construction from synthetic weights
# ================================================================================================
def engine_config(d: PI0Dims | None, v: VerifierDims | None, max_R: int, max_K: int, **kw):
    from .engine import EngineConfig
    args = dict(max_rephrases=max_R, max_samples=max_K)
    if d is not None:
        args.update(vis_layers=d.vis_layers, vis_width=d.vis_width, vis_heads=d.vis_heads, vis_mlp=d.vis_mlp,
                    vis_patch=d.vis_patch, vis_image=d.vis_image, layers=d.layers, lm_width=d.lm_width,
                    lm_mlp=d.lm_mlp, heads=d.heads, head_dim=d.head_dim, ex_width=d.ex_width, ex_mlp=d.ex_mlp,
                    vocab=d.vocab, max_state_dim=d.max_state_dim, max_action_dim=d.max_action_dim,
                    chunk_size=d.chunk_size, max_lang_len=d.max_lang_len, num_steps=d.num_steps)
    else:
        args.update(layers=0, vis_layers=0)
    if v is not None:
        args.update(vf_image=v.image, vf_patch=v.patch, vf_width=v.width, vf_layers=v.layers, vf_heads=v.heads,
                    vf_mlp=v.mlp, vf_text_layers=v.text_layers, vf_text_ctx=v.text_ctx, vf_vocab=v.vocab,
                    vf_members=v.members, vf_embed=v.embed, vf_pool_heads=v.pool_heads,
                    vf_pool_layers=v.pool_layers, vf_traj_layers=v.traj_layers, vf_traj_ff=v.traj_ff,
                    vf_history=v.history, vf_action_dim=v.action_dim)
    args.update(kw)
    return EngineConfig(**args)


def build_engine(d: PI0Dims | None, w: dict | None, v: VerifierDims | None, vw: dict | None, max_R: int,
                 max_K: int, device="cuda", **kw):
    """Engine with the given synthetic weights bound under their reference state-dict names."""
    from .engine import Engine
    eng = Engine(engine_config(d, v, max_R, max_K, **kw), device=device)
    if w is not None:
        for k, t in w.items():
            eng.bind("model." + k, t.to(device))
    if vw is not None:
        for k, t in vw.items():
            eng.bind(k, t.to(device))
    eng.finalize()
    return eng
