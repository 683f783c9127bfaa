"""Shared test helpers: build an Engine from oracle-generated weights, compare tensors."""
from __future__ import annotations

import torch

from oracle import pi0_oracle as O


def engine_config_from_dims(d: O.PI0Dims, max_R: int, max_K: int, **kw):
    from cover_vla_b200.engine import EngineConfig
    return EngineConfig(vis_layers=d.vis_layers, vis_width=d.vis_width, vis_heads=d.vis_heads, vis_mlp=d.vis_mlp,
                        vis_patch=d.vis_patch, vis_image=d.vis_image, layers=d.layers, lm_width=d.lm_width,
                        lm_mlp=d.lm_mlp, heads=d.heads, head_dim=d.head_dim, ex_width=d.ex_width, ex_mlp=d.ex_mlp,
                        vocab=d.vocab, max_state_dim=d.max_state_dim, max_action_dim=d.max_action_dim,
                        chunk_size=d.chunk_size, max_lang_len=d.max_lang_len, num_steps=d.num_steps,
                        max_rephrases=max_R, max_samples=max_K, **kw)


def build_pi0_engine(d: O.PI0Dims, w: dict, max_R: int, max_K: int, **kw):
    from cover_vla_b200.engine import Engine
    eng = Engine(engine_config_from_dims(d, max_R, max_K, **kw))
    for k, v in w.items():
        eng.bind("model." + k, v.cuda())
    eng.finalize()
    return eng


def rel_l2(x, y):
    x, y = x.float().cpu(), y.float().cpu()
    return ((x - y).norm() / (y.norm() + 1e-12)).item()


def max_abs(x, y):
    return (x.float().cpu() - y.float().cpu()).abs().max().item()


def verifier_config_kwargs(v):
    return dict(vf_image=v.image, vf_patch=v.patch, vf_width=v.width, vf_layers=v.layers, vf_heads=v.heads,
                vf_mlp=v.mlp, vf_text_layers=v.text_layers, vf_text_ctx=v.text_ctx, vf_vocab=v.vocab,
                vf_members=v.members, vf_embed=v.embed, vf_pool_heads=v.pool_heads, vf_pool_layers=v.pool_layers,
                vf_traj_layers=v.traj_layers, vf_traj_ff=v.traj_ff, vf_history=v.history, vf_action_dim=v.action_dim)


def build_full_engine(d, w, v, vw, max_R, max_K, **kw):
    """pi0 + verifier in one handle."""
    from cover_vla_b200.engine import Engine
    eng = Engine(engine_config_from_dims(d, max_R, max_K, **verifier_config_kwargs(v), **kw))
    for k, t in w.items():
        eng.bind("model." + k, t.cuda())
    for k, t in vw.items():
        eng.bind(k, t.cuda())
    eng.finalize()
    return eng
