"""TEST INFRASTRUCTURE - import shim that runs the UNMODIFIED reference arithmetic files from
/root/reference in this container.  Only `oracle/make_golden.py` and
`tests/test_oracle_vs_reference.py` use it; it never runs on the GPU box (no /root/reference there)
and nothing under cover_vla_b200/ imports it.

What is real reference code after `install()`:
  lerobot/common/policies/pi0/{modeling_pi0,paligemma_with_expert,flex_attention}.py
  lerobot/common/policies/normalize.py, lerobot/common/utils/utils.py, lerobot/configs/types.py
  bridge_verifier/ensemble_eval/{model,efficient_ensemble_merged,finetune_trajectory_bridge_ddp}.py
What is stubbed (package __init__ chains need draccus/jsonlines/imageio/open_clip/timm):
  lerobot.common.policies.pretrained.PreTrainedPolicy, ...pi0.configuration_pi0.PI0Config (fields of
  configuration_pi0.py:27-80), timm.layers.mlp.Mlp (fc1 -> GELU(erf) -> fc2), an empty open_clip.
"""
from __future__ import annotations

import sys
import types
from dataclasses import dataclass, field
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn as nn

REF = Path("/root/reference")
_installed = False


def available() -> bool:
    return (REF / "lerobot_custom/lerobot/common/policies/pi0/modeling_pi0.py").exists()


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("/root/reference is not present (the shim only works in the authoring container)")
    base = str(REF / "lerobot_custom/lerobot")
    for name, sub in [("lerobot", ""), ("lerobot.common", "/common"), ("lerobot.common.policies", "/common/policies"),
                      ("lerobot.common.policies.pi0", "/common/policies/pi0"), ("lerobot.configs", "/configs"),
                      ("lerobot.common.utils", "/common/utils")]:
        m = types.ModuleType(name)
        m.__path__ = [base + sub]
        sys.modules[name] = m

    pre = types.ModuleType("lerobot.common.policies.pretrained")

    class PreTrainedPolicy(nn.Module):
        def __init__(self, config, *a, **k):
            super().__init__()
            self.config = config

    pre.PreTrainedPolicy = PreTrainedPolicy
    sys.modules["lerobot.common.policies.pretrained"] = pre

    cfgm = types.ModuleType("lerobot.common.policies.pi0.configuration_pi0")

    @dataclass
    class PI0Config:
        n_obs_steps: int = 1
        chunk_size: int = 50
        n_action_steps: int = 50
        normalization_mapping: dict = field(default_factory=dict)
        max_state_dim: int = 32
        max_action_dim: int = 32
        resize_imgs_with_padding: tuple = (224, 224)
        empty_cameras: int = 0
        adapt_to_pi_aloha: bool = False
        use_delta_joint_actions_aloha: bool = False
        tokenizer_max_length: int = 48
        proj_width: int = 1024
        num_steps: int = 10
        use_cache: bool = True
        attention_implementation: str = "eager"
        freeze_vision_encoder: bool = True
        train_expert_only: bool = False
        train_state_proj: bool = True
        paligemma_pretrained_path: str | None = None

    cfgm.PI0Config = PI0Config
    sys.modules["lerobot.common.policies.pi0.configuration_pi0"] = cfgm

    # transformers >= 4.52 compat: restore the 4.48.3 attribute layout / get_image_features scaling
    from transformers import PaliGemmaForConditionalGeneration as PG

    PG.language_model = property(lambda s: SimpleNamespace(model=s.model.language_model))
    PG.vision_tower = property(lambda s: s.model.vision_tower)

    def _get_image_features(self, x):
        out = self.model.get_image_features(x)
        feats = out.pooler_output if hasattr(out, "pooler_output") else out
        return feats / (self.config.text_config.hidden_size ** 0.5)

    PG.get_image_features = _get_image_features

    # verifier stubs
    timm = types.ModuleType("timm")
    timm_layers = types.ModuleType("timm.layers")
    timm_mlp = types.ModuleType("timm.layers.mlp")

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, **kw):
            super().__init__()
            self.fc1 = nn.Linear(in_features, hidden_features or in_features)
            self.act = nn.GELU()
            self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    timm_mlp.Mlp = Mlp
    timm.layers = timm_layers
    timm_layers.mlp = timm_mlp
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.layers", timm_layers)
    sys.modules.setdefault("timm.layers.mlp", timm_mlp)
    oc = types.ModuleType("open_clip")
    oc.create_model_from_pretrained = None
    oc.get_tokenizer = None
    sys.modules.setdefault("open_clip", oc)
    bv = types.ModuleType("bridge_verifier")
    bv.__path__ = [str(REF / "bridge_verifier")]
    sys.modules["bridge_verifier"] = bv
    ee = types.ModuleType("bridge_verifier.ensemble_eval")
    ee.__path__ = [str(REF / "bridge_verifier/ensemble_eval")]
    sys.modules["bridge_verifier.ensemble_eval"] = ee
    _installed = True


def pi0_modules():
    install()
    import lerobot.common.policies.pi0.modeling_pi0 as M
    import lerobot.common.policies.pi0.paligemma_with_expert as P
    return M, P


def build_pi0(dims: dict, chunk_size=4, tokenizer_max_length=72, num_steps=10):
    """Reference PI0FlowMatching with the layer sizes in `dims` (see oracle/pi0_oracle.py:PI0Dims).
    Full-size = the defaults of paligemma_with_expert.py:81-150."""
    M, P = pi0_modules()
    from transformers import CONFIG_MAPPING
    cfg = sys.modules["lerobot.common.policies.pi0.configuration_pi0"].PI0Config(
        chunk_size=chunk_size, n_action_steps=chunk_size, tokenizer_max_length=tokenizer_max_length,
        num_steps=num_steps, proj_width=dims["ex_width"], max_state_dim=dims["max_state_dim"],
        max_action_dim=dims["max_action_dim"], paligemma_pretrained_path=None)
    base = P.PaliGemmaWithExpertConfig(paligemma_pretrained_path=None)
    n_img = (dims["vis_image"] // dims["vis_patch"]) ** 2
    pg = CONFIG_MAPPING["paligemma"](
        _vocab_size=dims["vocab"], bos_token_id=2, eos_token_id=1, hidden_size=dims["lm_width"],
        image_token_index=dims["vocab"], model_type="paligemma", pad_token_id=0, projection_dim=dims["lm_width"],
        text_config={"hidden_activation": "gelu_pytorch_tanh", "hidden_size": dims["lm_width"],
                     "intermediate_size": dims["lm_mlp"], "model_type": "gemma",
                     "num_attention_heads": dims["heads"], "num_hidden_layers": dims["layers"],
                     "num_image_tokens": n_img, "num_key_value_heads": 1, "head_dim": dims["head_dim"],
                     "torch_dtype": "float32", "vocab_size": dims["vocab"]},
        vision_config={"hidden_size": dims["vis_width"], "intermediate_size": dims["vis_mlp"],
                       "model_type": "siglip_vision_model", "num_attention_heads": dims["vis_heads"],
                       "num_hidden_layers": dims["vis_layers"], "num_image_tokens": n_img,
                       "patch_size": dims["vis_patch"], "image_size": dims["vis_image"],
                       "projection_dim": dims["lm_width"], "projector_hidden_act": "gelu_fast",
                       "torch_dtype": "float32", "vision_use_head": False})
    ge = CONFIG_MAPPING["gemma"](
        attention_bias=False, attention_dropout=0.0, bos_token_id=2, eos_token_id=1, head_dim=dims["head_dim"],
        hidden_act="gelu_pytorch_tanh", hidden_activation="gelu_pytorch_tanh", hidden_size=dims["ex_width"],
        initializer_range=0.02, intermediate_size=dims["ex_mlp"], max_position_embeddings=8192, model_type="gemma",
        num_attention_heads=dims["heads"], num_hidden_layers=dims["layers"], num_key_value_heads=1, pad_token_id=0,
        rms_norm_eps=1e-06, rope_theta=10000.0, torch_dtype="float32", use_cache=True, vocab_size=dims["vocab"])
    base.paligemma_config = pg
    base.gemma_expert_config = ge
    orig = P.PaliGemmaWithExpertConfig
    try:
        M.PaliGemmaWithExpertConfig = lambda **kw: base
        model = M.PI0FlowMatching(cfg)
    finally:
        M.PaliGemmaWithExpertConfig = orig
    # transformers >= 4.5x made Gemma's embed_tokens a *scaled* embedding (x sqrt(hidden)); the
    # reference targets 4.48.3 where it is a plain nn.Embedding and modeling_pi0.py:553 applies the
    # scale itself.  Neutralise the built-in scale so the 4.48.3 arithmetic is what runs.
    et = model.paligemma_with_expert.paligemma.model.language_model.embed_tokens
    if hasattr(et, "embed_scale"):
        et.embed_scale = torch.ones_like(et.embed_scale)
    return model.eval(), cfg


def verifier_modules():
    install()
    import bridge_verifier.ensemble_eval.model as VM
    import bridge_verifier.ensemble_eval.efficient_ensemble_merged as EM
    return VM, EM
