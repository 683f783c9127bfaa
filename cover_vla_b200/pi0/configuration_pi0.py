"""Host-side mirror of the reference PI0Config (lerobot_custom/lerobot/common/policies/pi0/
configuration_pi0.py:27-153): same field names and defaults, plus the model dimensions the reference
hard-codes in paligemma_with_expert.py:81-150 (they are configuration here so small test models work)."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class PolicyFeature:
    type: str
    shape: tuple


@dataclass
class PI0Config:
    n_obs_steps: int = 1
    chunk_size: int = 50
    n_action_steps: int = 50
    normalization_mapping: dict = field(default_factory=lambda: {"VISUAL": "IDENTITY", "STATE": "IDENTITY",
                                                                  "ACTION": "IDENTITY"})
    max_state_dim: int = 32
    max_action_dim: int = 32
    resize_imgs_with_padding: tuple | None = (224, 224)
    empty_cameras: int = 0
    adapt_to_pi_aloha: bool = False
    use_delta_joint_actions_aloha: bool = False
    tokenizer_max_length: int = 48
    proj_width: int = 1024
    num_steps: int = 10
    use_cache: bool = True
    attention_implementation: str = "eager"
    device: str = "cuda"
    # features (reference: PreTrainedConfig.input_features / output_features)
    input_features: dict = field(default_factory=lambda: {
        "observation.images.top": PolicyFeature("VISUAL", (3, 224, 224)),
        "observation.state": PolicyFeature("STATE", (7,)),
    })
    output_features: dict = field(default_factory=lambda: {"action": PolicyFeature("ACTION", (7,))})
    # model dimensions (reference hard-codes these: paligemma_with_expert.py:81-150)
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
    ex_mlp: int = 4096
    vocab: int = 257152
    # workspace sizing of the engine
    max_rephrases: int = 8
    max_samples: int = 5

    @property
    def image_features(self) -> dict:
        return {k: v for k, v in self.input_features.items() if v.type == "VISUAL"}

    @property
    def action_feature(self) -> PolicyFeature:
        return self.output_features["action"]

    @property
    def robot_state_feature(self) -> PolicyFeature:
        return self.input_features["observation.state"]

    def validate_features(self):
        if not self.image_features:
            raise ValueError("PI0Config needs at least one VISUAL input feature")
        if self.n_action_steps > self.chunk_size:
            raise ValueError(f"n_action_steps ({self.n_action_steps}) must be <= chunk_size ({self.chunk_size})")
        if self.attention_implementation != "eager":
            raise ValueError("only the 'eager' attention semantics are implemented (the Bridge checkpoints use it)")

    @classmethod
    def bridge(cls, **kw) -> "PI0Config":
        """INT-ACT/config/models/pi0_finetune_bridge.json: chunk 4, 72 tokens, one 224^2 camera, IDENTITY norms."""
        return cls(chunk_size=4, n_action_steps=4, tokenizer_max_length=72, **kw)
