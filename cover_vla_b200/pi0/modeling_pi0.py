"""Drop-in host surface for the reference pi0 policy, running on the coverb200 CUDA engine.

Mirrors (same names, argument meaning, return types and error behaviour):
  PI0Policy.select_action / reset / prepare_images / prepare_state / prepare_language
      lerobot_custom/lerobot/common/policies/pi0/modeling_pi0.py:263-307, :255-257, :344-441
  PI0FlowMatching.sample_actions / sample_noise
      modeling_pi0.py:672-715, :502-510
The arithmetic is NOT here: sample_actions marshals pointers into cvb_pi0_sample (include/coverb200.h).
There is no CPU path; constructing the policy without the built library or without a GPU raises.
"""
from __future__ import annotations

import json
from collections import deque
from pathlib import Path

import torch
import torch.nn.functional as F
from torch import Tensor

from ..engine import Engine, EngineConfig
from .configuration_pi0 import PI0Config, PolicyFeature

OBS_ROBOT = "observation.state"
ACTION = "action"


def resize_with_pad(img, width, height, pad_value=-1):
    # modeling_pi0.py:131-150 (host glue; identity for 224x224 inputs)
    if img.ndim != 4:
        raise ValueError(f"(b,c,h,w) expected, but {img.shape}")
    cur_height, cur_width = img.shape[2:]
    if (cur_height, cur_width) == (height, width):
        return img
    ratio = max(cur_width / width, cur_height / height)
    resized_height = int(cur_height / ratio)
    resized_width = int(cur_width / ratio)
    resized = F.interpolate(img, size=(resized_height, resized_width), mode="bilinear", align_corners=False)
    pad_height = max(0, int(height - resized_height))
    pad_width = max(0, int(width - resized_width))
    return F.pad(resized, (pad_width, 0, pad_height, 0), value=pad_value)


def pad_vector(vector, new_dim):
    # modeling_pi0.py:153-165
    if vector.shape[-1] == new_dim:
        return vector
    shape = list(vector.shape)
    current_dim = shape[-1]
    shape[-1] = new_dim
    new_vector = torch.zeros(*shape, dtype=vector.dtype, device=vector.device)
    new_vector[..., :current_dim] = vector
    return new_vector


def engine_config_from_policy_config(c: PI0Config, **extra) -> EngineConfig:
    try:  # one image stream per camera key of the policy config (the Bridge checkpoints have one)
        extra.setdefault("num_cameras", max(1, len(c.image_features)))
    except Exception:  # noqa: BLE001 - configs without feature declarations
        extra.setdefault("num_cameras", 1)
    return EngineConfig(vis_layers=c.vis_layers, vis_width=c.vis_width, vis_heads=c.vis_heads, vis_mlp=c.vis_mlp,
                        vis_patch=c.vis_patch, vis_image=c.vis_image, layers=c.layers, lm_width=c.lm_width,
                        lm_mlp=c.lm_mlp, heads=c.heads, head_dim=c.head_dim, ex_width=c.proj_width, ex_mlp=c.ex_mlp,
                        vocab=c.vocab, max_state_dim=c.max_state_dim, max_action_dim=c.max_action_dim,
                        chunk_size=c.chunk_size, max_lang_len=c.tokenizer_max_length, num_steps=c.num_steps,
                        max_rephrases=c.max_rephrases, max_samples=c.max_samples, **extra)


class PI0FlowMatching:
    """modeling_pi0.py:448-752 - only the inference surface (`sample_actions`) exists here."""

    def __init__(self, config: PI0Config, engine: Engine):
        self.config = config
        self.engine = engine
        # set True when the caller guarantees the CoVer batch layout (one observation, rephrase-major
        # rows, K samples per rephrase): skips the two device->host checks below
        self.assume_cover_layout = False
        self.samples_per_rephrase: int | None = None
        # bound on valid language tokens per prompt known WITHOUT a device sync (set by PI0Policy.prepare_language from
        # the tokenizer's host output); None = process all tokenizer_max_length rows
        self.lang_len_hint: int | None = None

    def sample_noise(self, shape, device, noise_std=1.0):
        return torch.normal(mean=0.0, std=noise_std, size=shape, dtype=torch.float32, device=device)

    def _layout(self, images, lang_tokens):
        """(R, K) of the batch: rows are rephrase-major with K identical token rows per rephrase
        (run_simpler_eval_with_openpi.py:305-319).  Anything else degrades to R = N, K = 1."""
        N = lang_tokens.shape[0]
        if self.assume_cover_layout and self.samples_per_rephrase:
            K = self.samples_per_rephrase
            if N % K != 0:
                raise ValueError(f"batch size {N} is not a multiple of samples_per_rephrase {K}")
            return N // K, K
        _, counts = torch.unique_consecutive(lang_tokens, dim=0, return_counts=True)
        counts = counts.tolist()  # one device->host sync
        if len(set(counts)) == 1:
            return len(counts), counts[0]
        return N, 1

    def sample_actions(self, images, img_masks, lang_tokens, lang_masks, state, noise=None, noise_std=1.0) -> Tensor:
        """Full inference forward: (batch_size x chunk_size x max_action_dim) actions.

        One engine call handles one observation: the reference caller replicates the same image and state
        over the batch (run_simpler_eval_with_openpi.py:312-313); batches holding several distinct
        observations are split into per-observation calls.
        """
        cfg = self.config
        bsize = state.shape[0]
        # Cameras (modeling_pi0.py:344-387, 529-547): a camera whose mask is False for the whole batch is an "empty camera"
        # (-1 image): its tokens are masked as keys, do not advance the position ids and are never read, so it is dropped
        # (exact).  The present cameras are stacked [batch, C, 3, H, W] in the order the reference concatenates them.
        if img_masks is not None and len(img_masks) == len(images) and not self.assume_cover_layout:
            keep = []
            for cam, m in enumerate(img_masks):
                n_on = int(m.sum())
                if n_on not in (0, m.numel()):
                    raise NotImplementedError("a camera must be present (or absent) for the whole batch")
                if n_on:
                    keep.append(cam)
            if not keep:
                raise ValueError("every camera is masked out")
            images = [images[cam] for cam in keep]
        if len(images) > max(1, self.engine.cfg.num_cameras):
            raise ValueError(f"{len(images)} cameras, but the engine was built for {self.engine.cfg.num_cameras} "
                             "(EngineConfig.num_cameras / len(config.image_features))")
        self.engine.set_active_cameras(len(images))
        img = images[0] if len(images) == 1 else torch.stack(list(images), dim=1)
        device = state.device
        if device.type != "cuda":
            raise RuntimeError("coverb200 has no CPU path: move the observation to a CUDA device")
        if noise is None:
            noise = self.sample_noise((bsize, cfg.chunk_size, cfg.max_action_dim), device, noise_std)
        img = img.to(torch.float32)
        state = state.to(torch.float32)
        lang_tokens = lang_tokens.to(torch.int64)
        lang_len = lang_masks.sum(dim=1).to(torch.int32)
        if not self.assume_cover_layout:
            prefix_ok = lang_masks == (torch.arange(lang_masks.shape[1], device=device)[None, :] < lang_len[:, None])
            same_obs = (img == img[:1]).all() & (state == state[:1]).all()
            prefix_ok, same_obs = bool(prefix_ok.all()), bool(same_obs)
            hint = int(lang_len.max())  # these checks already synchronise: the exact bound is free here
            if not prefix_ok:
                raise NotImplementedError("language masks must be right-padded (tokenizer padding_side='right')")
            if not same_obs:
                out = torch.empty_like(noise)
                for i in range(bsize):  # distinct observations: one engine call each
                    out[i:i + 1] = self.engine.pi0_sample(img[i].contiguous(), lang_tokens[i:i + 1].contiguous(),
                                                          lang_len[i:i + 1].contiguous(), state[i].contiguous(),
                                                          noise[i:i + 1].contiguous(), K=1)
                return out
        else:
            hint = self.lang_len_hint
        R, K = self._layout(images, lang_tokens)
        if R > self.engine.cfg.max_rephrases or K > self.engine.cfg.max_samples:
            # e.g. two adjacent rephrases tokenise identically (duplicates, truncation) so the rows do not group into the
            # configured R x K, or the batch is larger than the workspace: the reference handles any batch, so fall back
            # to row-by-row prompts (K = 1) in chunks of max_rephrases - same arithmetic, no de-duplication
            out = torch.empty_like(noise)
            step = self.engine.cfg.max_rephrases
            for a in range(0, bsize, step):
                b = min(bsize, a + step)
                out[a:b] = self.engine.pi0_sample(img[0].contiguous(), lang_tokens[a:b].contiguous(),
                                                  lang_len[a:b].contiguous(), state[0].contiguous(),
                                                  noise[a:b].contiguous(), K=1, lang_len_max=hint)
            return out
        return self.engine.pi0_sample(img[0].contiguous(), lang_tokens[::K].contiguous(), lang_len[::K].contiguous(),
                                      state[0].contiguous(), noise.contiguous(), K=K, lang_len_max=hint)


class _Normalize:
    """lerobot_custom/lerobot/common/policies/normalize.py:153-183 / :227-254 for the three modes."""

    def __init__(self, features: dict, mapping: dict, stats: dict | None, inverse: bool):
        self.features, self.mapping, self.stats, self.inverse = features, mapping, stats or {}, inverse

    def __call__(self, batch: dict) -> dict:
        batch = dict(batch)
        for key, ft in self.features.items():
            mode = self.mapping.get(ft.type, "IDENTITY")
            mode = getattr(mode, "name", mode)
            if mode == "IDENTITY" or key not in batch:
                continue
            st = self.stats.get(key)
            if st is None:
                raise ValueError(f"normalisation mode {mode} for '{key}' needs dataset statistics")
            x = batch[key]
            if mode == "MEAN_STD":
                mean, std = st["mean"].to(x), st["std"].to(x)
                batch[key] = x * std + mean if self.inverse else (x - mean) / (std + 1e-8)
            elif mode == "MIN_MAX":
                lo, hi = st["min"].to(x), st["max"].to(x)
                batch[key] = (x + 1) / 2 * (hi - lo) + lo if self.inverse else (x - lo) / (hi - lo + 1e-8) * 2 - 1
            else:
                raise ValueError(mode)
        return batch


class PI0Policy:
    """modeling_pi0.py:216-307.  `select_action` keeps the reference's (customised) contract: it returns the
    internal deque of n_action_steps tensors [batch, action_dim]; the caller copies and clears it
    (run_simpler_eval_with_openpi.py:324-326)."""

    config_class = PI0Config
    name = "pi0"

    def __init__(self, config: PI0Config, state_dict: dict | None = None, dataset_stats: dict | None = None,
                 language_tokenizer=None, engine: Engine | None = None, engine_extra: dict | None = None):
        config.validate_features()
        self.config = config
        self.normalize_inputs = _Normalize(config.input_features, config.normalization_mapping, dataset_stats, False)
        self.unnormalize_outputs = _Normalize(config.output_features, config.normalization_mapping, dataset_stats, True)
        self.language_tokenizer = language_tokenizer
        self._prompt_cache: dict = {}
        if engine is None:
            if state_dict is None:
                raise ValueError("PI0Policy needs the reference state dict (or a finalized Engine)")
            engine = Engine(engine_config_from_policy_config(config, **(engine_extra or {})), device=config.device)
            engine.load_state_dict(state_dict)
            engine.finalize()
        self.engine = engine
        self.model = PI0FlowMatching(config, engine)
        self._preprocess_adapter = None
        self.reset()

    # -- nn.Module-ish conveniences the caller uses
    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("coverb200 has no CPU path")
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter(())

    @classmethod
    def from_pretrained(cls, pretrained_name_or_path, config: PI0Config | None = None, language_tokenizer=None, **kw):
        """Load `config.json` + `model.safetensors` written by the reference's save_pretrained
        (policies/pretrained.py:71-151).  Tensor names are bound unchanged."""
        from safetensors.torch import load_file
        path = Path(pretrained_name_or_path)
        if not path.is_dir():
            raise FileNotFoundError(f"{path} is not a local directory (no network access: hub ids are not resolved)")
        if config is None:
            raw = json.loads((path / "config.json").read_text())
            known = {f for f in PI0Config.__dataclass_fields__}
            args = {k: v for k, v in raw.items() if k in known and k not in ("input_features", "output_features",
                                                                             "normalization_mapping")}
            # features and the normalisation mapping as the reference's PreTrainedConfig stores them
            # (configs/policies.py:130-176): {"name": {"type": "VISUAL", "shape": [3, 224, 224]}}, {"VISUAL": "IDENTITY"}
            for fk in ("input_features", "output_features"):
                if isinstance(raw.get(fk), dict):
                    args[fk] = {name: PolicyFeature(str(ft["type"]).split(".")[-1], tuple(ft["shape"]))
                                for name, ft in raw[fk].items()}
            if isinstance(raw.get("normalization_mapping"), dict):
                args["normalization_mapping"] = {str(k).split(".")[-1]: str(v).split(".")[-1]
                                                 for k, v in raw["normalization_mapping"].items()}
            if isinstance(args.get("resize_imgs_with_padding"), list):
                args["resize_imgs_with_padding"] = tuple(args["resize_imgs_with_padding"])
            config = PI0Config(**args)
        sd = load_file(str(path / "model.safetensors"))
        # dataset statistics travel as buffers of the (un)normalisation modules (normalize.py:40-43):
        # normalize_inputs.buffer_<feature with . -> _>.<mean|std|min|max>
        if "dataset_stats" not in kw:
            stats = {}
            for k, v in sd.items():
                if k.startswith(("normalize_inputs.buffer_", "unnormalize_outputs.buffer_")):
                    feat, stat = k.split("buffer_", 1)[1].rsplit(".", 1)
                    stats.setdefault(feat, {})[stat] = v
            if stats:
                names = list(config.input_features) + list(config.output_features)
                kw["dataset_stats"] = {n: stats[n.replace(".", "_")] for n in names if n.replace(".", "_") in stats}
        need = {t for t, m in config.normalization_mapping.items() if str(getattr(m, "name", m)) != "IDENTITY"}
        used = {ft.type for ft in list(config.input_features.values()) + list(config.output_features.values())}
        if need & used and not kw.get("dataset_stats"):
            raise ValueError(f"checkpoint config normalises {sorted(need & used)} features but carries no dataset statistics: "
                             "pass dataset_stats= (the reference would fail on its infinite placeholder buffers)")
        if language_tokenizer is None:
            try:
                from transformers import AutoTokenizer
                language_tokenizer = AutoTokenizer.from_pretrained("google/paligemma-3b-pt-224", local_files_only=True)
            except Exception:
                language_tokenizer = None  # callers may pass pre-tokenised `lang_tokens` / `lang_masks`
        return cls(config, state_dict=sd, language_tokenizer=language_tokenizer, **kw)

    def reset(self):
        """This should be called whenever the environment is reset."""
        self._action_queue = deque([], maxlen=self.config.n_action_steps)

    @torch.no_grad()
    def select_action(self, batch: dict, noise: Tensor | None = None, noise_std: float = 1.0):
        if self.config.adapt_to_pi_aloha:
            raise NotImplementedError("adapt_to_pi_aloha is not part of the CoVer path")
        batch = self.normalize_inputs(batch)
        if len(self._action_queue) == 0:
            images, img_masks = self.prepare_images(batch)
            state = self.prepare_state(batch)
            lang_tokens, lang_masks = self.prepare_language(batch)
            actions = self.model.sample_actions(images, img_masks, lang_tokens, lang_masks, state, noise=noise,
                                                noise_std=noise_std)
            actions = actions[:, : self.config.n_action_steps]
            original_action_dim = self.config.action_feature.shape[0]
            actions = actions[:, :, :original_action_dim]
            actions = self.unnormalize_outputs({"action": actions})["action"]
            self._action_queue.extend(actions.transpose(0, 1))
        return self._action_queue

    def prepare_images(self, batch):
        images, img_masks = [], []
        present = [k for k in self.config.image_features if k in batch]
        if len(present) == 0:
            raise ValueError(f"All image features are missing from the batch. At least one expected. "
                             f"(batch: {batch.keys()}) (image_features:{self.config.image_features})")
        for key in present:
            img = batch[key]
            if self.config.resize_imgs_with_padding is not None:
                img = resize_with_pad(img, *self.config.resize_imgs_with_padding, pad_value=0)
            mask = torch.ones(img.shape[0], dtype=torch.bool, device=img.device)
            images.append(img)
            img_masks.append(mask)
        # Missing keys: the reference appends up to config.empty_cameras images of -1 with an all-False mask (:377-385).
        # Those tokens are masked as keys, do not advance the position ids and are never read, so they are not created
        # here at all (exact; sample_actions also drops cameras whose mask is all False).
        return images, img_masks

    def prepare_state(self, batch):
        return pad_vector(batch[OBS_ROBOT], self.config.max_state_dim)

    def prepare_language(self, batch):
        device = batch[OBS_ROBOT].device
        if "lang_tokens" in batch:  # pre-tokenised (tests / hosts without the PaliGemma tokenizer)
            self.model.lang_len_hint = None  # a bound left by an earlier tokenised call must not truncate these prompts
            return batch["lang_tokens"].to(device), batch["lang_masks"].to(device=device, dtype=torch.bool)
        if self.language_tokenizer is None:
            raise RuntimeError("no language tokenizer available offline: pass `lang_tokens` / `lang_masks` in the batch")
        tasks = [t if t.endswith("\n") else f"{t}\n" for t in batch["task"]]
        # per-task prompt cache: the prompts of an episode are static between the instruction
        # swaps of run_simpler_eval_with_openpi.py:409, so the tokenizer run, the H2D copy and the length bound of the
        # last few prompt sets are kept (device tensors; same values the tokenizer would return again)
        key = (tuple(tasks), str(device), id(self.language_tokenizer))
        hit = self._prompt_cache.get(key)
        if hit is None:
            tok = self.language_tokenizer.__call__(tasks, padding="max_length", padding_side="right",
                                                   max_length=self.config.tokenizer_max_length, return_tensors="pt",
                                                   truncation=True)
            hit = (tok["input_ids"].to(device=device), tok["attention_mask"].to(device=device, dtype=torch.bool),
                   int(tok["attention_mask"].sum(dim=1).max()))  # host tensor: no device sync
            if len(self._prompt_cache) >= 8:
                self._prompt_cache.pop(next(iter(self._prompt_cache)))
            self._prompt_cache[key] = hit
        self.model.lang_len_hint = hit[2]
        return hit[0], hit[1]
