"""Drop-in host surface for the reference CoVer verifier, running on the coverb200 CUDA engine.

Mirrors bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:
  EfficientEnsembleMerged.__init__                            :25-186 (checkpoint layout, default config :41-48)
  .compute_max_similarity_scores_batch                        :309-454
  .predict / .fuse_embeddings / .extract_shared_features      :295-307, :249-293, :188-192
Same argument meaning and return types; the arithmetic runs in cvb_verifier_score (include/coverb200.h).
What is deliberately different: only (images[0], instructions[0]) is encoded - the
reference consumes row 0 of its similarity matrix only (:422-425) - and an unchanged (image, instruction)
pair is not re-encoded on the second call of a decision (the 1-candidate gate, then the N candidates).
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

from ..engine import Engine, EngineConfig

_VF_DEFAULT = dict(vf_image=384, vf_patch=16, vf_width=1024, vf_layers=24, vf_heads=16, vf_mlp=4096,
                   vf_text_layers=24, vf_text_ctx=64, vf_vocab=256000, vf_embed=512, vf_pool_heads=8,
                   vf_pool_layers=4, vf_traj_layers=4, vf_traj_ff=1024, vf_history=10, vf_action_dim=7)


def default_preprocess(image_size: int):
    """open_clip's SigLIP image transform: squash-resize (bicubic) -> [0,1] -> (x - 0.5) / 0.5."""
    def _pre(image):
        from PIL import Image
        if isinstance(image, np.ndarray):
            image = Image.fromarray(image.astype("uint8"))
        image = image.convert("RGB").resize((image_size, image_size), Image.BICUBIC)
        x = torch.from_numpy(np.asarray(image).copy()).permute(2, 0, 1).to(torch.float32) / 255.0
        return (x - 0.5) / 0.5
    return _pre


class _SiglipShim:
    def __init__(self, context_length):
        self.context_length = context_length


class EfficientEnsembleMerged:
    def __init__(self, merged_checkpoint_path=None, device="cuda", *, ensemble_components=None,
                 trunk_state_dict=None, tokenizer=None, preprocess=None, vf_config: dict | None = None,
                 max_candidates: int = 40, engine: Engine | None = None):
        """merged_checkpoint_path: the reference's merged .pt (dict with 'ensemble_components': list of
        per-member dicts of state dicts, :94-160).  trunk_state_dict: `siglip_model.state_dict()` of the
        open_clip model 'hf-hub:timm/ViT-L-16-SigLIP2-384' (open_clip / timm / the hub are not available
        offline, so the trunk weights and the tokenizer / preprocess callables are injected)."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("coverb200 has no CPU path")
        meta = {}
        if merged_checkpoint_path is not None:
            ckpt = torch.load(merged_checkpoint_path, map_location="cpu", weights_only=False)
            ensemble_components = ckpt["ensemble_components"]
            meta = ckpt
        if engine is None and ensemble_components is None:
            raise ValueError("need a merged checkpoint path, ensemble_components, or a finalized Engine")
        # weights-only checkpoints default to the CoVer-BridgeV2 configuration (:41-48)
        self.backbone = meta.get("backbone", "hf-hub:timm/ViT-L-16-SigLIP2-384")
        self.use_transformer = meta.get("use_transformer", True)
        self.history_length = meta.get("history_length", 10)
        self.action_dim = meta.get("action_dim", 7)
        if engine is None:
            cfgd = dict(_VF_DEFAULT)
            if not self.use_transformer:
                # MLP action encoder (:161-171): the engine takes vf_traj_layers = 0 and the hidden width (512 in the
                # reference; read from the checkpoint) in vf_traj_ff
                w0 = ensemble_components[0]["complex_action_encoder"]["0.weight"]
                cfgd["vf_traj_layers"], cfgd["vf_traj_ff"] = 0, int(w0.shape[0])
            cfgd.update(vf_config or {})
            cfgd["vf_history"], cfgd["vf_action_dim"] = self.history_length, self.action_dim
            self.num_models = len(ensemble_components)
            ecfg = EngineConfig(layers=0, vis_layers=0, vf_members=self.num_models, max_rephrases=max_candidates,
                                max_samples=1, **cfgd)
            engine = Engine(ecfg, device=self.device)
            if trunk_state_dict is None:
                # exactly what the reference does (efficient_ensemble_merged.py:57-69): the shared trunk, its image
                # transform and its tokenizer come from open_clip, not from the checkpoint
                try:
                    import open_clip
                except ImportError as e:
                    raise RuntimeError("the SigLIP2 trunk is loaded with open_clip.create_model_from_pretrained("
                                       f"{self.backbone!r}) like the reference does, but open_clip is not installed: "
                                       "install open_clip_torch + timm, or pass trunk_state_dict= (and tokenizer= / "
                                       "preprocess=)") from e
                siglip_model, oc_preprocess = open_clip.create_model_from_pretrained(self.backbone)
                trunk_state_dict = siglip_model.state_dict()
                if preprocess is None:
                    preprocess = oc_preprocess
                if tokenizer is None:
                    tokenizer = open_clip.get_tokenizer(self.backbone)
                del siglip_model
            for k, v in trunk_state_dict.items():
                k = k if k.startswith("verifier.trunk.") else "verifier.trunk." + k
                if k in engine_required(engine):
                    engine.bind(k, v.to(torch.bfloat16))
            for m, comp in enumerate(ensemble_components):
                for cname, sd in comp.items():
                    if not isinstance(sd, dict):
                        continue  # action_padding_value (python float, -5.0); None for the unused action encoder
                    for k, v in sd.items():
                        key = f"verifier.{m}.{cname}.{k}"
                        if key in engine_required(engine):
                            engine.bind(key, v.to(torch.float32))
            engine.finalize()
        else:
            self.num_models = engine.cfg.vf_members
        self.engine = engine
        self.tokenizer = tokenizer
        self.preprocess = preprocess or default_preprocess(engine.cfg.vf_image)
        # RGB uint8 frames (PIL images / arrays) take the device-side transform unless a custom preprocess was injected
        self._device_preprocess = preprocess is None
        self.siglip_model = _SiglipShim(engine.cfg.vf_text_ctx)
        self._ctx_key = None

    # ------------------------------------------------------------------------------------------
    def _tokens(self, instruction) -> torch.Tensor:
        if isinstance(instruction, str):
            if self.tokenizer is None:
                raise RuntimeError("no tokenizer available offline: pass pre-tokenised instructions (int64 tensors)")
            t = self.tokenizer([instruction], context_length=self.siglip_model.context_length)
        else:
            t = instruction
        t = torch.as_tensor(t)
        if t.ndim == 2:
            t = t[0]
        return t.to(torch.int64)

    def _set_context(self, image, instruction):
        """Stage (image, instruction) on the device; returns (img, tok, recompute)."""
        if isinstance(image, torch.Tensor):
            img = image.to(torch.float32)
            raw = img.cpu().numpy().tobytes() if not img.is_cuda else None
        else:
            arr = np.asarray(image)
            raw = arr.tobytes()
            if self._device_preprocess and arr.dtype == np.uint8 and arr.ndim == 3 and arr.shape[2] == 3:
                # open_clip's transform on the device (bit-exact with PIL's bicubic resize): one uint8 H2D copy
                from .. import preprocess as _pp
                img = _pp.verifier_image(torch.from_numpy(np.array(arr)).to(self.device),
                                         self.engine.cfg.vf_image)[0]
            else:
                img = self.preprocess(image)
        tok = self._tokens(instruction)
        key = None
        if raw is not None:
            key = (hashlib.blake2b(raw, digest_size=16).digest(), len(raw), tuple(tok.tolist()))
        # reuse only if the pair is unchanged AND nobody else (CoverStep, verifier_context, ...) rewrote the engine's
        # context since this wrapper last did (Engine.ctx_generation)
        recompute = key is None or key != self._ctx_key or self.engine.ctx_generation != getattr(self, "_ctx_gen", -1)
        self._ctx_key = key
        if not recompute:
            return None, None, False
        # per-task prompt cache: the image changes every tick, the instruction only at the swaps of
        # run_simpler_eval_with_openpi.py:409 - while the tokens are the ones this wrapper encoded last and nobody else
        # rewrote the engine's context, the text tower is skipped (cvb_verifier_hold_text)
        tkey = tuple(tok.tolist())
        self._hold_text = (tkey == getattr(self, "_text_key", None) and
                           self.engine.ctx_generation == getattr(self, "_ctx_gen", -1))
        self._text_key = tkey
        self._ctx_gen = self.engine.ctx_generation + 1  # the verifier_score(recompute_context=True) that follows bumps it
        return img.to(self.device).contiguous(), tok.to(self.device).contiguous(), True

    def _pad(self, all_action_histories) -> torch.Tensor:
        # :379-390
        H = self.engine.cfg.vf_history
        out = np.full((len(all_action_histories), H, self.engine.cfg.vf_action_dim), -5.0, dtype=np.float64)
        for i, ah in enumerate(all_action_histories):
            ah = np.asarray(ah)
            if ah.ndim == 1:
                ah = ah[:, None]
            if len(ah) > H:
                raise ValueError(f"action history longer than {H} steps")
            out[i, H - len(ah):] = ah
        return torch.tensor(out, dtype=torch.float32)

    def compute_max_similarity_scores_batch(self, images, instructions, all_action_histories,
                                            cfg_repeat_language_instructions=1):
        num_actions = len(all_action_histories)
        group_size = cfg_repeat_language_instructions
        if num_actions % group_size != 0:
            raise ValueError("number of action histories must be a multiple of cfg_repeat_language_instructions")
        num_groups = num_actions // group_size
        img, tok, recompute = self._set_context(images[0], instructions[0])
        traj = self._pad(all_action_histories).to(self.device, non_blocking=True)
        scores, gmean, bidx, bscore = self.engine.verifier_score(img, tok, traj, num_groups, group_size,
                                                                 recompute_context=recompute,
                                                                 hold_text=recompute and self._hold_text)
        out = torch.stack([bscore[0], bidx[0].to(torch.float32)]).cpu()  # the one D2H sync (:439 does .item())
        max_score, gidx = float(out[0]), int(out[1])
        all_same = len(set(instructions)) == 1 if isinstance(instructions[0], str) else False
        if all_same and len(images) > 1:
            max_instruction = instructions[0]
        else:
            max_instruction = instructions[min((gidx // group_size) * group_size, len(instructions) - 1)]
        self.last_scores = scores
        return max_score, max_instruction, all_action_histories[gidx], torch.tensor(gidx, dtype=torch.int64)

    def predict(self, image, instruction, possible_action_histories):
        # :295-307 - diagonal of [N,512]@[512,N] with N identical image-text rows == the score vector
        img, tok, recompute = self._set_context(image, instruction)
        traj = self._pad(possible_action_histories).to(self.device)
        scores, *_ = self.engine.verifier_score(img, tok, traj, 0, 1, recompute_context=recompute,
                                                hold_text=recompute and self._hold_text)
        scores = scores.cpu().numpy()
        idx = int(scores.argmax())
        return possible_action_histories[idx], {str(i): float(scores[i]) for i in range(len(scores))}

    def extract_shared_features(self, img_tensor, text_tokens):
        """(patch_features [1,Np,W], text_features [1,ctx,W]) fp32, L2-normalised (:188-192)."""
        cfg = self.engine.cfg
        dummy = torch.full((1, cfg.vf_history, cfg.vf_action_dim), 0.0, device=self.device)
        self._ctx_key = self._text_key = None
        self.engine.verifier_score(img_tensor.reshape(3, cfg.vf_image, cfg.vf_image).to(self.device, torch.float32).contiguous(),
                                   text_tokens.reshape(-1).to(self.device, torch.int64).contiguous(), dummy, 0, 1)
        Np = (cfg.vf_image // cfg.vf_patch) ** 2
        patch = self.engine.debug("vf_patch_features", (1, Np, cfg.vf_width), torch.float32)
        text = self.engine.debug("vf_text_features", (1, cfg.vf_text_ctx, cfg.vf_width), torch.float32)
        return patch, text

    def fuse_embeddings(self, image, instruction, action_histories):
        """(fused_image_text [N,512] (identical rows), fused_action [N,512]) as in :249-293."""
        cfg = self.engine.cfg
        img, tok, recompute = self._set_context(image, instruction)
        traj = self._pad(action_histories).to(self.device)
        N = traj.shape[0]
        self.engine.verifier_score(img, tok, traj, 0, 1, recompute_context=recompute)
        it = self.engine.debug("vf_it_emb", (cfg.vf_members, cfg.vf_embed), torch.float32)
        act = self.engine.debug("vf_act_emb", (cfg.vf_members, N, cfg.vf_embed), torch.float32)
        fit = it.mean(dim=0, keepdim=True)
        fit = fit / fit.norm(dim=-1, keepdim=True)
        fact = act.mean(dim=0)
        fact = fact / fact.norm(dim=-1, keepdim=True)
        return fit.expand(N, -1), fact


def engine_required(engine: Engine) -> set:
    cache = getattr(engine, "_required_cache", None)
    if cache is None:
        cache = set(engine.required_weights())
        engine._required_cache = cache
    return cache
