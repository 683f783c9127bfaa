"""Python handle on the engine-level C ABI (cvb_create / cvb_bind_weight / cvb_finalize / cvb_pi0_sample /
cvb_verifier_score ...).  PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields

import torch

from . import _lib

_CFG_FIELDS = [
    "struct_size",
    "vis_layers", "vis_width", "vis_heads", "vis_mlp", "vis_patch", "vis_image",
    "layers", "lm_width", "lm_mlp", "heads", "head_dim", "ex_width", "ex_mlp", "vocab",
    "max_state_dim", "max_action_dim", "chunk_size", "max_lang_len", "num_steps",
    "max_rephrases", "max_samples",
    "vf_image", "vf_patch", "vf_width", "vf_layers", "vf_heads", "vf_mlp",
    "vf_text_layers", "vf_text_ctx", "vf_vocab",
    "vf_members", "vf_embed", "vf_pool_heads", "vf_pool_layers", "vf_traj_layers", "vf_traj_ff",
    "vf_history", "vf_action_dim",
    "use_cuda_graph",
    "max_observations",
    "num_cameras",
]


class CvbConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in _CFG_FIELDS]


@dataclass
class EngineConfig:
    """Mirror of cvb_config (include/coverb200.h).  Defaults = the Bridge pi0 checkpoint
    (INT-ACT/config/models/pi0_finetune_bridge.json) and the CoVer-BridgeV2 verifier."""
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
    max_rephrases: int = 8
    max_samples: int = 5
    vf_image: int = 384
    vf_patch: int = 16
    vf_width: int = 1024
    vf_layers: int = 24
    vf_heads: int = 16
    vf_mlp: int = 4096
    vf_text_layers: int = 24
    vf_text_ctx: int = 64
    vf_vocab: int = 256000
    vf_members: int = 0
    vf_embed: int = 512
    vf_pool_heads: int = 8
    vf_pool_layers: int = 4
    vf_traj_layers: int = 4
    vf_traj_ff: int = 1024
    vf_history: int = 10
    vf_action_dim: int = 7
    use_cuda_graph: int = 1
    max_observations: int = 1
    num_cameras: int = 1

    def to_c(self) -> CvbConfig:
        c = CvbConfig()
        c.struct_size = C.sizeof(CvbConfig)
        for f in fields(self):
            setattr(c, f.name, int(getattr(self, f.name)))
        return c

    @property
    def n_img_tokens(self) -> int:
        return (self.vis_image // self.vis_patch) ** 2


_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.int64: 2, torch.int32: 3, torch.uint8: 4}


class Engine:
    """Owns one cvb_handle on the current CUDA device."""

    def __init__(self, cfg: EngineConfig, device: str | torch.device = "cuda"):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CvbError("coverb200 has no CPU path: the engine needs a CUDA device")
        self._declare()
        self._h = C.c_void_p()
        ccfg = cfg.to_c()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cvb_create(C.byref(ccfg), C.byref(self._h)))
        self._keep = {}  # tensors whose memory the handle borrows
        self.finalized = False
        # bumped by every call that rewrites the verifier's image/text context, so a host-side cache of "the context is
        # still (image, instruction) X" (EfficientEnsembleMerged) notices another caller's write
        self.ctx_generation = 0

    def _declare(self):
        L = self.lib
        L.cvb_create.argtypes = [C.POINTER(CvbConfig), C.POINTER(C.c_void_p)]
        L.cvb_destroy.argtypes = [C.c_void_p]
        L.cvb_destroy.restype = None
        L.cvb_bind_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.cvb_required_weight_count.argtypes = [C.c_void_p]
        L.cvb_required_weight_name.argtypes = [C.c_void_p, C.c_int]
        L.cvb_required_weight_name.restype = C.c_char_p
        L.cvb_required_weight_dtype.argtypes = [C.c_void_p, C.c_int]
        L.cvb_finalize.argtypes = [C.c_void_p, C.c_void_p]
        L.cvb_pi0_sample.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.cvb_pi0_set_lang_len_hint.argtypes = [C.c_void_p, C.c_int]
        L.cvb_pi0_set_active_cameras.argtypes = [C.c_void_p, C.c_int]
        L.cvb_verifier_hold_text.argtypes = [C.c_void_p, C.c_int]
        L.cvb_pi0_run_phase.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.cvb_debug_copy.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.cvb_debug_copy.restype = C.c_int64
        L.cvb_verifier_score.argtypes = [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
        L.cvb_verifier_set_features.argtypes = [C.c_void_p] * 4
        L.cvb_verifier_context.argtypes = [C.c_void_p] * 4
        L.cvb_cover_step.argtypes = ([C.c_void_p] * 6 + [C.c_int, C.c_int] + [C.c_void_p] * 2 +
                                     [C.POINTER(C.c_double)] * 2 + [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7)
        L.cvb_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cvb_pi0_sample_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.cvb_pi0_run_phase_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.cvb_cover_step_batch.argtypes = ([C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int] + [C.c_void_p] * 2 +
                                           [C.POINTER(C.c_double)] * 2 + [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7)

    # ------------------------------------------------------------------ weights
    def required_weights(self) -> list[str]:
        n = self.lib.cvb_required_weight_count(self._h)
        return [self.lib.cvb_required_weight_name(self._h, i).decode() for i in range(n)]

    def bind(self, key: str, t: torch.Tensor):
        if t.device != self.device and not (t.is_cuda and self.device.index is None):
            t = t.to(self.device)
        t = t.contiguous()
        if t.dtype not in _DTYPES:
            raise _lib.CvbError(f"unsupported dtype for {key}: {t.dtype}")
        shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
        _lib.check(self.lib.cvb_bind_weight(self._h, key.encode(), _lib.ptr(t), _DTYPES[t.dtype], t.dim(), shape))
        self._keep[key] = t

    def required_dtypes(self) -> dict:
        """canonical weight name -> torch dtype the engine expects (bf16 / fp32)."""
        inv = {v: k for k, v in _DTYPES.items()}
        n = self.lib.cvb_required_weight_count(self._h)
        return {self.lib.cvb_required_weight_name(self._h, i).decode(): inv[self.lib.cvb_required_weight_dtype(self._h, i)]
                for i in range(n)}

    def load_state_dict(self, sd: dict, strict_unused: bool = False):
        """Bind a reference state dict (names unchanged).  Like the reference's load_state_dict -
        which copies into parameters already cast by to_bfloat16_like_physical_intelligence
        (paligemma_with_expert.py:216-227) - the STORED dtype does not matter: every tensor is cast to the dtype the
        engine expects.  Tensors the sampling path never reads (lm_head, the tied embedding copy, normalisation
        buffers, ...) are skipped instead of being copied to the GPU; strict_unused=True raises on them."""
        from .synthetic import canonical_key
        want = self.required_dtypes()
        unused = []
        for k, v in sd.items():
            if not isinstance(v, torch.Tensor):
                continue
            ck = canonical_key(k)
            if ck not in want:
                unused.append(k)
                continue
            self.bind(k, v.to(device=self.device, dtype=want[ck]))
        if strict_unused and unused:
            raise _lib.CvbError(f"state dict has tensors the engine does not use: {unused[:5]} ...")
        return unused

    def finalize(self, release_repacked: bool = True):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cvb_finalize(self._h, _lib.stream_ptr()))
            torch.cuda.synchronize()
        self.finalized = True
        if release_repacked:
            # q/k/v, gate/up and the patch embedding were repacked into library-owned memory
            for k in list(self._keep):
                if any(s in k for s in ("q_proj.weight", "k_proj.weight", "v_proj.weight", "gate_proj.weight",
                                        "up_proj.weight", "patch_embedding.weight")) and "verifier." not in k:
                    del self._keep[k]
                elif ("q_proj.bias" in k or "k_proj.bias" in k or "v_proj.bias" in k) and "verifier." not in k:
                    del self._keep[k]

    # ------------------------------------------------------------------ pi0
    def set_lang_len_hint(self, max_valid_tokens: int | None):
        """Bound on valid language tokens per prompt (None / 0 = none).  The host that tokenised the prompts knows it
        without a device sync; the prefix then skips the right-padding rows (exact: SURVEY.md F11)."""
        n = int(max_valid_tokens or 0)
        if n != getattr(self, "_lang_hint", 0):
            _lib.check(self.lib.cvb_pi0_set_lang_len_hint(self._h, n))
            self._lang_hint = n

    def verifier_hold_text(self, hold: bool):
        """Per-task prompt cache: while held, context computations reuse the resident text-tower features (the caller
        vouches that the instruction tokens did not change; include/coverb200.h cvb_verifier_hold_text)."""
        if bool(hold) != getattr(self, "_hold_text", False):
            _lib.check(self.lib.cvb_verifier_hold_text(self._h, 1 if hold else 0))
            self._hold_text = bool(hold)

    def set_active_cameras(self, cameras: int | None):
        """Cameras per observation in the following calls (None / 0 = cfg.num_cameras); masked cameras are dropped by the
        caller, which is exact (include/coverb200.h)."""
        n = int(cameras or 0)
        if n == max(1, self.cfg.num_cameras):
            n = 0
        if n != getattr(self, "_active_cams", 0):
            _lib.check(self.lib.cvb_pi0_set_active_cameras(self._h, n))
            self._active_cams = n

    def _image_numel(self):
        cams = getattr(self, "_active_cams", 0) or max(1, self.cfg.num_cameras)
        return cams * 3 * self.cfg.vis_image ** 2

    def pi0_sample(self, image, lang_tokens, lang_len, state, noise, K: int, out=None, lang_len_max: int | None = None):
        """image f32 [3,H,W] ([C,3,H,W] with C cameras); lang_tokens i64 [R,L]; lang_len i32 [R]; state f32
        [max_state_dim]; noise f32 [R*K, chunk, max_action_dim] -> actions f32 (same shape).  Asynchronous."""
        cfg = self.cfg
        R = lang_tokens.shape[0]
        assert image.dtype == torch.float32 and image.numel() == self._image_numel()
        assert lang_tokens.dtype == torch.int64 and lang_tokens.shape[1] == cfg.max_lang_len
        assert lang_len.dtype == torch.int32 and lang_len.numel() == R
        assert state.dtype == torch.float32 and state.numel() == cfg.max_state_dim
        assert noise.dtype == torch.float32 and tuple(noise.shape) == (R * K, cfg.chunk_size, cfg.max_action_dim)
        for t in (image, lang_tokens, lang_len, state, noise):
            assert t.is_cuda and t.is_contiguous()
        if out is None:
            out = torch.empty_like(noise)
        self.set_lang_len_hint(lang_len_max)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cvb_pi0_sample(self._h, _lib.ptr(image), _lib.ptr(lang_tokens), _lib.ptr(lang_len),
                                               _lib.ptr(state), _lib.ptr(noise), R, K, _lib.ptr(out),
                                               _lib.stream_ptr()))
        return out

    def pi0_run_phase(self, phase: int, R: int, K: int, B: int = 1):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cvb_pi0_run_phase_batch(self._h, phase, B, R, K, _lib.stream_ptr()))

    def pi0_sample_batch(self, images, lang_tokens, lang_len, states, noise, K: int, out=None,
                         lang_len_max: int | None = None):
        """B observations in one pass (cvb_pi0_sample_batch): images f32 [B,3,H,W]; lang_tokens i64 [B,R,L]; lang_len
        i32 [B,R]; states f32 [B,max_state_dim]; noise f32 [B,R*K,chunk,max_action_dim] -> actions (same shape)."""
        cfg = self.cfg
        B, R = lang_tokens.shape[0], lang_tokens.shape[1]
        assert B <= cfg.max_observations, "engine was built with a smaller max_observations"
        assert images.dtype == torch.float32 and images.shape[0] == B and images.numel() == B * self._image_numel()
        assert lang_tokens.dtype == torch.int64 and lang_tokens.shape[2] == cfg.max_lang_len
        assert lang_len.dtype == torch.int32 and tuple(lang_len.shape) == (B, R)
        assert states.dtype == torch.float32 and tuple(states.shape) == (B, cfg.max_state_dim)
        assert noise.dtype == torch.float32 and tuple(noise.shape) == (B, R * K, cfg.chunk_size, cfg.max_action_dim)
        for t in (images, lang_tokens, lang_len, states, noise):
            assert t.is_cuda and t.is_contiguous()
        if out is None:
            out = torch.empty_like(noise)
        self.set_lang_len_hint(lang_len_max)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cvb_pi0_sample_batch(self._h, B, _lib.ptr(images), _lib.ptr(lang_tokens), _lib.ptr(lang_len),
                                                     _lib.ptr(states), _lib.ptr(noise), R, K, _lib.ptr(out),
                                                     _lib.stream_ptr()))
        return out

    def cover_step_batch(self, images, lang_tokens, lang_len, states, noise, K: int, vf_images, vf_tokens, p01, p99,
                         past=None, n_future: int | None = None, lang_len_max: int | None = None, hold_text: bool = False):
        """B whole decisions in one graph (cvb_cover_step_batch).  Shapes as pi0_sample_batch plus vf_images f32
        [B,3,S,S], vf_tokens i64 [B,ctx], past f32 [B,num_past,7] or None.  Returns device tensors (actions
        [B,N,chunk,A], traj [B,N,H,7], scores [B,N], group_mean [B,R], best_idx i32 [B], best_score [B])."""
        cfg = self.cfg
        self.verifier_hold_text(hold_text)  # per-task prompt cache: explicit per call, never inherited from another caller
        B, R = lang_tokens.shape[0], lang_tokens.shape[1]
        N = R * K
        assert B <= cfg.max_observations, "engine was built with a smaller max_observations"
        assert tuple(noise.shape) == (B, N, cfg.chunk_size, cfg.max_action_dim) and noise.dtype == torch.float32
        assert images.shape[0] == B and images.numel() == B * self._image_numel() and tuple(states.shape) == (B, cfg.max_state_dim)
        assert tuple(vf_images.shape) == (B, 3, cfg.vf_image, cfg.vf_image) and tuple(vf_tokens.shape) == (B, cfg.vf_text_ctx)
        assert tuple(lang_len.shape) == (B, R) and lang_tokens.shape[2] == cfg.max_lang_len
        for t in (images, lang_tokens, lang_len, states, noise, vf_images, vf_tokens):
            assert t.is_cuda and t.is_contiguous()
        assert lang_tokens.dtype == torch.int64 and lang_len.dtype == torch.int32 and vf_tokens.dtype == torch.int64
        num_past = 0 if past is None else int(past.shape[1])
        if past is not None:
            assert past.dtype == torch.float32 and past.is_cuda and past.is_contiguous() and tuple(past.shape) == (B, num_past, 7)
        n_future = n_future or cfg.chunk_size
        dev = self.device
        actions = torch.empty_like(noise)
        traj = torch.empty(B, N, cfg.vf_history, 7, dtype=torch.float32, device=dev)
        scores = torch.empty(B, N, dtype=torch.float32, device=dev)
        gmean = torch.empty(B, R, dtype=torch.float32, device=dev)
        bidx = torch.zeros(B, dtype=torch.int32, device=dev)
        bscore = torch.zeros(B, dtype=torch.float32, device=dev)
        a = (C.c_double * 6)(*p01)
        b = (C.c_double * 6)(*p99)
        self.set_lang_len_hint(lang_len_max)
        self.ctx_generation += 1
        with torch.cuda.device(dev):
            _lib.check(self.lib.cvb_cover_step_batch(self._h, B, _lib.ptr(images), _lib.ptr(lang_tokens), _lib.ptr(lang_len),
                                                     _lib.ptr(states), _lib.ptr(noise), R, K, _lib.ptr(vf_images),
                                                     _lib.ptr(vf_tokens), a, b, _lib.ptr(past), num_past, n_future,
                                                     _lib.ptr(actions), _lib.ptr(traj), _lib.ptr(scores), _lib.ptr(gmean),
                                                     _lib.ptr(bidx), _lib.ptr(bscore), _lib.stream_ptr()))
        return actions, traj, scores, gmean, bidx, bscore

    # ------------------------------------------------------------------ verifier
    def verifier_score(self, image, text_tokens, traj, R: int, K: int, recompute_context: bool = True,
                       hold_text: bool = False):
        """image f32 [3,S,S] (or None to reuse the context); text_tokens i64 [ctx]; traj f32 [N,H,A] left-padded
        with -5.  Returns device tensors (scores [N], group_mean [R], best_idx i32 [1], best_score [1]).
        R == 0: scores only."""
        cfg = self.cfg
        self.verifier_hold_text(hold_text)  # per-task prompt cache: explicit per call, never inherited from another caller
        N = traj.shape[0]
        assert traj.dtype == torch.float32 and tuple(traj.shape[1:]) == (cfg.vf_history, cfg.vf_action_dim)
        assert traj.is_cuda and traj.is_contiguous()
        if image is not None:
            assert image.dtype == torch.float32 and image.numel() == 3 * cfg.vf_image ** 2 and image.is_contiguous()
            assert text_tokens.dtype == torch.int64 and text_tokens.numel() == cfg.vf_text_ctx
        scores = torch.empty(N, dtype=torch.float32, device=self.device)
        gmean = torch.empty(max(R, 1), dtype=torch.float32, device=self.device)
        bidx = torch.zeros(1, dtype=torch.int32, device=self.device)
        bscore = torch.zeros(1, dtype=torch.float32, device=self.device)
        if recompute_context:
            self.ctx_generation += 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cvb_verifier_score(self._h, _lib.ptr(image), _lib.ptr(text_tokens), _lib.ptr(traj),
                                                   N, R, K, _lib.ptr(scores), _lib.ptr(gmean), _lib.ptr(bidx),
                                                   _lib.ptr(bscore), int(recompute_context), _lib.stream_ptr()))
        return scores, gmean, bidx, bscore

    def cover_step(self, image, lang_tokens, lang_len, state, noise, K: int, vf_image, vf_tokens, p01, p99, past=None,
                   n_future: int | None = None, lang_len_max: int | None = None, hold_text: bool = False):
        """One whole decision (cvb_cover_step): sample -> format -> score -> select in one graph.  Returns device
        tensors (actions [N,chunk,A], traj [N,H,7], scores [N], group_mean [R], best_idx i32 [1], best_score [1])."""
        cfg = self.cfg
        self.verifier_hold_text(hold_text)  # per-task prompt cache: explicit per call, never inherited from another caller
        R = lang_tokens.shape[0]
        N = R * K
        assert tuple(noise.shape) == (N, cfg.chunk_size, cfg.max_action_dim) and noise.dtype == torch.float32
        for t in (image, lang_tokens, lang_len, state, noise, vf_image, vf_tokens):
            assert t.is_cuda and t.is_contiguous()
        assert lang_tokens.dtype == torch.int64 and lang_len.dtype == torch.int32 and vf_tokens.dtype == torch.int64
        num_past = 0 if past is None else int(past.shape[0])
        if past is not None:
            assert past.dtype == torch.float32 and past.is_cuda and past.is_contiguous() and past.shape[1] == 7
        n_future = n_future or cfg.chunk_size
        actions = torch.empty_like(noise)
        traj = torch.empty(N, cfg.vf_history, 7, dtype=torch.float32, device=self.device)
        scores = torch.empty(N, dtype=torch.float32, device=self.device)
        gmean = torch.empty(R, dtype=torch.float32, device=self.device)
        bidx = torch.zeros(1, dtype=torch.int32, device=self.device)
        bscore = torch.zeros(1, dtype=torch.float32, device=self.device)
        a = (C.c_double * 6)(*p01)
        b = (C.c_double * 6)(*p99)
        self.set_lang_len_hint(lang_len_max)
        self.ctx_generation += 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cvb_cover_step(self._h, _lib.ptr(image), _lib.ptr(lang_tokens), _lib.ptr(lang_len),
                                               _lib.ptr(state), _lib.ptr(noise), R, K, _lib.ptr(vf_image),
                                               _lib.ptr(vf_tokens), a, b, _lib.ptr(past), num_past, n_future,
                                               _lib.ptr(actions), _lib.ptr(traj), _lib.ptr(scores), _lib.ptr(gmean),
                                               _lib.ptr(bidx), _lib.ptr(bscore), _lib.stream_ptr()))
        return actions, traj, scores, gmean, bidx, bscore

    def verifier_context(self, image, text_tokens, hold_text: bool = False):
        """Image/text side only (trunk + image-text heads) on the current stream; pair with
        verifier_score(None, None, traj, ..., recompute_context=False)."""
        cfg = self.cfg
        self.verifier_hold_text(hold_text)  # per-task prompt cache: explicit per call, never inherited from another caller
        assert image.dtype == torch.float32 and image.numel() == 3 * cfg.vf_image ** 2 and image.is_contiguous()
        assert text_tokens.dtype == torch.int64 and text_tokens.numel() == cfg.vf_text_ctx
        self.ctx_generation += 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cvb_verifier_context(self._h, _lib.ptr(image), _lib.ptr(text_tokens), _lib.stream_ptr()))

    def verifier_set_features(self, patch, text):
        assert patch.dtype == torch.float32 and text.dtype == torch.float32 and patch.is_cuda and text.is_cuda
        self.ctx_generation += 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cvb_verifier_set_features(self._h, _lib.ptr(patch.contiguous()),
                                                          _lib.ptr(text.contiguous()), _lib.stream_ptr()))

    def select(self, scores, R: int, K: int):
        """group-mean -> argmax group -> argmax inside it over a (gathered) fp32 score vector."""
        assert scores.dtype == torch.float32 and scores.is_cuda and scores.numel() == R * K
        gmean = torch.empty(R, dtype=torch.float32, device=scores.device)
        bidx = torch.zeros(1, dtype=torch.int32, device=scores.device)
        bscore = torch.zeros(1, dtype=torch.float32, device=scores.device)
        with torch.cuda.device(scores.device):
            _lib.check(self.lib.cvb_select(_lib.ptr(scores.contiguous()), R, K, _lib.ptr(gmean), _lib.ptr(bidx),
                                           _lib.ptr(bscore), _lib.stream_ptr()))
        return gmean, bidx, bscore

    def debug(self, name: str, shape, dtype) -> torch.Tensor:
        out = torch.zeros(shape, dtype=dtype, device=self.device)
        n = self.lib.cvb_debug_copy(self._h, name.encode(), _lib.ptr(out), out.numel() * out.element_size(),
                                    _lib.stream_ptr())
        if n < 0:
            _lib.check(int(n))
        torch.cuda.synchronize()
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.cvb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
