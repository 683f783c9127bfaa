"""Device-side observation pre-processing.

`policy_image` replaces BridgeSimplerAdapter.preprocess (INT-ACT/src/experiments/env_adapters/simpler.py:43-65): the uint8
simulator frame is copied to the device once and resized with cv2's INTER_LANCZOS4 arithmetic (bit-exact), then scaled to
[-1, 1] exactly as src/utils/pipeline.py:34-69 does."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def policy_image(frame_u8_hwc: torch.Tensor, size: int = 224, return_u8: bool = False):
    """frame uint8 [H, W, 3] (CUDA, contiguous) -> float32 [1, 3, size, size] in [-1, 1] (and the uint8 HWC resize)."""
    assert frame_u8_hwc.dtype == torch.uint8 and frame_u8_hwc.is_cuda and frame_u8_hwc.is_contiguous()
    H, W, ch = frame_u8_hwc.shape
    assert ch == 3, "expected an RGB frame"
    lib = _lib.load()
    out = torch.empty(1, 3, size, size, dtype=torch.float32, device=frame_u8_hwc.device)
    u8 = torch.empty(size, size, 3, dtype=torch.uint8, device=frame_u8_hwc.device) if return_u8 else None
    lib.cvb_preprocess_policy_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_void_p]
    with torch.cuda.device(frame_u8_hwc.device):
        _lib.check(lib.cvb_preprocess_policy_image(_lib.ptr(frame_u8_hwc), H, W, size, size, _lib.ptr(u8), _lib.ptr(out),
                                                   _lib.stream_ptr()))
    return (out, u8) if return_u8 else out


def verifier_image(frame_u8_hwc: torch.Tensor, size: int = 384, return_u8: bool = False):
    """open_clip's SigLIP image transform (PIL bicubic squash-resize -> [0, 1] -> (x - 0.5) / 0.5) on the device:
    frame uint8 [H, W, 3] (CUDA, contiguous) -> float32 [1, 3, size, size]; the uint8 resize is bit-exact with Pillow."""
    assert frame_u8_hwc.dtype == torch.uint8 and frame_u8_hwc.is_cuda and frame_u8_hwc.is_contiguous()
    H, W, ch = frame_u8_hwc.shape
    assert ch == 3, "expected an RGB frame"
    lib = _lib.load()
    out = torch.empty(1, 3, size, size, dtype=torch.float32, device=frame_u8_hwc.device)
    u8 = torch.empty(size, size, 3, dtype=torch.uint8, device=frame_u8_hwc.device) if return_u8 else None
    lib.cvb_preprocess_verifier_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                  C.c_void_p]
    with torch.cuda.device(frame_u8_hwc.device):
        _lib.check(lib.cvb_preprocess_verifier_image(_lib.ptr(frame_u8_hwc), H, W, size, size, _lib.ptr(u8), _lib.ptr(out),
                                                     _lib.stream_ptr()))
    return (out, u8) if return_u8 else out


def verifier_frame(frame_u8_hwc: torch.Tensor, size: int = 256) -> torch.Tensor:
    """process_raw_image_to_jpg (eval_utils.py:228-286) on the device: tf.image.resize(frame, (size, size), BILINEAR,
    antialias=True) + uint8 cast.  frame uint8 [H, W, 3] (CUDA, contiguous) -> uint8 [size, size, 3]."""
    assert frame_u8_hwc.dtype == torch.uint8 and frame_u8_hwc.is_cuda and frame_u8_hwc.is_contiguous()
    H, W, ch = frame_u8_hwc.shape
    assert ch == 3, "expected an RGB frame"
    lib = _lib.load()
    out = torch.empty(size, size, 3, dtype=torch.uint8, device=frame_u8_hwc.device)
    scratch = torch.empty(size * W * 3, dtype=torch.float32, device=frame_u8_hwc.device)
    lib.cvb_resize_bilinear_antialias_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                     C.c_void_p]
    with torch.cuda.device(frame_u8_hwc.device):
        _lib.check(lib.cvb_resize_bilinear_antialias_u8(_lib.ptr(frame_u8_hwc), H, W, size, size, _lib.ptr(scratch),
                                                        _lib.ptr(out), _lib.stream_ptr()))
    return out


def verifier_image_from_raw(frame_u8_hwc: torch.Tensor, size: int = 384):
    """The reference's whole verifier-side image path from ONE uint8 H2D copy of the simulator frame: bilinear-antialias
    256^2 (process_raw_image_to_jpg) -> open_clip's PIL-bicubic size^2 transform + normalisation.  -> f32 [1, 3, size, size]."""
    return verifier_image(verifier_frame(frame_u8_hwc, 256), size)
