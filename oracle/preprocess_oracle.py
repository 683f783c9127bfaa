"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the policy-side observation pre-processing.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Follows (reference, /root/reference):
  * BridgeSimplerAdapter.preprocess   INT-ACT/src/experiments/env_adapters/simpler.py:43-65: cv2.resize(obs, (224, 224),
    interpolation=cv2.INTER_LANCZOS4) on the uint8 HWC frame, then process_images(rescale 1/255, mean 0.5, std 0.5)
  * process_images / rescale / normalize   INT-ACT/src/utils/pipeline.py:34-69 (float32 torch arithmetic)
Third-party arithmetic: OpenCV (opencv-python, unpinned in /root/reference/requirements.txt; 4.13.0 in this image).  The
8-bit INTER_LANCZOS4 path of cv::resize is restated from its published algorithm (imgproc/src/resize.cpp: resizeGeneric_
with HResizeLanczos4<uchar,int,short> / VResizeLanczos4<..., FixedPtCast<int,uchar,22>>; interpolateLanczos4; coefficients
quantised to 11 bits with cvRound; replicated borders; no anti-aliasing) and PINNED bit for bit against cv2.resize itself:
tests/test_preprocess.py runs cv2 live (it is part of the image, here and on the GPU box) and replays the committed
tests/golden/preprocess_lanczos4.npz made by oracle/make_golden_preprocess.py.
"""
from __future__ import annotations

import math

import numpy as np

_S45 = 0.70710678118654752440084436210485
_CS = [(1, 0), (-_S45, -_S45), (0, 1), (_S45, -_S45), (-1, 0), (_S45, _S45), (0, -1), (-_S45, _S45)]
_PI = 3.1415926535897932384626433832795
COEF_BITS = 11  # INTER_RESIZE_COEF_BITS


def lanczos4_coeffs(x) -> np.ndarray:
    """cv::interpolateLanczos4 (float / double mix as in OpenCV)."""
    x = np.float32(x)
    co = np.zeros(8, dtype=np.float32)
    s = np.float32(0)
    y0 = -(float(x) + 3) * _PI * 0.25
    s0, c0 = math.sin(y0), math.cos(y0)
    for i in range(8):
        y0_ = np.float32(x + np.float32(3) - np.float32(i))
        if abs(y0_) >= np.float32(1e-6):
            y = -float(y0_) * _PI * 0.25
            co[i] = np.float32((_CS[i][0] * s0 + _CS[i][1] * c0) / (y * y))
        else:
            co[i] = np.float32(1e30)
        s = np.float32(s + co[i])
    s = np.float32(np.float32(1) / s)
    return (co * s).astype(np.float32)


def lanczos4_tables(src: int, dst: int):
    """Per destination index: first-tap source offset (sx, taps at sx - 3 .. sx + 4) and 8 int16 coefficients."""
    scale = 1.0 / (dst / src)
    ofs = np.zeros(dst, dtype=np.int32)
    coef = np.zeros((dst, 8), dtype=np.int16)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        ofs[d] = s
        c = lanczos4_coeffs(f)
        coef[d] = np.clip(np.rint(c * np.float32(1 << COEF_BITS)), -32768, 32767).astype(np.int16)
    return ofs, coef


def resize_lanczos4_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LANCZOS4) for uint8 HWC images, bit-exact."""
    H, W, _ = img.shape
    xo, xa = lanczos4_tables(W, dw)
    yo, ya = lanczos4_tables(H, dh)
    src = img.astype(np.int64)
    tmp = np.zeros((H, dw, img.shape[2]), dtype=np.int64)
    for k in range(8):
        tmp += src[:, np.clip(xo + k - 3, 0, W - 1), :] * xa[None, :, k, None].astype(np.int64)
    out = np.zeros((dh, dw, img.shape[2]), dtype=np.int64)
    for k in range(8):
        out += tmp[np.clip(yo + k - 3, 0, H - 1), :, :] * ya[:, k, None, None].astype(np.int64)
    out = (out + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS)
    return np.clip(out, 0, 255).astype(np.uint8)


def policy_image(img_u8_hwc: np.ndarray, size: int = 224):
    """simpler.py:47-65 -> (uint8 [size, size, 3], float32 [1, 3, size, size] in [-1, 1])."""
    small = resize_lanczos4_u8(img_u8_hwc, size, size)
    x = small.transpose(2, 0, 1)[None].astype(np.float32) * np.float32(1 / 255.0)   # pipeline.py:34-39
    x = (x - np.float32(0.5)) / np.float32(0.5)                                     # pipeline.py:42-55
    return small, x.astype(np.float32)


# ------------------------------------------------------------------------------------------------------------------
# Verifier image: open_clip's SigLIP transform as applied at bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:
# 249-254 (self.preprocess): PIL Image.resize((384, 384), BICUBIC) -> ToTensor -> Normalize(0.5, 0.5).  Pillow's 8-bit
# resampler (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
# Vertical_8bpc; Pillow unpinned in the reference, 12.2.0 here) restated and pinned bit for bit against PIL itself.
_PIL_PREC = 22


def _pil_bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_tables(in_size: int, out_size: int):
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ww = 0.0
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [0.0] * ksize
        for x in range(xmax):
            w = _pil_bicubic((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x] * (1 << _PIL_PREC)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resize_pil_bicubic_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """np.asarray(Image.fromarray(img).resize((dw, dh), Image.BICUBIC)) for uint8 RGB images, bit-exact."""
    H, W, ch = img.shape
    bx, kx = pil_tables(W, dw)
    by, ky = pil_tables(H, dh)
    src = img.astype(np.int64)
    half = 1 << (_PIL_PREC - 1)
    tmp = np.zeros((H, dw, ch), dtype=np.int64)
    for xx in range(dw):
        xmin, n = bx[xx]
        ss = np.full((H, ch), half, dtype=np.int64)
        for x in range(n):
            ss += src[:, xmin + x, :] * kx[xx, x]
        tmp[:, xx, :] = np.clip(ss >> _PIL_PREC, 0, 255)
    out = np.zeros((dh, dw, ch), dtype=np.int64)
    for yy in range(dh):
        ymin, n = by[yy]
        ss = np.full((dw, ch), half, dtype=np.int64)
        for y in range(n):
            ss += tmp[ymin + y] * ky[yy, y]
        out[yy] = np.clip(ss >> _PIL_PREC, 0, 255)
    return out.astype(np.uint8)


def verifier_image(img_u8_hwc: np.ndarray, size: int = 384):
    """-> (uint8 [size, size, 3], float32 [1, 3, size, size]): resize, ToTensor (x / 255), Normalize(0.5, 0.5)."""
    small = resize_pil_bicubic_u8(img_u8_hwc, size, size)
    x = small.transpose(2, 0, 1)[None].astype(np.float32) / np.float32(255)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return small, x.astype(np.float32)


# ------------------------------------------------------------------------------------------------------------------
# process_raw_image_to_jpg (CoVer_VLA/inference/experiments/robot/simpler/eval_utils.py:228-286):
#   tf.image.resize(image_u8, (256, 256), method=BILINEAR, preserve_aspect_ratio=False, antialias=True) -> tf.cast(uint8)
# Third-party arithmetic: TensorFlow ("tensorflow" is an unpinned dependency of CoVer_VLA/inference/pyproject.toml; NOT
# installed here and not vendored).  Restated from its published algorithm - tensorflow/python/ops/image_ops_impl.py
# (resize_images_v2 with antialias -> scale_and_translate, scale = float32(new) / float32(old), translation 0) and
# tensorflow/core/kernels/image/scale_and_translate_op.cc (ComputeSpansCore with the triangle kernel of radius 1,
# kernel_scale = max(1 / scale, 1); GatherRows then GatherColumns, float32, sequential multiply-adds).  PARITY UNPINNED
# against TensorFlow itself: there is no TensorFlow to run and the reference holds no fixture for this step; the CUDA
# kernel is bit-exact against THIS restatement (tests/test_preprocess.py), and the structural properties any correct
# triangle-filter resize has (constant images stay constant, weights sum to one, identity at equal sizes) are tested.
def tf_spans(in_size: int, out_size: int):
    f32 = np.float32
    scale = f32(out_size) / f32(in_size)
    inv_scale = f32(1.0) / scale
    kernel_scale = max(inv_scale, f32(1.0))
    radius = f32(1.0)
    span = min(2 * int(math.ceil(float(radius * kernel_scale))) + 1, in_size)
    starts = np.zeros(out_size, dtype=np.int64)
    weights = np.zeros((out_size, span), dtype=np.float32)
    for x in range(out_size):
        col_f = f32(x) + f32(0.5)
        sample_f = f32(col_f * inv_scale)
        if sample_f < 0 or sample_f > in_size:
            continue
        s0 = int(math.ceil(float(f32(f32(sample_f - f32(radius * kernel_scale)) - f32(0.5)))))
        s1 = int(math.floor(float(f32(f32(sample_f + f32(radius * kernel_scale)) - f32(0.5)))))
        s0 = min(max(s0, 0), in_size - 1)
        s1 = min(max(s1, 0), in_size - 1) + 1
        tmp = np.zeros(s1 - s0, dtype=np.float32)
        total = f32(0.0)
        for i in range(s1 - s0):
            kernel_pos = f32(f32(f32(s0 + i) + f32(0.5)) - sample_f)
            v = abs(f32(kernel_pos / kernel_scale))
            w = max(f32(0.0), f32(f32(1.0) - v))
            total = f32(total + w)
            tmp[i] = w
        starts[x] = s0
        if abs(total) >= 1000.0 * np.finfo(np.float32).tiny:
            one_over = f32(f32(1.0) / total)
            n = min(len(tmp), span)
            weights[x, :n] = (tmp[:n] * one_over).astype(np.float32)
    return starts, weights, span


def tf_resize_bilinear_antialias_u8(img: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [dh, dw, C]; float32 sequential accumulation in source order, truncating cast."""
    H, W, C = img.shape
    ys, wy, sy = tf_spans(H, dh)
    xs, wx, sx = tf_spans(W, dw)
    src = img.astype(np.float32)
    inter = np.zeros((dh, W, C), dtype=np.float32)
    for k in range(sy):  # rows first (GatherRows)
        idx = ys + k
        ok = idx < H
        rows = src[np.minimum(idx, H - 1)]  # [dh, W, C]
        term = (rows * wy[:, k][:, None, None]).astype(np.float32)
        inter = np.where(ok[:, None, None], (inter + term).astype(np.float32), inter)
    out = np.zeros((dh, dw, C), dtype=np.float32)
    for k in range(sx):  # then columns (GatherColumns)
        idx = xs + k
        ok = idx < W
        cols = inter[:, np.minimum(idx, W - 1)]  # [dh, dw, C]
        term = (cols * wx[:, k][None, :, None]).astype(np.float32)
        out = np.where(ok[None, :, None], (out + term).astype(np.float32), out)
    return np.clip(out, 0.0, 255.0).astype(np.int32).astype(np.uint8)  # tf.cast truncates toward zero
