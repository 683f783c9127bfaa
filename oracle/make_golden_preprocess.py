"""Generates tests/golden/preprocess_lanczos4.npz with cv2.resize itself (INTER_LANCZOS4, the call of
INT-ACT/src/experiments/env_adapters/simpler.py:47-51) and torch's float32 arithmetic of pipeline.py:34-69.

    python -m oracle.make_golden_preprocess
"""
import hashlib
from pathlib import Path

import cv2
import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent


def reference(img, size):
    small = cv2.resize(img, (size, size), interpolation=cv2.INTER_LANCZOS4)
    images = torch.as_tensor(small, dtype=torch.uint8).permute(2, 0, 1)[None]
    images = images * (1 / 255.0)
    images = (images - torch.tensor([0.5, 0.5, 0.5])[None, :, None, None]) / torch.tensor([0.5, 0.5, 0.5])[None, :, None, None]
    return small, images.numpy()


def main():
    rng = np.random.default_rng(0)
    out = {"cv2_version": np.array(cv2.__version__)}
    small_cases = [(120, 160, 64), (64, 64, 64), (50, 37, 96), (33, 200, 56)]
    for i, (H, W, S) in enumerate(small_cases):
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        if i == 1:
            img[:] = np.where((np.add.outer(np.arange(H), np.arange(W)) % 2 == 0)[..., None], 255, 0)  # checkerboard extremes
        u8, f32 = reference(img, S)
        out[f"img{i}"], out[f"size{i}"], out[f"u8_{i}"], out[f"f32_{i}"] = img, S, u8, f32
    out["n"] = len(small_cases)
    # the real shape (480 x 640 simulator frame -> 224 x 224): the input is regenerated from the seed, only digests are stored
    img = np.random.default_rng(1234).integers(0, 256, size=(480, 640, 3), dtype=np.uint8)
    u8, f32 = reference(img, 224)
    out["full_u8_sha256"] = np.array(hashlib.sha256(u8.tobytes()).hexdigest())
    out["full_f32_sha256"] = np.array(hashlib.sha256(f32.tobytes()).hexdigest())
    out["full_u8_sum"] = np.array(int(u8.astype(np.int64).sum()))
    # verifier side: PIL bicubic + torchvision-style ToTensor / Normalize (open_clip's SigLIP transform)
    from PIL import Image
    import PIL
    out["pil_version"] = np.array(PIL.__version__)

    def reference_v(img, size):
        u8 = np.asarray(Image.fromarray(img).convert("RGB").resize((size, size), Image.BICUBIC))
        t = torch.from_numpy(u8.copy()).permute(2, 0, 1)[None].to(torch.float32).div(255)
        t = t.sub(0.5).div(0.5)
        return u8, t.numpy()

    vcases = [(64, 64, 96), (120, 160, 64), (50, 37, 96)]
    for i, (H, W, S) in enumerate(vcases):
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        if i == 0:
            img[:] = np.where((np.add.outer(np.arange(H), np.arange(W)) % 2 == 0)[..., None], 255, 0)
        u8, f32 = reference_v(img, S)
        out[f"vimg{i}"], out[f"vsize{i}"], out[f"vu8_{i}"], out[f"vf32_{i}"] = img, S, u8, f32
    out["vn"] = len(vcases)
    img = np.random.default_rng(4321).integers(0, 256, size=(256, 256, 3), dtype=np.uint8)  # the 256 x 256 frame of eval_utils.py:273-283
    u8, f32 = reference_v(img, 384)
    out["vfull_u8_sha256"] = np.array(hashlib.sha256(u8.tobytes()).hexdigest())
    out["vfull_f32_sha256"] = np.array(hashlib.sha256(f32.tobytes()).hexdigest())
    np.savez_compressed(ROOT / "tests/golden/preprocess_lanczos4.npz", **out)
    print("wrote golden, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
