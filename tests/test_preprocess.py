"""Policy-side observation pre-processing: cv2.resize INTER_LANCZOS4 + process_images.

Reference: BridgeSimplerAdapter.preprocess INT-ACT/src/experiments/env_adapters/simpler.py:43-65, src/utils/pipeline.py:34-69."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as P

GOLD = Path(__file__).resolve().parent / "golden" / "preprocess_lanczos4.npz"


def _full_frame():
    return np.random.default_rng(1234).integers(0, 256, size=(480, 640, 3), dtype=np.uint8)


def test_oracle_matches_cv2_golden_bit_exact():
    z = np.load(GOLD)
    for i in range(int(z["n"])):
        u8, f32 = P.policy_image(z[f"img{i}"], int(z[f"size{i}"]))
        assert np.array_equal(u8, z[f"u8_{i}"]) and np.array_equal(f32, z[f"f32_{i}"])
    u8, f32 = P.policy_image(_full_frame(), 224)  # the simulator's 480 x 640 frame -> 224 x 224: digests of cv2's output
    assert hashlib.sha256(u8.tobytes()).hexdigest() == str(z["full_u8_sha256"])
    assert hashlib.sha256(f32.tobytes()).hexdigest() == str(z["full_f32_sha256"])
    assert int(u8.astype(np.int64).sum()) == int(z["full_u8_sum"])


def test_oracle_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for H, W, S in [(480, 640, 224), (224, 224, 224), (97, 301, 224), (512, 512, 64)]:
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        ref = cv2.resize(img, (S, S), interpolation=cv2.INTER_LANCZOS4)
        assert np.array_equal(P.resize_lanczos4_u8(img, S, S), ref), (H, W, S)


@pytest.mark.gpu
def test_cuda_policy_image_is_bit_exact():
    from cover_vla_b200 import preprocess
    z = np.load(GOLD)
    for i in range(int(z["n"])):
        img = torch.from_numpy(z[f"img{i}"]).cuda()
        f32, u8 = preprocess.policy_image(img, int(z[f"size{i}"]), return_u8=True)
        torch.cuda.synchronize()
        assert np.array_equal(u8.cpu().numpy(), z[f"u8_{i}"])
        assert np.array_equal(f32.cpu().numpy(), z[f"f32_{i}"])
    f32, u8 = preprocess.policy_image(torch.from_numpy(_full_frame()).cuda(), 224, return_u8=True)
    assert hashlib.sha256(u8.cpu().numpy().tobytes()).hexdigest() == str(z["full_u8_sha256"])
    assert hashlib.sha256(f32.cpu().numpy().tobytes()).hexdigest() == str(z["full_f32_sha256"])


@pytest.mark.gpu
def test_cuda_policy_image_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    from cover_vla_b200 import preprocess
    rng = np.random.default_rng(9)
    for H, W, S in [(480, 640, 224), (100, 130, 224), (224, 224, 224), (1080, 1920, 224)]:
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        ref = cv2.resize(img, (S, S), interpolation=cv2.INTER_LANCZOS4)
        _, u8 = preprocess.policy_image(torch.from_numpy(img).cuda(), S, return_u8=True)
        assert np.array_equal(u8.cpu().numpy(), ref), (H, W, S)


# ---------------------------------------------------------------------------------------------------------------
# verifier image: open_clip's SigLIP transform (PIL bicubic 384, ToTensor, Normalize) as applied at
# bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:249-254
# ---------------------------------------------------------------------------------------------------------------
def _vfull_frame():
    return np.random.default_rng(4321).integers(0, 256, size=(256, 256, 3), dtype=np.uint8)


def test_verifier_image_oracle_matches_pil_golden_bit_exact():
    z = np.load(GOLD)
    for i in range(int(z["vn"])):
        u8, f32 = P.verifier_image(z[f"vimg{i}"], int(z[f"vsize{i}"]))
        assert np.array_equal(u8, z[f"vu8_{i}"]) and np.array_equal(f32, z[f"vf32_{i}"])
    u8, f32 = P.verifier_image(_vfull_frame(), 384)
    assert hashlib.sha256(u8.tobytes()).hexdigest() == str(z["vfull_u8_sha256"])
    assert hashlib.sha256(f32.tobytes()).hexdigest() == str(z["vfull_f32_sha256"])


def test_verifier_image_oracle_matches_live_pil_and_the_host_mirror():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(6)
    for H, W, S in [(256, 256, 384), (480, 640, 384), (100, 77, 64), (384, 384, 384)]:
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).convert("RGB").resize((S, S), Image.BICUBIC))
        assert np.array_equal(P.resize_pil_bicubic_u8(img, S, S), ref), (H, W, S)
    # the host-side mirror of the reference transform (cover_vla_b200.verifier) produces the same float32 tensor
    from cover_vla_b200.verifier import efficient_ensemble_merged as E
    img = _vfull_frame()
    host = E.default_preprocess(384)(Image.fromarray(img))
    assert np.array_equal(host.numpy()[None], P.verifier_image(img, 384)[1])


@pytest.mark.gpu
def test_cuda_verifier_image_is_bit_exact():
    from cover_vla_b200 import preprocess
    z = np.load(GOLD)
    for i in range(int(z["vn"])):
        f32, u8 = preprocess.verifier_image(torch.from_numpy(z[f"vimg{i}"]).cuda(), int(z[f"vsize{i}"]), return_u8=True)
        torch.cuda.synchronize()
        assert np.array_equal(u8.cpu().numpy(), z[f"vu8_{i}"]) and np.array_equal(f32.cpu().numpy(), z[f"vf32_{i}"])
    f32, u8 = preprocess.verifier_image(torch.from_numpy(_vfull_frame()).cuda(), 384, return_u8=True)
    assert hashlib.sha256(u8.cpu().numpy().tobytes()).hexdigest() == str(z["vfull_u8_sha256"])
    assert hashlib.sha256(f32.cpu().numpy().tobytes()).hexdigest() == str(z["vfull_f32_sha256"])


@pytest.mark.gpu
def test_cuda_verifier_image_matches_live_pil():
    Image = pytest.importorskip("PIL.Image")
    from cover_vla_b200 import preprocess
    rng = np.random.default_rng(10)
    for H, W, S in [(256, 256, 384), (480, 640, 384), (1080, 1920, 384), (64, 48, 384)]:
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).convert("RGB").resize((S, S), Image.BICUBIC))
        _, u8 = preprocess.verifier_image(torch.from_numpy(img).cuda(), S, return_u8=True)
        assert np.array_equal(u8.cpu().numpy(), ref), (H, W, S)
