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


# ---------------------------------------------------------------------------------------------------------------
# verifier frame: process_raw_image_to_jpg (eval_utils.py:228-286) = tf.image.resize(BILINEAR, antialias=True) -> uint8.
# TensorFlow is absent (parity unpinned upstream): the restatement's structural properties, then CUDA == restatement.
# ---------------------------------------------------------------------------------------------------------------
def test_tf_antialias_resize_restatement_properties():
    rng = np.random.default_rng(8)
    for n_in, n_out in [(480, 256), (640, 256), (256, 256), (100, 256), (1080, 256)]:
        starts, w, span = P.tf_spans(n_in, n_out)
        assert w.dtype == np.float32 and (w >= 0).all()
        assert np.abs(w.sum(1) - 1).max() < 1e-6                      # normalised triangle weights
        assert (starts >= 0).all() and (starts + (w > 0).sum(1) <= n_in).all()
        centre = (np.arange(n_out) + 0.5) * n_in / n_out               # the filter is centred on the sample position
        mean_src = (w * (starts[:, None] + np.arange(span)[None, :] + 0.5)).sum(1)
        inside = (centre > span) & (centre < n_in - span)
        assert np.abs(mean_src - centre)[inside].max() < 0.1               # (a sampled triangle is only nearly symmetric)
    img = rng.integers(0, 256, size=(64, 48, 3), dtype=np.uint8)
    assert np.array_equal(P.tf_resize_bilinear_antialias_u8(img, 64, 48), img)   # identity at equal sizes
    flat = np.full((480, 640, 3), 200, np.uint8)
    out = P.tf_resize_bilinear_antialias_u8(flat, 256, 256)
    assert set(np.unique(out)) <= {199, 200}                            # float32 weight sums of 1 - 1 ulp truncate down
    # down-scaling averages: a 2x2 checkerboard of 0 / 255 at 2:1 lands mid-grey, not on an aliased extreme
    cb = (np.indices((512, 512)).sum(0) % 2 * 255).astype(np.uint8)[:, :, None].repeat(3, 2)
    out = P.tf_resize_bilinear_antialias_u8(cb, 256, 256)
    assert 120 <= out.min() and out.max() <= 135


@pytest.mark.gpu
def test_cuda_tf_antialias_resize_is_bit_exact_vs_restatement():
    from cover_vla_b200 import preprocess
    rng = np.random.default_rng(12)
    for H, W, S in [(480, 640, 256), (512, 640, 256), (256, 256, 256), (100, 130, 256), (1080, 1920, 256), (37, 53, 64)]:
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        ref = P.tf_resize_bilinear_antialias_u8(img, S, S)
        out = preprocess.verifier_frame(torch.from_numpy(img).cuda(), S)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref), (H, W, S)
    # the whole verifier-side image path from one uint8 frame: 256^2 antialias -> PIL-bicubic 384^2 -> normalise
    img = rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8)
    f32 = preprocess.verifier_image_from_raw(torch.from_numpy(img).cuda(), 384)
    ref = P.verifier_image(P.tf_resize_bilinear_antialias_u8(img, 256, 256), 384)[1]
    assert np.array_equal(f32.cpu().numpy(), ref)


@pytest.mark.parametrize("H,W", [(480, 640), (512, 512), (300, 400), (256, 256), (720, 1280)])
def test_tf_resize_restatement_agrees_with_an_independent_implementation(H, W):
    """tf.image.resize(..., BILINEAR, antialias=True) is restated from TensorFlow's scale_and_translate (TensorFlow is not
    installed: parity unpinned against TF itself).  torch's F.interpolate(mode="bilinear", antialias=True) is an independent
    implementation of the same triangle-filter resize with half-pixel centres: after the reference's truncating uint8 cast
    the two may differ by one count where fp32 summation order moves a value across an integer, never by more."""
    import torch.nn.functional as F
    from oracle import preprocess_oracle as P
    rng = np.random.default_rng(H + W)
    yy, xx = np.mgrid[0:H, 0:W]
    smooth = np.stack([(np.sin(xx / 37.0) + np.cos(yy / 23.0)) * 60 + 128, xx * 255.0 / W, yy * 255.0 / H], -1)
    for img in (rng.integers(0, 256, (H, W, 3), dtype=np.uint8), smooth.clip(0, 255).astype(np.uint8)):
        ours = P.tf_resize_bilinear_antialias_u8(img, 256, 256)
        t = torch.from_numpy(img).permute(2, 0, 1)[None].float()
        ref = F.interpolate(t, size=(256, 256), mode="bilinear", antialias=True, align_corners=False)
        ref_u8 = ref[0].permute(1, 2, 0).numpy().astype(np.uint8)  # tf.cast(float -> uint8) truncates
        diff = np.abs(ours.astype(np.int32) - ref_u8.astype(np.int32))
        assert diff.max() <= 1 and (diff == 0).mean() > 0.97, (diff.max(), (diff == 0).mean())


@pytest.mark.skipif(not __import__("pathlib").Path("/root/reference/INT-ACT/src/utils/pipeline.py").exists(),
                    reason="/root/reference is not present")
def test_policy_image_oracle_matches_the_reference_preprocess_executed_unmodified():
    """The reference's own BridgeSimplerAdapter.preprocess (INT-ACT/src/experiments/env_adapters/simpler.py:43-93) with the real
    rescale / normalize / process_images of INT-ACT/src/utils/pipeline.py:34-69 and normalize_bound (base.py:8-18), AST-extracted
    and executed as they are (the modules' import chains need absent packages): its 'observation.images.top' tensor equals the
    oracle's policy_image bit for bit - cv2's Lanczos4 AND torch's float32 rescale / normalise arithmetic."""
    import ast
    import json
    import types
    cv2 = pytest.importorskip("cv2")
    from oracle.make_golden_exec import REF, _method_source
    src = (REF / "INT-ACT/src/utils/pipeline.py").read_text()
    ns = {"torch": torch, "np": np, "cv2": cv2,
          "IMAGENET_STANDARD_MEAN": torch.tensor([0.5, 0.5, 0.5]), "IMAGENET_STANDARD_STD": torch.tensor([0.5, 0.5, 0.5])}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("rescale", "normalize", "process_images"):
            exec(ast.get_source_segment(src, node), ns)
    exec(_method_source(REF / "INT-ACT/src/experiments/env_adapters/simpler.py", "SimplerAdapter", "preprocess"), ns)
    exec(_method_source(REF / "INT-ACT/src/experiments/env_adapters/base.py", "BaseEnvAdapter", "normalize_bound"), ns)
    stats = json.loads((REF / "INT-ACT/config/dataset/bridge_statistics.json").read_text())
    me = types.SimpleNamespace(image_size=(224, 224), state_normalization_type="bound", dataset_statistics=stats,
                               dtype=torch.float32, preprocess_proprio=lambda s: np.asarray(s, dtype=np.float64))
    me.normalize_bound = types.MethodType(ns["normalize_bound"], me)
    rng = np.random.default_rng(12)
    for H, W in [(480, 640), (224, 224), (97, 301)]:
        frame = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        out = ns["preprocess"](me, {"observation.images.top": frame, "observation.state": rng.normal(size=7), "task": "x"})
        img = out["observation.images.top"]
        assert img.dtype == torch.float32 and tuple(img.shape) == (1, 3, 224, 224)
        u8, f32 = P.policy_image(frame, 224)
        assert np.array_equal(img.numpy(), f32), (H, W)
        assert out["task"] == ["x"] and tuple(out["observation.state"].shape) == (1, 7)
