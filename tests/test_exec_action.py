"""Execution-format action + gripper vote: oracle vs the reference (golden + live), CUDA vs oracle.

Reference: run_simpler_eval_with_openpi.py:368-391, BridgeSimplerAdapter.postprocess simpler.py:123-166,
euler2axangle src/utils/geometry.py:261-436."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import exec_action_oracle as X

GOLD = Path(__file__).resolve().parent / "golden" / "exec_action.npz"
REF = Path("/root/reference")


def _cases():
    z = np.load(GOLD)
    for i in range(int(z["n"])):
        yield z[f"a{i}"], int(z[f"idx{i}"]), int(z[f"K{i}"]), z[f"ex{i}"], tuple(int(v) for v in z[f"v{i}"]), z["p01"], z["p99"]


def test_oracle_matches_reference_golden_bit_exact():
    n = 0
    for a, idx, K, ex, votes, p01, p99 in _cases():
        got, v = X.execution_action(a, idx, K, p01, p99)
        assert got.dtype == np.float64 and np.array_equal(got, ex, equal_nan=True), (idx, K, got, ex)
        assert v == votes
        n += 1
    assert n >= 10


def test_statistics_constants_match_the_golden_file():
    from cover_vla_b200 import cover
    z = np.load(GOLD)
    assert np.array_equal(np.array(cover.BRIDGE_ACTION_P01), z["p01"])
    assert np.array_equal(np.array(cover.BRIDGE_ACTION_P99), z["p99"])


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the authoring container")
def test_oracle_matches_live_reference():
    from oracle import make_golden_exec as G
    adapter, stats = G.reference_adapter()
    rng = np.random.default_rng(7)
    for R, K in [(4, 3), (2, 6)]:
        a = rng.uniform(-2, 2, size=(R * K, 4, 32)).astype(np.float32)
        a[:, :, 6] = rng.uniform(0, 1, size=(R * K, 4)).astype(np.float32)
        for idx in range(R * K):
            ref, rv = G.reference_execution_action(adapter, a, idx, K)
            got, v = X.execution_action(a, idx, K, stats["action"]["p01"], stats["action"]["p99"])
            assert np.array_equal(ref, got) and tuple(rv) == v


@pytest.mark.gpu
def test_cuda_execution_action_matches_oracle():
    from cover_vla_b200 import cover
    for a, idx, K, ex, votes, p01, p99 in _cases():
        act = torch.from_numpy(a).cuda()
        out, v = cover.execution_action(act, torch.tensor([idx], dtype=torch.int32, device="cuda"), K,
                                        p01=tuple(p01), p99=tuple(p99))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert np.array_equal(got[:3], ex[:3])                    # float64 de-normalisation: bit-exact
        assert np.allclose(got[3:6], ex[3:6], rtol=0, atol=1e-12)  # device sin / cos / acos differ from libm by ulps
        assert got[6] == ex[6] and tuple(v.cpu().tolist()) == votes


@pytest.mark.gpu
def test_cuda_execution_action_random_sweep():
    from cover_vla_b200 import cover
    rng = np.random.default_rng(3)
    R, K = 8, 5
    a = rng.uniform(-1.5, 1.5, size=(R * K, 4, 32)).astype(np.float32)
    a[:, :, 6] = rng.uniform(0, 1, size=(R * K, 4)).astype(np.float32)
    act = torch.from_numpy(a).cuda()
    for idx in range(0, R * K, 3):
        for step in (0, 2):
            out, v = cover.execution_action(act, torch.tensor([idx], dtype=torch.int32, device="cuda"), K, step=step)
            ref, rv = X.execution_action(a, idx, K, cover.BRIDGE_ACTION_P01, cover.BRIDGE_ACTION_P99, step=step)
            got = out.cpu().numpy()
            assert np.array_equal(got[:3], ref[:3]) and got[6] == ref[6] and tuple(v.cpu().tolist()) == rv
            assert np.allclose(got[3:6], ref[3:6], rtol=0, atol=1e-12)
