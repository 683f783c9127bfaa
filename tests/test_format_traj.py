"""Verifier-format trajectories: oracle vs the reference (golden + live), CUDA vs both.

Reference: process_inputs(verifier_action=True) eval_utils.py:172-221, postprocess_verifier simpler.py:96-121, 222-226,
denormalize_bound base.py:20-31, the -5 left padding efficient_ensemble_merged.py:378-390."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import exec_action_oracle as X

GOLD = Path(__file__).resolve().parent / "golden" / "format_traj.npz"
REF = Path("/root/reference")
HISTORY = 10


def _cases():
    z = np.load(GOLD)
    for i in range(int(z["n"])):
        yield z[f"a{i}"], z[f"past{i}"], z[f"traj{i}"], z["p01"], z["p99"]


def _caller_past(past):
    """run_simpler_eval_with_openpi.py:333 / eval_utils.py:209: the last min(len, 6) executed actions."""
    return past[-min(len(past), 6):] if len(past) else None


def test_oracle_matches_reference_golden_bit_exact():
    n = 0
    for a, past, traj, p01, p99 in _cases():
        got = X.verifier_trajectories(a, _caller_past(past), HISTORY, p01, p99, n_future=4)
        assert got.dtype == np.float32 and got.shape == traj.shape == (a.shape[0], HISTORY, 7)
        assert np.array_equal(got, traj)
        n += 1
    assert n >= 6


def test_golden_covers_the_edges():
    """the cases hold: no history, a full 6-step history (no padding left), > 6 executed actions (only the last 6 are
    used), the gripper threshold itself (0.5 -> 1, just below -> 0) and the normalisation bounds (-1 -> p01, 1 -> p99)."""
    cs = list(_cases())
    assert {len(p) for _, p, _, _, _ in cs} >= {0, 6, 9}
    a, past, traj, p01, p99 = cs[0]
    assert (traj[:, :6] == -5).all() and traj[0, 6, 6] == 1.0 and traj[0, 7, 6] == 0.0
    mid = 0.5 * (p99 - p01) + p01  # action 0 -> the centre of the range
    assert np.array_equal(traj[0, 8, :6], np.array([p01[0], p99[1], mid[2], p01[3], p99[4], mid[5]]).astype(np.float32))
    a, past, traj, _, _ = [c for c in cs if len(c[1]) == 9][0]
    assert np.array_equal(traj[:, :6], np.broadcast_to(past[-6:].astype(np.float32), (a.shape[0], 6, 7)))


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the authoring container")
def test_oracle_matches_live_reference():
    from oracle import make_golden_format as G
    process_inputs, stats = G.reference_process_inputs()
    rng = np.random.default_rng(11)
    for N, n_past in [(7, 0), (7, 3), (3, 8)]:
        a = rng.uniform(-2, 2, size=(N, 4, 32)).astype(np.float32)
        a[:, :, 6] = rng.uniform(0, 1, size=(N, 4)).astype(np.float32)
        past = [rng.uniform(-0.1, 0.1, size=7) for _ in range(n_past)]
        ref = G.reference_trajectories(process_inputs, a, past, 4)
        got = X.verifier_trajectories(a, _caller_past(np.array(past).reshape(n_past, 7)), HISTORY,
                                      stats["action"]["p01"], stats["action"]["p99"], n_future=4)
        assert np.array_equal(ref, got)


@pytest.mark.gpu
def test_cuda_format_trajectories_matches_reference_golden_bit_exact():
    from cover_vla_b200 import cover
    for a, past, traj, p01, p99 in _cases():
        act = torch.from_numpy(a).cuda()
        cp = _caller_past(past)
        pd = None if cp is None else torch.from_numpy(cp.astype(np.float32)).cuda()
        out = cover.format_trajectories(act, pd, HISTORY, 4, p01=tuple(p01), p99=tuple(p99))
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), torch.from_numpy(traj))
