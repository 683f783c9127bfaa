"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the execution-format action of the selected candidate and
the gripper vote of its K-sample group.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Follows (reference, /root/reference):
  * process_inputs(verifier_action=False)                 CoVer_VLA/inference/experiments/robot/simpler/eval_utils.py:172-221
  * BridgeSimplerAdapter.postprocess                      INT-ACT/src/experiments/env_adapters/simpler.py:123-166
  * BaseEnvAdapter.denormalize_bound                      INT-ACT/src/experiments/env_adapters/base.py:20-31
  * euler2axangle = quat2axangle(euler2quat(., 'sxyz'))   INT-ACT/src/utils/geometry.py:261-291, 294-362, 365-436
  * BridgeSimplerAdapter.postprocess_gripper              INT-ACT/src/experiments/env_adapters/simpler.py:211-220
  * the gripper vote                                      CoVer_VLA/.../run_simpler_eval_with_openpi.py:368-391

Pinned: tests/test_exec_action.py replays tests/golden/exec_action.npz, which oracle/make_golden_exec.py produced by
executing the reference's own source lines (AST-extracted, unmodified) in the authoring container; where /root/reference
is present the same test also runs them live.
"""
from __future__ import annotations

import math

import numpy as np

_FLOAT_EPS = np.finfo(np.float64).eps


def euler2quat_sxyz(ai, aj, ak):
    """geometry.py:294-362 for axes='sxyz' (firstaxis 0, parity 0, repetition 0, frame 0 -> i, j, k = 1, 2, 3)."""
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si = math.cos(ai), math.sin(ai)
    cj, sj = math.cos(aj), math.sin(aj)
    ck, sk = math.cos(ak), math.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    q = np.empty((4,))
    q[0] = cj * cc + sj * ss
    q[1] = cj * sc - sj * cs
    q[2] = cj * ss + sj * cc
    q[3] = cj * cs - sj * sc
    return q


def quat2axangle(quat):
    """geometry.py:365-436 (identity_thresh=None -> 3 eps)."""
    quat = np.asarray(quat)
    Nq = np.sum(quat ** 2)
    if not np.isfinite(Nq):
        return np.array([1.0, 0, 0]), float("nan")
    identity_thresh = _FLOAT_EPS * 3
    if Nq < _FLOAT_EPS ** 2:
        return np.array([1.0, 0, 0]), 0.0
    if Nq != 1:
        quat = quat / math.sqrt(Nq)
    xyz = quat[1:]
    len2 = np.sum(xyz ** 2)
    if len2 < identity_thresh ** 2:
        return np.array([1.0, 0, 0]), 0.0
    theta = 2 * math.acos(max(min(quat[0], 1), -1))
    return xyz / math.sqrt(len2), theta


def postprocess_execution(action_1x7: np.ndarray, p01, p99) -> np.ndarray:
    """simpler.py:123-166 on one [1, 7] float32 action (normalisation type "bound")."""
    actions = np.asarray(action_1x7)
    p01 = np.asarray(p01, dtype=np.float64)[:6]
    p99 = np.asarray(p99, dtype=np.float64)[:6]
    raw6 = (actions[:, :-1] - (-1)) / (1 - (-1)) * (p99 - p01) + p01          # base.py:29-31
    raw = np.concatenate([raw6, actions[:, -1:]], axis=1)
    out = np.zeros((len(raw), 7))
    for idx, ra in enumerate(raw):
        roll, pitch, yaw = ra[3:6]
        ax, angle = quat2axangle(euler2quat_sxyz(roll, pitch, yaw))
        grip = 2.0 * (ra[-1] > 0.5) - 1.0                                      # simpler.py:215
        out[idx] = np.concatenate([ra[:3], ax * angle, [grip]])
    return out


def execution_action(actions: np.ndarray, best_idx: int, K: int, p01, p99, step: int = 0):
    """actions f32 [N, chunk, >=7].  Returns (execute_action f64 [7], (close_votes, open_votes)),
    run_simpler_eval_with_openpi.py:368-391 with num_past-th = first future step."""
    a = np.asarray(actions, dtype=np.float32)[:, :, :7]
    g0 = (best_idx // K) * K
    execs = [postprocess_execution(a[m, step][None, :], p01, p99)[0] for m in range(g0, g0 + K)]
    execute_action = execs[best_idx - g0].copy()
    grippers = np.stack(execs)[:, -1]
    close_votes = int((grippers >= 0).sum())
    open_votes = int((grippers < 0).sum())
    if close_votes > open_votes:
        execute_action[-1] = 1.0
    elif open_votes > close_votes:
        execute_action[-1] = -1.0
    else:
        execute_action[-1] = 1.0 if execute_action[-1] >= 0 else -1.0
    execute_action[-1] = float(np.sign(execute_action[-1]))
    return execute_action, (close_votes, open_votes)


def verifier_trajectories(actions: np.ndarray, past, history: int, p01, p99, n_future: int | None = None) -> np.ndarray:
    """Verifier-format trajectories of all N candidates: f32 [N, history, 7].

    Follows process_inputs(verifier_action=True) (eval_utils.py:172-221): every future step of every candidate goes through
    BridgeSimplerAdapter.postprocess_verifier (simpler.py:96-121: denormalize_bound on the 6 pose dims in float64 - no
    clipping, base.py:20-31 - and the gripper binarised to 0 / 1 at 0.5, simpler.py:222-226), the caller's last <= 6
    executed actions are put in front of each candidate's futures, and EfficientEnsembleMerged left-pads with -5 to
    `history` steps before the float32 cast (efficient_ensemble_merged.py:378-390).  `past` = None or [num_past, 7].
    Pinned by tests/golden/format_traj.npz (oracle/make_golden_format.py ran the reference's own lines)."""
    a = np.asarray(actions, dtype=np.float32)[:, :, :7]
    if n_future is not None:
        a = a[:, :n_future]
    p01 = np.asarray(p01, dtype=np.float64)[:6]
    p99 = np.asarray(p99, dtype=np.float64)[:6]
    fut = np.zeros(a.shape, dtype=np.float64)
    fut[:, :, :6] = (a[:, :, :6] - (-1)) / (1 - (-1)) * (p99 - p01) + p01       # base.py:29-30
    fut[:, :, 6] = np.where(a[:, :, 6] < 0.5, 0, 1)                             # simpler.py:225
    out = []
    for n in range(a.shape[0]):
        rows = fut[n]
        if past is not None and len(past) > 0:
            rows = np.concatenate([np.asarray(past, dtype=np.float64).reshape(-1, 7), rows], axis=0)   # eval_utils.py:213-216
        if len(rows) < history:
            rows = np.vstack([np.ones((history - len(rows), 7)) * -5, rows])    # efficient_ensemble_merged.py:384-386
        out.append(rows)
    return np.array(out).astype(np.float32)
