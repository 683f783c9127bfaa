"""Episode-batched driver on the GPU: several environments sharing cvb_cover_step_batch launches
must produce, episode by episode, exactly what each environment produces when it is stepped alone through the
single-observation surfaces in the reference's order of operations (run_simpler_eval_with_openpi.py:231-441) - every score,
index, instruction, executed action and episode length, bit for bit."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import exec_action_oracle as X
from oracle import pi0_oracle as O
from oracle import verifier_oracle as V
from tests.helpers import build_full_engine

pytestmark = pytest.mark.gpu

FH, FW = 48, 64


class ToyEnv:
    """Deterministic simulator stand-in: frames and states depend on the seed and on every executed action."""

    def reset(self, task, seed):
        self.k, self.acc = 0, hashlib.sha256(str((task, seed)).encode()).digest()
        return {"acc": self.acc}

    def step(self, action):
        a = np.asarray(action, dtype=np.float64)
        assert a.shape == (7,) and a[6] in (-1.0, 1.0)
        self.acc = hashlib.sha256(self.acc + a.tobytes()).digest()
        self.k += 1
        return {"acc": self.acc}, bool(self.acc[0] < 20 and self.k > 4)

    def frame(self, obs):
        return np.random.default_rng(int.from_bytes(obs["acc"][:8], "little")).integers(0, 256, size=(FH, FW, 3), dtype=np.uint8)

    def state(self, obs):
        return np.random.default_rng(int.from_bytes(obs["acc"][8:16], "little")).normal(size=7)


def _noise(record, t, shape):
    g = torch.Generator().manual_seed(record.task * 100003 + record.trial * 1009 + t)
    return torch.randn(shape, generator=g)


def _tasks(d, v, P, n_tasks):
    from cover_vla_b200.episodes import TaskPrompts
    out = []
    for k in range(n_tasks):
        g = torch.Generator().manual_seed(500 + k)
        lens = torch.randint(4, min(12, d.max_lang_len) + 1, (P,), generator=g)
        tok = torch.randint(3, d.vocab - 1, (P, d.max_lang_len), generator=g)
        mask = torch.arange(d.max_lang_len)[None, :] < lens[:, None]
        tok = torch.where(mask, tok, torch.zeros_like(tok))
        vt = torch.randint(1, v.vocab - 1, (P, v.text_ctx), generator=g)
        out.append(TaskPrompts.build([f"task {k} / instruction {i}" for i in range(P)], lambda s, t=tok, m=mask: (t, m),
                                     lambda s, x=vt: x, "cuda"))
    return out


def _episode_alone(eng, tp, task, trial, seed, R, K, n, max_steps, gate):
    """One environment, the reference's loop, through the single-observation surfaces only."""
    from cover_vla_b200 import preprocess
    from cover_vla_b200.cover import (BRIDGE_ACTION_P01, BRIDGE_ACTION_P99, BatchedCoverStep, CoverInputs, execution_action)
    from cover_vla_b200.episodes import EpisodeRecord
    cfg = eng.cfg
    step = BatchedCoverStep(eng, K, n_future=n)
    env = ToyEnv()
    obs = env.reset(task, seed)
    rec = EpisodeRecord(task=task, trial=trial, seed=seed)
    t, history, cur, done = 0, [], 0, False
    while t < max_steps:
        if t % n == 0:
            rows = [cur] + list(range(1, R))                                              # :299-302
            frame = torch.from_numpy(env.frame(obs)).cuda()
            state = torch.zeros(cfg.max_state_dim)
            state[:7] = torch.from_numpy(env.state(obs).astype(np.float32))
            num_past = min(len(history), 6)                                               # :333
            past = torch.from_numpy(np.stack(history[-num_past:])).cuda() if num_past else None
            x = CoverInputs(image=preprocess.policy_image(frame, cfg.vis_image)[0], lang_tokens=tp.pi0_tokens[rows].contiguous(),
                            lang_len=tp.pi0_len[rows].contiguous(), state=state.cuda(),
                            noise=_noise(rec, t, (R * K, cfg.chunk_size, cfg.max_action_dim)).cuda(),
                            vf_image=preprocess.verifier_image_from_raw(frame, cfg.vf_image)[0], vf_tokens=tp.vf_tokens[cur],
                            past=past, lang_len_max=tp.lang_len_max)
            actions, traj, scores, gmean, bidx, bscore = step.sample_and_score(BatchedCoverStep.stack([x]))
            s0 = float(scores[0, 0])                                                      # :344-352 (the gate call's score)
            idx, score = (0, s0) if s0 >= gate else (int(bidx[0]), float(bscore[0]))      # :355-366
            cur = rows[idx // K]                                                          # :366, :409
            idx_dev = torch.tensor([idx], dtype=torch.int32, device="cuda")
            # :368-391 for step 0 (vote of the K-sample group), :411-417 for the queued steps (converted alone)
            execs = [execution_action(actions[0], idx_dev, K if i == 0 else 1, i)[0].cpu().numpy() for i in range(n)]
            hist = X.verifier_trajectories(actions[0].cpu().numpy(), None, cfg.vf_history, BRIDGE_ACTION_P01, BRIDGE_ACTION_P99,
                                           n_future=n)[idx, cfg.vf_history - n:]
            assert np.array_equal(hist, traj[0, idx, cfg.vf_history - n:].cpu().numpy())
            queue = list(zip(execs, hist))
            rec.verifier_scores.append(score)
            rec.selected_indices.append(idx)
        else:
            rec.verifier_scores.append(None)
            rec.selected_indices.append(None)
        ex, hi = queue.pop(0)
        rec.selected_instructions.append(tp.instructions[cur])
        rec.execute_actions.append(ex)
        rec.step_timestamps.append(t)
        history.append(hi)
        obs, done = env.step(ex)
        if done:
            break
        t += 1
    rec.success, rec.episode_length = bool(done), t
    return rec


@pytest.mark.parametrize("gate", [0.1, -1e9, 1e9])
def test_batched_episodes_equal_episodes_stepped_alone(gate):
    from cover_vla_b200.episodes import EpisodeBatchDriver
    d, v = O.TINY, V.VTINY
    R, K, n, B, max_steps = 3, 2, d.chunk_size, 3, 11
    w, vw = O.make_pi0_weights(d, 0), V.make_verifier_weights(v, 0)
    eng = build_full_engine(d, w, v, vw, R, K, max_observations=B)
    tasks = _tasks(d, v, R + 1, 2)
    work = [(task, trial, 1000 + 10 * task + trial) for task in range(2) for trial in range(3)]
    drv = EpisodeBatchDriver(eng, [ToyEnv() for _ in range(B)], tasks, work, R, K, n_action_steps=n, max_steps=max_steps,
                             gate_threshold=gate, noise_fn=_noise)
    recs = drv.run()
    torch.cuda.synchronize()
    assert len(recs) == len(work) and drv.batched_calls < drv.decisions
    swaps = 0
    for r in recs:
        ref = _episode_alone(eng, tasks[r.task], r.task, r.trial, r.seed, R, K, n, max_steps, gate)
        assert (r.success, r.episode_length) == (ref.success, ref.episode_length)
        assert r.verifier_scores == ref.verifier_scores and r.selected_indices == ref.selected_indices
        assert r.selected_instructions == ref.selected_instructions and r.step_timestamps == ref.step_timestamps
        assert len(r.execute_actions) == len(ref.execute_actions)
        assert all(np.array_equal(a, b) for a, b in zip(r.execute_actions, ref.execute_actions))
        swaps += r.selected_instructions[0] != tasks[r.task].instructions[0] or len(set(r.selected_instructions)) > 1
    if gate == -1e9:   # always confident: candidate 0, the task description never changes
        assert swaps == 0 and all(i in (0, None) for r in recs for i in r.selected_indices)
    if gate == 1e9:    # never confident: the N-candidate selection decides, and some episode switches instruction
        picked = [i for r in recs for i in r.selected_indices if i is not None]
        assert swaps > 0 and any(i >= K for i in picked), picked
    eng.close()
