"""Episode-batched driver: the per-episode state machine of
run_simpler_eval_with_openpi.py:231-441 kept for several environments at once.

CPU part: the bookkeeping (decision ticks, action queue, verifier-format history tail, instruction swap, termination,
environment reuse, grouping by history length) against a one-environment restatement of the reference loop, with the
device decision replaced by a deterministic function of EVERYTHING the driver hands it (a wrong history tail, prompt set,
frame or state changes the outcome).  The real decision path is covered by tests/test_episodes_gpu.py."""
import hashlib
import types

import numpy as np
import pytest
import torch

from cover_vla_b200.episodes import MAX_PAST, EpisodeBatchDriver, TaskPrompts

H, W = 12, 16


class ToyEnv:
    """Deterministic stand-in for the simulator: what it shows depends on the seed and on every action it was given."""

    def reset(self, task, seed):
        self.seed, self.k, self.acc = seed, 0, hashlib.sha256(str((task, seed)).encode()).digest()
        return self._obs()

    def _obs(self):
        return {"k": self.k, "acc": self.acc}

    def step(self, action):
        a = np.asarray(action, dtype=np.float64)
        assert a.shape == (7,)
        self.acc = hashlib.sha256(self.acc + a.tobytes()).digest()
        self.k += 1
        done = self.acc[0] < 6 and self.k > 5   # ~2 % per step: some episodes end early, others run to max_steps
        return self._obs(), done

    def frame(self, obs):
        rng = np.random.default_rng(int.from_bytes(obs["acc"][:8], "little"))
        return rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)

    def state(self, obs):
        rng = np.random.default_rng(int.from_bytes(obs["acc"][8:16], "little"))
        return rng.normal(size=7)


def fake_decision(frame, state, past, rows, task, R, K, n):
    """Deterministic stand-in for cvb_cover_step_batch + post-processing: a function of all decision inputs."""
    h = hashlib.sha256()
    for part in (frame.tobytes(), np.asarray(state, dtype=np.float32).tobytes(),
                 b"" if past is None else np.asarray(past, dtype=np.float32).tobytes(), str((rows, task)).encode()):
        h.update(part)
    rng = np.random.default_rng(int.from_bytes(h.digest()[:8], "little"))
    s0 = rng.uniform(-0.3, 0.3)
    idx = 0 if s0 >= 0.1 else int(rng.integers(0, R * K))
    return {"idx": idx, "score": float(s0 if idx == 0 else rng.uniform(-0.3, 0.3)),
            "exec": rng.normal(size=(n, 7)), "hist": rng.normal(size=(n, 7)).astype(np.float32)}


class FakeDeviceDriver(EpisodeBatchDriver):
    groups = None

    def _decide(self, group):
        num_past = min(len(group[0].history), MAX_PAST)
        assert all(min(len(s.history), MAX_PAST) == num_past for s in group)   # one history length per batched call
        assert len(group) <= self.engine.cfg.max_observations
        self.groups.append(len(group))
        out = []
        for s in group:
            hi = self._host_inputs(s, num_past)
            rows = self.tasks[s.task].prompt_rows(s.current, self.R)
            out.append(fake_decision(hi["frame"], hi["state"][:7], hi["past"], rows, s.task, self.R, self.K, self.n_action_steps))
        self.decisions += len(group)
        self.batched_calls += 1
        return out


def reference_episode(task, trial, seed, prompts, R, K, n, max_steps, wait):
    """run_simpler_eval_with_openpi.py:231-441 for ONE environment, in the reference's order of operations."""
    env = ToyEnv()
    obs = env.reset(task, seed)                                              # :194, :231
    t, action_history, task_description = 0, [], 0                           # :234-236, :220-223 (instruction ids)
    rec = types.SimpleNamespace(scores=[], instr=[], acts=[], ts=[], idx=[])
    done = False
    while t < max_steps + wait:                                              # :259
        if t < wait:                                                         # :261-265
            obs, done = env.step(np.array([0, 0, 0, 0, 0, 0, -1], dtype=np.float64))
            t += 1
            continue
        if t % n == 0:                                                       # :322, :329
            unique_prompts = [task_description] + list(range(1, R))          # :299-302
            num_past = min(len(action_history), 6)                           # :333
            past = np.stack(action_history[-num_past:]) if num_past else None
            r = fake_decision(env.frame(obs), env.state(obs), past, unique_prompts, task, R, K, n)
            max_instruction = unique_prompts[r["idx"] // K]                  # :366 (gate pass: idx 0 -> task_description)
            execute_action = r["exec"][0]                                    # :372-391
            queue = [(r["exec"][i], r["hist"][i]) for i in range(1, n)]      # :393-400
            rec.scores.append(r["score"]); rec.idx.append(r["idx"])
            rec.instr.append(prompts.instructions[max_instruction])
            task_description = max_instruction                               # :409
            history_row = r["hist"][0]                                       # :427
        else:
            execute_action, history_row = queue.pop(0)                       # :411-417, :429-432
            rec.scores.append(None); rec.idx.append(None)
            rec.instr.append(prompts.instructions[task_description])
        rec.acts.append(execute_action.copy()); rec.ts.append(t)
        action_history.append(history_row)                                   # :433
        obs, done = env.step(execute_action)                                 # :436
        if done:                                                             # :438-441
            break
        t += 1
    rec.success, rec.length = bool(done), t                                  # :455-456
    return rec


def _prompts(P):
    return TaskPrompts([f"instruction {i}" for i in range(P)], torch.zeros(P, 8, dtype=torch.int64),
                       torch.full((P,), 3, dtype=torch.int32), torch.zeros(P, 4, dtype=torch.int64), 3)


def _engine(max_obs):
    cfg = types.SimpleNamespace(chunk_size=4, max_observations=max_obs, max_rephrases=8, max_samples=5, max_state_dim=32,
                                vf_history=10, max_action_dim=32, vis_image=224, vf_image=384)
    return types.SimpleNamespace(cfg=cfg, device="cpu")


@pytest.mark.parametrize("n_envs,wait,max_steps", [(1, 0, 30), (3, 0, 41), (4, 4, 26), (8, 0, 150)])
def test_driver_matches_the_reference_loop(n_envs, wait, max_steps):
    R, K, n = 4, 3, 4
    tasks = [_prompts(R), _prompts(R + 2)]
    work = [(task, trial, 1000 + trial) for task in range(2) for trial in range(7)]
    FakeDeviceDriver.groups = []
    drv = FakeDeviceDriver(_engine(n_envs), [ToyEnv() for _ in range(n_envs)], tasks, work, R, K, n_action_steps=n,
                           max_steps=max_steps, num_steps_wait=wait)
    recs = drv.run()
    assert len(recs) == len(work) and not drv.active
    assert {(r.task, r.trial, r.seed) for r in recs} == set(work)
    n_early = 0
    for r in recs:
        ref = reference_episode(r.task, r.trial, r.seed, tasks[r.task], R, K, n, max_steps, wait)
        assert (r.success, r.episode_length) == (ref.success, ref.length)
        assert r.verifier_scores == ref.scores and r.selected_indices == ref.idx
        assert r.selected_instructions == ref.instr and r.step_timestamps == ref.ts
        assert len(r.execute_actions) == len(ref.acts)
        assert all(np.array_equal(a, b) for a, b in zip(r.execute_actions, ref.acts))
        n_early += r.success
    assert drv.decisions == sum(sum(s is not None for s in r.verifier_scores) for r in recs)
    if n_envs > 1:
        assert max(FakeDeviceDriver.groups) > 1 and drv.batched_calls < drv.decisions   # decisions really share launches
    if max_steps == 41:
        assert 0 < n_early < len(work)   # both kinds of episode end are exercised


def test_instruction_swap_indexes_the_prompt_cache():
    tp = _prompts(8)
    assert tp.prompt_rows(0, 8) == [0, 1, 2, 3, 4, 5, 6, 7]
    assert tp.prompt_rows(5, 8) == [5, 1, 2, 3, 4, 5, 6, 7]   # the chosen rephrase leads; the original is dropped (:299-302)
    assert tp.prompt_rows(3, 1) == [3]


def test_constructor_rejects_what_the_engine_cannot_hold():
    tasks = [_prompts(4)]
    with pytest.raises(ValueError):
        EpisodeBatchDriver(_engine(2), [ToyEnv() for _ in range(3)], tasks, [], 4, 3)
    with pytest.raises(ValueError):
        EpisodeBatchDriver(_engine(2), [ToyEnv()], tasks, [], 9, 3)
    with pytest.raises(ValueError):
        EpisodeBatchDriver(_engine(2), [ToyEnv()], tasks, [], 5, 3)          # the task has only 4 instructions
    with pytest.raises(ValueError):
        EpisodeBatchDriver(_engine(2), [ToyEnv()], tasks, [], 4, 3, num_steps_wait=2)


def _rank_worker(rank, world, port, work, R, K, n, max_steps, q):
    import os
    import torch.distributed as dist
    from cover_vla_b200.episodes import gather_records, shard_work
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tasks = [_prompts(R), _prompts(R + 2)]
        FakeDeviceDriver.groups = []
        mine = shard_work(work, world, rank)
        drv = FakeDeviceDriver(_engine(2), [ToyEnv(), ToyEnv()], tasks, mine, R, K, n_action_steps=n, max_steps=max_steps)
        drv.run()
        allrecs = gather_records(drv.finished)
        q.put((rank, len(mine), [(r.task, r.trial, r.seed, r.success, r.episode_length, r.selected_indices,
                                  [a.tolist() for a in r.execute_actions]) for r in allrecs]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_episode_parallel_ranks_gather_the_same_records(world):
    """configs[4] across ranks: every rank drives its round-robin share of the episodes; the gathered, ordered records are
    identical on every rank and equal each episode stepped alone (no data-path collective, one gather at the end)."""
    import socket
    import torch.multiprocessing as mp
    R, K, n, max_steps = 4, 3, 4, 25
    work = [(task, trial, 1000 + trial) for task in range(2) for trial in range(4)]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, world, port, work, R, K, n, max_steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[1] for r in res) == len(work)
    tasks = [_prompts(R), _prompts(R + 2)]
    first = res[0][2]
    assert [(t, tr, sd) for t, tr, sd, *_ in first] == sorted(work)
    for rank, _, recs in res:
        assert recs == first
    for task, trial, seed, success, length, idx, acts in first:
        ref = reference_episode(task, trial, seed, tasks[task], R, K, n, max_steps, 0)
        assert (success, length, idx) == (ref.success, ref.length, ref.idx)
        assert acts == [a.tolist() for a in ref.acts]
