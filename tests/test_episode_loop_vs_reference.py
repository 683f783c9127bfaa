"""Episode-batched driver against the reference's OWN episode loop, executed unmodified.

tests/test_episodes.py checks `EpisodeBatchDriver` against a hand-written restatement of the loop.  Here the loop is the
reference's: the statements of one trial of `eval_simpler` - `env.reset` to `episode_data['episode_length'] = t`,
CoVer_VLA/inference/experiments/robot/simpler/run_simpler_eval_with_openpi.py:231-455 - are AST-extracted and executed as they
are, together with the reference's real `process_inputs` / `convert_maniskill_with_bridge_adapter` (eval_utils.py:138-221) and
the real adapter methods they call (`postprocess`, `postprocess_gripper`, `postprocess_verifier`,
`postprocess_gripper_verifier`, `denormalize_bound`, `euler2axangle`).  Only what cannot exist offline is replaced: the
simulator (a deterministic toy whose frames depend on every action it was given), the policy (raw action chunks = a hash of
everything `select_action` is handed) and the verifier (each candidate's score = a hash of exactly what the ensemble would
see: frame, instruction, the padded float32 trajectory; the reference's selection rule applied to those scores).

The driver side replaces the device (`_decide`) by the SAME toy policy / toy scores pushed through the oracle's restatements
of the formatting kernels (oracle/exec_action_oracle.py - the CUDA kernels are bit-exact against them, tests/test_format_traj.py,
tests/test_exec_action.py).  Everything else - decision ticks, the 0.1 gate, the winner's queue, the gripper vote, the <= 6-row
float history tail, the instruction swap, termination, slots taking the next episode - is the driver's own bookkeeping, and
must reproduce the reference loop's records exactly, for several environments sharing the batched calls.

CPU, authoring container only (skipped where /root/reference is absent)."""
import ast
import hashlib
import textwrap
import types
from collections import deque
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present")

DRIVER = Path("/root/reference/CoVer_VLA/inference/experiments/robot/simpler/run_simpler_eval_with_openpi.py")
H, W = 12, 16


# ----------------------------------------------------------------------------------------------------------------------
# toys shared by both sides
# ----------------------------------------------------------------------------------------------------------------------
class ToyEnv:
    """Deterministic stand-in for the simulator: what it shows depends on the task, the seed and every action so far."""

    def reset(self, task, seed):
        self.k, self.acc = 0, hashlib.sha256(str((task, seed)).encode()).digest()
        return {"k": self.k, "acc": self.acc}

    def step(self, action):
        a = np.asarray(action, dtype=np.float64)
        assert a.shape == (7,)
        self.acc = hashlib.sha256(self.acc + a.tobytes()).digest()
        self.k += 1
        done = self.acc[0] < 2 and self.k > 5   # ~0.8 % per step: some episodes end early, others reach max_steps
        return {"k": self.k, "acc": self.acc}, done

    def frame(self, obs):
        rng = np.random.default_rng(int.from_bytes(obs["acc"][:8], "little"))
        return rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)

    def state(self, obs):
        rng = np.random.default_rng(int.from_bytes(obs["acc"][8:16], "little"))
        return rng.normal(size=7).astype(np.float32)


def toy_policy(frame, state, rows, task, K, n):
    """Raw action chunks f32 [R*K, n, 7] (policy output range, gripper in [0, 1]) from everything select_action is handed."""
    h = hashlib.sha256(np.ascontiguousarray(frame).tobytes() + np.asarray(state, dtype=np.float32).tobytes() +
                       str((list(rows), task)).encode())
    rng = np.random.default_rng(int.from_bytes(h.digest()[:8], "little"))
    a = rng.uniform(-1.3, 1.3, size=(len(rows) * K, n, 7)).astype(np.float32)
    a[:, :, 6] = rng.uniform(0.0, 1.0, size=a.shape[:2]).astype(np.float32)
    return a


def toy_score(frame, instruction_id, traj_f32):
    """One candidate's score from what the ensemble sees of it; ~1/3 of the decisions clear the 0.1 gate."""
    h = hashlib.sha256(np.ascontiguousarray(frame).tobytes() + str(instruction_id).encode() +
                       np.ascontiguousarray(traj_f32, dtype=np.float32).tobytes())
    return float(np.random.default_rng(int.from_bytes(h.digest()[:8], "little")).uniform(-0.3, 0.2))


def select(scores, K):
    """efficient_ensemble_merged.py:417-447 (first maximum wins, like torch.max)."""
    s = np.asarray(scores, dtype=np.float64).reshape(-1, K)
    g = int(np.argmax(s.mean(axis=1)))
    k = int(np.argmax(s[g]))
    return g * K + k, float(s[g, k])


# ----------------------------------------------------------------------------------------------------------------------
# the reference side: its own lines
# ----------------------------------------------------------------------------------------------------------------------
def _reference_functions():
    """process_inputs / convert_maniskill_with_bridge_adapter with an adapter carrying the reference's unmodified methods."""
    from oracle.make_golden_exec import reference_adapter, _method_source, REF
    from oracle.make_golden_format import EVAL_UTILS, _function_source, reference_padding
    adapter, stats = reference_adapter()  # postprocess, postprocess_gripper, denormalize_bound (+ the real geometry.py)
    simpler = REF / "INT-ACT/src/experiments/env_adapters/simpler.py"
    ns = {"np": np}
    exec(_method_source(simpler, "SimplerAdapter", "postprocess_verifier"), ns)
    exec(_method_source(simpler, "BridgeSimplerAdapter", "postprocess_gripper_verifier"), ns)
    type(adapter).postprocess_verifier = ns["postprocess_verifier"]
    type(adapter).postprocess_gripper_verifier = ns["postprocess_gripper_verifier"]
    mod = {"np": np, "create_bridge_adapter_wrapper": lambda temp: adapter}
    exec(_function_source(EVAL_UTILS, "convert_maniskill_with_bridge_adapter"), mod)
    exec(_function_source(EVAL_UTILS, "process_inputs"), mod)
    return mod["process_inputs"], mod["convert_maniskill_with_bridge_adapter"], reference_padding, stats


def _reference_trial_function(namespace):
    """One trial of eval_simpler as a function: the statements from `obs, reset_info = env.reset(...)` (:231) to
    `episode_data['episode_length'] = t` (:455), unmodified."""
    src = DRIVER.read_text()
    loop = next(n for n in ast.walk(ast.parse(src))
                if isinstance(n, ast.For) and isinstance(n.target, ast.Name) and n.target.id == "trail_idx")
    texts = [ast.get_source_segment(src, s) for s in loop.body]
    first = next(i for i, t in enumerate(texts) if t.startswith("obs, reset_info = env.reset("))
    last = next(i for i, t in enumerate(texts) if t.startswith("episode_data['episode_length'] = t"))
    assert loop.body[first].lineno == 231 and loop.body[last].lineno == 455, (loop.body[first].lineno, loop.body[last].lineno)
    lines = src.splitlines()[loop.body[first].lineno - 1:loop.body[last].end_lineno]   # whole source lines, as written
    body = textwrap.indent(textwrap.dedent("\n".join(lines)), "    ")
    fn = ("def reference_trial(cfg, env, seeds, task_description, original_task_description, rephrased_list, pi0_policy,\n"
          "                    ensemble_model, preprocess_adapter, action_noise_std, log_file, task_episodes=0,\n"
          "                    task_successes=0, total_successes=0, total_episodes=0, action_queue=None):\n" + body +
          "\n    return episode_data\n")
    exec(compile(fn, str(DRIVER), "exec"), namespace)
    return namespace["reference_trial"]


class _Silent:
    def write(self, *_):
        pass


class _Bar:
    def __init__(self, *a, **k):
        pass

    def update(self, *_):
        pass

    def set_description(self, *_):
        pass

    def close(self):
        pass


def reference_episode(task, trial_seed, instructions, R, K, n, wait):
    process_inputs, convert, reference_padding, stats = _reference_functions()
    toy = ToyEnv()

    class Env:  # the gym-style surface the loop uses (:231, :262, :436)
        def reset(self, seed):
            return toy.reset(task, seed), {}

        def step(self, action):
            obs, done = toy.step(np.asarray(action, dtype=np.float64))
            return obs, 0.0, done, False, {}

    class Adapter:  # preprocess_adapter.preprocess (:279-284): frame and proprioception to tensors
        def preprocess(self, d):
            return {"observation.images.top": torch.from_numpy(d["observation.images.top"].copy())[None],
                    "observation.state": torch.from_numpy(toy.state(d["observation.state"]))[None], "task": d["task"]}

    def select_action(observation, noise_std=1.0):  # PI0Policy.select_action: a deque of n tensors [N, 7]
        tasks = observation["task"]
        assert len(tasks) == R * K and all(tasks[i] == tasks[i - i % K] for i in range(R * K))
        rows = [instructions.index(t) for t in tasks[::K]]
        raw = toy_policy(observation["observation.images.top"][0].numpy(), observation["observation.state"][0].numpy(),
                         rows, task, K, n)
        assert torch.equal(observation["observation.images.top"][0], observation["observation.images.top"][-1])
        return deque([torch.from_numpy(np.ascontiguousarray(raw[:, i])) for i in range(n)], maxlen=n)

    class Ensemble:  # compute_max_similarity_scores_batch -> (max_score, max_instruction, max_action_history, index tensor)
        def compute_max_similarity_scores_batch(self, images, instructions, all_action_histories, cfg_repeat_language_instructions=1):
            traj = reference_padding(all_action_histories).numpy()        # the reference's own -5 padding + float32 cast
            scores = [toy_score(images[0], globals_instr.index(instructions[0]), traj[c]) for c in range(len(traj))]
            gidx, best = select(scores, cfg_repeat_language_instructions)
            g = gidx // cfg_repeat_language_instructions
            all_same = len(set(instructions)) == 1
            mi = instructions[0] if (all_same and len(images) > 1) else \
                instructions[min(g * cfg_repeat_language_instructions, len(instructions) - 1)]
            return best, mi, all_action_histories[gidx], torch.tensor(gidx, dtype=torch.int64)

    globals_instr = instructions
    cfg = types.SimpleNamespace(task_suite_name="simpler_widowx", num_steps_wait=wait, model_family="pi0", obs_history=1,
                                policy_batch_inference_size=K, lang_rephrase_num=R, n_action_steps=n, use_verifier=True,
                                action_ensemble_temp=-0.8)
    ns = {"np": np, "torch": torch, "deque": deque, "tqdm": types.SimpleNamespace(tqdm=_Bar), "print": lambda *a, **k: None,
          "get_simpler_dummy_action": lambda family: np.array([0, 0, 0, 0, 0, 0, -1]),
          "get_image_from_maniskill2_obs_dict": lambda env, obs: toy.frame(obs),
          "process_raw_image_to_jpg": lambda img: img, "process_inputs": process_inputs,
          "convert_maniskill_with_bridge_adapter": convert}
    trial = _reference_trial_function(ns)
    policy = types.SimpleNamespace(config=types.SimpleNamespace(device="cpu", image_features={"observation.images.top": None}),
                                   select_action=select_action)
    data = trial(cfg, Env(), iter([trial_seed]), instructions[0], instructions[0], instructions[1:R], policy, Ensemble(),
                 Adapter(), 1.0, _Silent())
    return data, stats


# ----------------------------------------------------------------------------------------------------------------------
# the driver side: EpisodeBatchDriver with the device replaced by the toys + the oracle's formatting restatements
# ----------------------------------------------------------------------------------------------------------------------
def _driver_class(stats):
    from cover_vla_b200.episodes import MAX_PAST, EpisodeBatchDriver
    from oracle import exec_action_oracle as X
    p01, p99 = stats["action"]["p01"][:6], stats["action"]["p99"][:6]

    class ToyDeviceDriver(EpisodeBatchDriver):
        groups: list = []

        def _decide(self, group):
            num_past = min(len(group[0].history), MAX_PAST)
            assert all(min(len(s.history), MAX_PAST) == num_past for s in group)
            self.groups.append(len(group))
            out, n, Hh = [], self.n_action_steps, self.engine.cfg.vf_history
            for s in group:
                hi = self._host_inputs(s, num_past)
                rows = self.tasks[s.task].prompt_rows(s.current, self.R)
                raw = toy_policy(hi["frame"], hi["state"][:7], rows, s.task, self.K, n)          # cvb_pi0_sample
                traj = X.verifier_trajectories(raw, hi["past"], Hh, p01, p99, n)                  # format_traj_kernel
                scores = [toy_score(hi["frame"], s.current, traj[c]) for c in range(len(traj))]   # verifier, current instruction
                bidx, bscore = select(scores, self.K)                                             # select_kernel
                use0 = self.use_gate and scores[0] >= self.gate_threshold                         # the gate, as _decide does
                idx, score = (0, scores[0]) if use0 else (bidx, bscore)
                ex = [X.execution_action(raw, idx, self.K if i == 0 else 1, p01, p99, step=i)[0] for i in range(n)]
                out.append({"idx": idx, "score": float(score), "exec": np.stack(ex), "hist": traj[idx, Hh - n:].astype(np.float32)})
            self.decisions += len(group)
            self.batched_calls += 1
            return out

    return ToyDeviceDriver


@pytest.mark.parametrize("n_envs,wait,max_steps_note", [(1, 0, "one environment"), (3, 4, "three environments, 4 wait steps")])
def test_driver_reproduces_the_reference_loop_executed_unmodified(n_envs, wait, max_steps_note):
    from cover_vla_b200.episodes import TaskPrompts
    R, K, n = 4, 3, 4
    P = R + 1
    tasks = [TaskPrompts([f"task {t} instruction {i}" for i in range(P)], torch.zeros(P, 8, dtype=torch.int64),
                         torch.full((P,), 3, dtype=torch.int32), torch.zeros(P, 4, dtype=torch.int64), 3) for t in range(2)]
    work = [(task, trial, 1000 + trial) for task in range(2) for trial in range(3)]
    ref = {}
    stats = None
    for task, trial, seed in work:
        ref[(task, trial, seed)], stats = reference_episode(task, seed, tasks[task].instructions, R, K, n, wait)
    cfg = types.SimpleNamespace(chunk_size=4, max_observations=n_envs, max_rephrases=8, max_samples=5, max_state_dim=32,
                                vf_history=10, max_action_dim=32, vis_image=224, vf_image=384)
    Driver = _driver_class(stats)
    Driver.groups = []
    drv = Driver(types.SimpleNamespace(cfg=cfg, device="cpu"), [ToyEnv() for _ in range(n_envs)], tasks, work, R, K,
                 n_action_steps=n, max_steps=150, num_steps_wait=wait)   # 150: the reference's constant for simpler suites (:250)
    recs = drv.run()
    assert {(r.task, r.trial, r.seed) for r in recs} == set(work)
    n_gate_pass = n_swaps = n_early = 0
    for r in recs:
        e = ref[(r.task, r.trial, r.seed)]
        assert (r.success, r.episode_length) == (bool(e["success"]), e["episode_length"])
        assert r.step_timestamps == e["step_timestamps"]
        assert r.verifier_scores == e["verifier_scores"]
        assert r.selected_instructions == e["selected_instructions"]
        assert len(r.execute_actions) == len(e["execute_actions"])
        for a, b in zip(r.execute_actions, e["execute_actions"]):
            assert a.dtype == b.dtype == np.float64 and np.array_equal(a, b)
        n_gate_pass += sum(i == 0 for i in r.selected_indices if i is not None)
        n_swaps += len(set(r.selected_instructions)) - 1
        n_early += r.success
    # the comparison exercised both gate outcomes, instruction swaps and both kinds of episode end
    assert n_gate_pass > 0 and n_swaps > 0 and 0 < n_early < len(work)
    if n_envs > 1:
        assert max(Driver.groups) > 1 and drv.batched_calls < drv.decisions
