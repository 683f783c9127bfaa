"""Episode-batched CoVer driver (BASELINE.json configs[4]).

The reference evaluates ONE environment at a time (run_simpler_eval_with_openpi.py:190-449): every n_action_steps ticks it
samples N = R x K action chunks for the current frame (:296-326), formats them for the verifier (:338-341), asks the
verifier about candidate 0 and - when its score is below 0.1 - about all N (:344-366), executes the winner's first action
with the gripper voted by its K-sample group (:368-391), keeps the winner's remaining actions for the next ticks (:393-400,
:411-417), appends the executed action in verifier format to the history (:425-433) and makes the winning instruction the
task description of the next decision (:409).

`EpisodeBatchDriver` keeps exactly that per-episode state machine for SEVERAL environments that share one GPU and, per tick,
hands the observations of every environment that is at a decision tick to ONE `cvb_cover_step_batch` (the weights are
streamed once for all of them; `tests/test_batch_gpu.py`: B observations == B single calls bit for bit).  Everything between
the simulator frame and the executed action runs on the device: both image pre-processing chains from one uint8 H2D copy per
frame (`preprocess.py`), the decision, the gate, the execution-format action of all n_action_steps steps, the gripper vote
and the verifier-format history rows; one D2H read per group of decisions returns them.

Per-task prompt cache: a task's instructions (the original + its rephrases) are tokenised ONCE (`TaskPrompts`) for the
policy and for the verifier and stay on the device; the prompt set of a decision - `[task_description] + rephrases[:R-1]`
(:299-302) - is an index gather, and the instruction swap of :409 moves one index.

The simulator itself, its proprioception adapter and the rephrase generation are out of scope: an
environment is any object with the four methods of `EpisodeEnv`; a slot's environment object is re-used for whatever
(task, trial, seed) comes next, so `reset` receives the task.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Callable, Iterable, Protocol

import numpy as np
import torch

from . import preprocess
from .cover import BRIDGE_ACTION_P01, BRIDGE_ACTION_P99, BatchedCoverStep, CoverInputs, execution_action
from .engine import Engine

MAX_PAST = 6  # run_simpler_eval_with_openpi.py:333 / eval_utils.py:209: the verifier sees at most 6 executed actions


class EpisodeEnv(Protocol):
    """What the driver needs from a simulator wrapper (get_simpler_env + the adapter's proprioception, which stay host code)."""

    def reset(self, task: int, seed: int): ...           # -> obs; the environment of `task` (:194, :231)
    def step(self, action: np.ndarray): ...              # f64 [7] -> (obs, done: bool)     (:436)
    def frame(self, obs) -> np.ndarray: ...              # uint8 [H, W, 3]                  (:268)
    def state(self, obs) -> np.ndarray: ...              # float [<= max_state_dim]         (preprocess_adapter.preprocess, :284)


@dataclass
class TaskPrompts:
    """The tokenised instructions of one task, resident on the device for all its episodes.

    Row 0 is the original instruction, rows 1.. the rephrases in the order of the reference's `ert_rephrases` list
    (:213); `pi0_tokens` / `pi0_len` are the policy tokenisation (PI0Policy.prepare_language: "\\n" appended, right-padded
    to tokenizer_max_length), `vf_tokens` the verifier's (open_clip tokenizer, context_length)."""
    instructions: list[str]
    pi0_tokens: torch.Tensor   # i64 [P, L]
    pi0_len: torch.Tensor      # i32 [P]
    vf_tokens: torch.Tensor    # i64 [P, ctx]
    lang_len_max: int | None = None

    @staticmethod
    def build(instructions: list[str], pi0_tokenize: Callable, vf_tokenize: Callable, device) -> "TaskPrompts":
        """pi0_tokenize(list[str]) -> (i64 [P, L], bool [P, L]); vf_tokenize(list[str]) -> i64 [P, ctx].  One tokenizer run
        and one H2D copy per task."""
        tok, mask = pi0_tokenize(list(instructions))
        lens = mask.sum(dim=1).to(torch.int32)
        return TaskPrompts(list(instructions), tok.to(device=device, dtype=torch.int64).contiguous(),
                           lens.to(device).contiguous(), vf_tokenize(list(instructions)).to(device=device, dtype=torch.int64).contiguous(),
                           int(lens.max()))

    def prompt_rows(self, current: int, R: int) -> list[int]:
        """Instruction ids of `[task_description] + rephrased_list[:R - 1]` (:299-302) when the task description is
        instruction `current`."""
        return [current] + list(range(1, R))


@dataclass
class EpisodeRecord:
    """episode_data of the reference (:237-246), one per finished episode."""
    task: int
    trial: int
    seed: int
    success: bool = False
    episode_length: int = 0
    verifier_scores: list = field(default_factory=list)
    selected_instructions: list = field(default_factory=list)
    execute_actions: list = field(default_factory=list)
    step_timestamps: list = field(default_factory=list)
    selected_indices: list = field(default_factory=list)


@dataclass
class _Slot:
    env: EpisodeEnv
    task: int = -1
    current: int = 0          # instruction id of task_description
    t: int = 0
    obs: object = None
    history: list = field(default_factory=list)    # executed actions in verifier format, f32 [7] each (:425-433)
    queue_exec: list = field(default_factory=list)  # execution-format actions of the winner's remaining steps, f64 [7]
    queue_hist: list = field(default_factory=list)  # their verifier-format rows
    record: EpisodeRecord | None = None
    active: bool = False


class EpisodeBatchDriver:
    """Steps `len(envs)` environments in lock step; see the module docstring.

    engine: a finalized Engine built with max_observations >= the number of environments, max_rephrases >= R,
    max_samples >= K.  tasks: one TaskPrompts per task id.  work: iterable of (task id, trial index, seed) - the
    reference's loops over tasks / trials / seeds (:190-231) flattened by the caller."""

    def __init__(self, engine: Engine, envs: list[EpisodeEnv], tasks: list[TaskPrompts], work: Iterable[tuple[int, int, int]],
                 R: int, K: int, n_action_steps: int | None = None, max_steps: int = 150, num_steps_wait: int = 0,
                 gate_threshold: float = 0.1, use_verifier_gate: bool = True, noise_std: float = 1.0,
                 noise_fn: Callable | None = None, dummy_action: np.ndarray | None = None, seed: int = 0,
                 p01=BRIDGE_ACTION_P01, p99=BRIDGE_ACTION_P99):
        cfg = engine.cfg
        self.engine, self.tasks, self.R, self.K = engine, tasks, R, K
        self.n_action_steps = n_action_steps or cfg.chunk_size
        if self.n_action_steps > cfg.chunk_size:
            raise ValueError("n_action_steps exceeds the policy's chunk_size")
        if num_steps_wait % self.n_action_steps != 0:
            # the reference only samples when t % n_action_steps == 0 (:322): a wait that is not a multiple of it pops an
            # empty queue there
            raise ValueError("num_steps_wait must be a multiple of n_action_steps")
        if len(envs) > max(1, cfg.max_observations):
            raise ValueError(f"{len(envs)} environments, engine built for max_observations = {cfg.max_observations}")
        if R > cfg.max_rephrases or K > cfg.max_samples:
            raise ValueError("R / K exceed the engine's max_rephrases / max_samples")
        for tp in tasks:
            if tp.pi0_tokens.shape[0] < R:
                raise ValueError("a task needs at least R instructions (the original + R - 1 rephrases)")
        self.max_steps, self.num_steps_wait = max_steps, num_steps_wait
        self.gate_threshold, self.use_gate = gate_threshold, use_verifier_gate
        self.noise_std, self.noise_fn = noise_std, noise_fn
        self.dummy_action = np.array([0, 0, 0, 0, 0, 0, -1], dtype=np.float64) if dummy_action is None else dummy_action
        self.p01, self.p99 = p01, p99
        self.step_fn = BatchedCoverStep(engine, K, n_future=self.n_action_steps, p01=p01, p99=p99)
        self._gen, self._seed = None, seed  # created on first use (the noise of sample_noise)
        self._work = iter(work)
        self.slots = [_Slot(env=e) for e in envs]
        self.finished: list[EpisodeRecord] = []
        self.decisions = 0          # decisions taken (one per environment per n_action_steps ticks)
        self.batched_calls = 0      # cvb_cover_step_batch launches that served them
        self._last_group_key = None
        self._prompt_sets: dict = {}
        for s in self.slots:
            self._start_next(s)

    # ------------------------------------------------------------------ episode bookkeeping
    def _start_next(self, s: _Slot) -> None:
        nxt = next(self._work, None)
        if nxt is None:
            s.active = False
            return
        task, trial, seed = nxt
        s.task, s.current, s.t = task, 0, 0            # every trial starts from the original instruction (:220-223)
        s.history, s.queue_exec, s.queue_hist = [], [], []
        s.record = EpisodeRecord(task=task, trial=trial, seed=seed)
        s.obs = s.env.reset(task, seed)
        s.active = True

    def _finish(self, s: _Slot, success: bool) -> None:
        s.record.success, s.record.episode_length = bool(success), s.t
        self.finished.append(s.record)
        self._start_next(s)

    @property
    def active(self) -> bool:
        return any(s.active for s in self.slots)

    # ------------------------------------------------------------------ one group of decisions on the device
    def _prompt_set(self, task: int, current: int):
        """Policy tokens / lengths of `[task_description] + rephrases[:R - 1]` - a device-side gather, kept per
        (task, task description): no tokenizer run and no H2D copy on the decision path."""
        hit = self._prompt_sets.get((task, current))
        if hit is None:
            tp = self.tasks[task]
            rows = torch.tensor(tp.prompt_rows(current, self.R), dtype=torch.int64, device=tp.pi0_tokens.device)
            hit = (tp.pi0_tokens.index_select(0, rows).contiguous(), tp.pi0_len.index_select(0, rows).contiguous())
            self._prompt_sets[(task, current)] = hit
        return hit

    def _host_inputs(self, s: _Slot, num_past: int) -> dict:
        """What one environment contributes to a decision, still on the host: the raw frame, the padded state
        (pad_vector, modeling_pi0.py:438-441) and the last num_past executed actions in verifier format (:333, :338-341)."""
        state = np.zeros(self.engine.cfg.max_state_dim, dtype=np.float32)
        st = np.asarray(s.env.state(s.obs), dtype=np.float32).reshape(-1)
        state[: st.shape[0]] = st
        past = np.stack(s.history[-num_past:]).astype(np.float32) if num_past > 0 else None
        return {"frame": np.ascontiguousarray(s.env.frame(s.obs)), "state": state, "past": past}

    def _inputs(self, group: list[_Slot]) -> CoverInputs:
        dev, cfg = self.engine.device, self.engine.cfg
        xs = []
        num_past = min(len(group[0].history), MAX_PAST)
        for s in group:
            h = self._host_inputs(s, num_past)
            frame = torch.from_numpy(h["frame"]).to(dev, non_blocking=True)  # ONE uint8 H2D copy feeds both image chains
            tp = self.tasks[s.task]
            lang_tokens, lang_len = self._prompt_set(s.task, s.current)
            shape = (self.R * self.K, cfg.chunk_size, cfg.max_action_dim)
            if self.noise_fn is not None:
                noise = self.noise_fn(s.record, s.t, shape).to(device=dev, dtype=torch.float32)
            else:  # sample_noise (modeling_pi0.py:502-510)
                if self._gen is None:
                    self._gen = torch.Generator(device=dev).manual_seed(self._seed)
                noise = torch.randn(shape, generator=self._gen, device=dev, dtype=torch.float32) * self.noise_std
            past = None if h["past"] is None else torch.from_numpy(h["past"]).to(dev)
            xs.append(CoverInputs(image=preprocess.policy_image(frame, cfg.vis_image)[0],
                                  lang_tokens=lang_tokens, lang_len=lang_len,
                                  state=torch.from_numpy(h["state"]).to(dev), noise=noise,
                                  vf_image=preprocess.verifier_image_from_raw(frame, cfg.vf_image)[0],
                                  vf_tokens=tp.vf_tokens[s.current], past=past, lang_len_max=tp.lang_len_max))
        return BatchedCoverStep.stack(xs)

    def _decide(self, group: list[_Slot]) -> list[dict]:
        """One cvb_cover_step_batch for the group (same history length) + the device-side post-processing; ONE D2H read.
        Per environment: idx, score, exec f64 [n_action_steps, 7], hist f32 [n_action_steps, 7]."""
        n, H = self.n_action_steps, self.engine.cfg.vf_history
        # per-task prompt cache, device side: the verifier's text tower is skipped when this call sees the same
        # environments with the same instructions in the same batch slots as the previous one
        key = tuple((id(s), s.task, s.current) for s in group)
        self.step_fn.hold_text = key == self._last_group_key
        self._last_group_key = key
        actions, traj, scores, gmean, bidx, bscore = self.step_fn.sample_and_score(self._inputs(group))
        B = actions.shape[0]
        if self.use_gate:  # :344-366: candidate 0 under the current task description is kept when it clears the threshold
            use0 = scores[:, 0] >= self.gate_threshold
        else:
            use0 = torch.zeros(B, dtype=torch.bool, device=actions.device)
        idx = torch.where(use0, torch.zeros_like(bidx), bidx)
        score = torch.where(use0, scores[:, 0], bscore)
        rows = []
        for b in range(B):
            ib = idx[b:b + 1].to(torch.int32).contiguous()
            # step 0: gripper voted by the winner's K-sample group (:368-391); later steps come out of the queue and are
            # converted alone (:411-417) = a group of one
            ex = [execution_action(actions[b], ib, self.K if i == 0 else 1, i, self.p01, self.p99)[0] for i in range(n)]
            hist = traj[b].index_select(0, ib.to(torch.int64))[0, H - n:]      # the winner's verifier-format rows
            rows.append(torch.cat([ib.to(torch.float64), score[b:b + 1].to(torch.float64), torch.cat(ex),
                                   hist.reshape(-1).to(torch.float64)]))
        packed = torch.stack(rows).cpu().numpy()
        self.decisions += B
        self.batched_calls += 1
        return [{"idx": int(p[0]), "score": float(p[1]), "exec": p[2:2 + 7 * n].reshape(n, 7).copy(),
                 "hist": p[2 + 7 * n:].reshape(n, 7).astype(np.float32)} for p in packed]

    # ------------------------------------------------------------------ the tick
    def tick(self) -> None:
        """Advance every active environment by one simulator step."""
        live = [s for s in self.slots if s.active]
        acts: dict[int, np.ndarray] = {}
        for s in live:
            if s.t < self.num_steps_wait:                                          # :261-265
                acts[id(s)] = self.dummy_action
        deciding = [s for s in live if id(s) not in acts and s.t % self.n_action_steps == 0]
        decided: set[int] = set()
        # one batched call per history length (0 .. 6; every episode reaches 6 after two decisions)
        for _, grp in itertools.groupby(sorted(deciding, key=lambda s: min(len(s.history), MAX_PAST)),
                                        key=lambda s: min(len(s.history), MAX_PAST)):
            grp = list(grp)
            for s, r in zip(grp, self._decide(grp)):
                tp = self.tasks[s.task]
                g = r["idx"] // self.K
                if g > 0:                                                          # :366, :409 task_description = max_instruction
                    s.current = tp.prompt_rows(s.current, self.R)[g]
                s.queue_exec, s.queue_hist = list(r["exec"]), list(r["hist"])
                s.record.verifier_scores.append(r["score"])
                s.record.selected_indices.append(r["idx"])
                decided.add(id(s))
        for s in live:
            if id(s) in acts:
                continue
            if id(s) not in decided:
                s.record.verifier_scores.append(None)                             # :418-422
                s.record.selected_indices.append(None)
            ex, hi = s.queue_exec.pop(0), s.queue_hist.pop(0)
            s.record.selected_instructions.append(self.tasks[s.task].instructions[s.current])
            s.record.execute_actions.append(ex.copy())
            s.record.step_timestamps.append(s.t)
            s.history.append(hi)                                                   # :425-433
            acts[id(s)] = ex
        for s in live:
            s.obs, done = s.env.step(acts[id(s)])                                  # :436
            waiting = s.t < self.num_steps_wait
            if done and not waiting:                                               # :438-441
                self._finish(s, True)
                continue
            s.t += 1
            if s.t >= self.max_steps + self.num_steps_wait:                        # :259
                self._finish(s, False)

    def run(self, max_ticks: int | None = None) -> list[EpisodeRecord]:
        ticks = 0
        while self.active and (max_ticks is None or ticks < max_ticks):
            self.tick()
            ticks += 1
        return self.finished


# ----------------------------------------------------------------------------------------------------
# episode-parallel over ranks (BASELINE.json configs[4]: 64 observations over 8 GPUs)
# ----------------------------------------------------------------------------------------------------
def shard_work(work: Iterable[tuple[int, int, int]], world_size: int, rank: int) -> list[tuple[int, int, int]]:
    """Round-robin share of the (task, trial, seed) list for one rank: episodes are independent, so there is no data-path
    collective - every rank drives its own environments with its own EpisodeBatchDriver."""
    return [w for i, w in enumerate(work) if i % world_size == rank]


def gather_records(records: list[EpisodeRecord], group=None) -> list[EpisodeRecord]:
    """The one collective of the episode-parallel mode: every rank's finished episodes, gathered once at the end (host
    objects over the process group's backend) and ordered by (task, trial, seed) so the result does not depend on the
    number of ranks."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(records, key=lambda r: (r.task, r.trial, r.seed))
    parts: list = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, records, group=group)
    return sorted((r for part in parts for r in part), key=lambda r: (r.task, r.trial, r.seed))
