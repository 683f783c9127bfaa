"""One CoVer decision on the device: sample N = R*K action chunks (pi0) -> verifier-format trajectories ->
ensemble scores -> group-mean / argmax, with no host round trip in between.

This is the body of run_simpler_eval_with_openpi.py:322-409 minus the simulator: `select_action` (:324),
`process_inputs(verifier_action=True)` (:338), the 1-candidate gate (:344-352) and the N-candidate call
(:355-363).  The gate needs no second pass: its score is scores[0] of the same launch.

Multi-GPU: candidates shard by REPHRASE - rank g owns rephrases
[g*R/G, (g+1)*R/G) and all their K samples, so every prefix is computed exactly once globally and group
means stay rank-local.  The only exchange is an all-gather of N/G fp32 scores (+ the N/G x chunk x 7
actions so every rank can return the winner); every rank then runs the same select kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

import ctypes as C
import torch

from . import _lib
from .engine import Engine

# INT-ACT/config/dataset/bridge_statistics.json "action" p01 / p99 (first 6 dims; the gripper is not scaled)
BRIDGE_ACTION_P01 = (-0.028539552688598632, -0.041432044506073, -0.025977383628487588,
                     -0.08020886614918708, -0.09213060349225997, -0.2054861941933632)
BRIDGE_ACTION_P99 = (0.028122276067733765, 0.040630316659808145, 0.03994889184832546,
                     0.08121915772557152, 0.07724379181861864, 0.20214049845933896)


def format_trajectories(actions, past, history: int, n_future: int, p01=BRIDGE_ACTION_P01, p99=BRIDGE_ACTION_P99,
                        out=None):
    """actions f32 [N, chunk, stride>=7] (device) -> verifier trajectories f32 [N, history, 7] (device)."""
    lib = _lib.load()
    N, chunk, stride = actions.shape
    num_past = 0 if past is None else past.shape[0]
    if out is None:
        out = torch.empty(N, history, 7, dtype=torch.float32, device=actions.device)
    a = (C.c_double * 6)(*p01)
    b = (C.c_double * 6)(*p99)
    lib.cvb_format_trajectories.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                            C.POINTER(C.c_double), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p]
    with torch.cuda.device(actions.device):
        _lib.check(lib.cvb_format_trajectories(_lib.ptr(actions), N, chunk, stride, a, b, _lib.ptr(past), num_past,
                                               history, n_future, _lib.ptr(out), _lib.stream_ptr()))
    return out


def execution_action(actions, best_idx, K: int, step: int = 0, p01=BRIDGE_ACTION_P01, p99=BRIDGE_ACTION_P99):
    """Execution-format action of candidate best_idx (device i32 tensor) + gripper vote of its K-sample group.

    actions f32 [N, chunk, stride>=7] (device) -> (f64 [7] = xyz, axis*angle, gripper +-1; i32 [2] = close / open votes).
    Mirrors run_simpler_eval_with_openpi.py:368-391 (process_inputs(verifier_action=False) + the vote)."""
    lib = _lib.load()
    N, chunk, stride = actions.shape
    out = torch.empty(7, dtype=torch.float64, device=actions.device)
    votes = torch.empty(2, dtype=torch.int32, device=actions.device)
    a = (C.c_double * 6)(*p01)
    b = (C.c_double * 6)(*p99)
    lib.cvb_execution_action.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
    with torch.cuda.device(actions.device):
        _lib.check(lib.cvb_execution_action(_lib.ptr(actions), N, chunk, stride, a, b, _lib.ptr(best_idx), int(K), int(step),
                                            _lib.ptr(out), _lib.ptr(votes), _lib.stream_ptr()))
    return out, votes


@dataclass
class CoverInputs:
    """Device-resident inputs of one decision."""
    image: torch.Tensor        # f32 [3, 224, 224] in [-1, 1]
    lang_tokens: torch.Tensor  # i64 [R, L] (one row per unique rephrase)
    lang_len: torch.Tensor     # i32 [R]
    state: torch.Tensor        # f32 [max_state_dim]
    noise: torch.Tensor        # f32 [R*K, chunk, max_action_dim]
    vf_image: torch.Tensor     # f32 [3, 384, 384] (verifier preprocessing)
    vf_tokens: torch.Tensor    # i64 [ctx] (the CURRENT task description)
    past: torch.Tensor | None = None  # f32 [num_past, 7] verifier-format action history tail
    lang_len_max: int | None = None   # host-known bound on valid tokens per prompt (lets the prefix skip padding rows)


class CoverStep:
    def __init__(self, engine: Engine, samples_per_rephrase: int, n_future: int | None = None,
                 p01=BRIDGE_ACTION_P01, p99=BRIDGE_ACTION_P99):
        self.engine = engine
        self.K = samples_per_rephrase
        self.n_future = n_future or engine.cfg.chunk_size
        self.p01, self.p99 = p01, p99
        # the verifier's image/text side does not depend on the sampled actions: it runs on a second stream,
        # concurrently with the (latency-bound, SM-underfilling) denoise loop of the sampler
        self.overlap_context = True
        # default: the whole decision is one C-ABI call / one CUDA graph (cvb_cover_step) with the verifier context forked
        # after the prefix; False = three calls (cvb_pi0_sample, cvb_format_trajectories, cvb_verifier_score)
        self.fused = True
        # per-task prompt cache: set True while the verifier instruction (x.vf_tokens) is the one of
        # the previous decision - the verifier's text tower is then skipped (cvb_verifier_hold_text); the caller owns this
        # promise, nothing is compared on the device.  Default False: every decision encodes the text, like the reference.
        self.hold_text = False
        self._side = torch.cuda.Stream(device=engine.device)

    def sample_and_score(self, x: CoverInputs):
        """Asynchronous; returns device tensors (actions, traj, scores, group_mean, best_idx, best_score)."""
        e = self.engine
        R = x.lang_tokens.shape[0]
        if self.fused:
            return e.cover_step(x.image, x.lang_tokens, x.lang_len, x.state, x.noise, self.K, x.vf_image, x.vf_tokens,
                                self.p01, self.p99, past=x.past, n_future=self.n_future, lang_len_max=x.lang_len_max,
                                hold_text=self.hold_text)
        if self.overlap_context:
            cur = torch.cuda.current_stream(e.device)
            self._side.wait_stream(cur)
            # critical path first: one graph launch for the whole sampler, then the side work
            actions = e.pi0_sample(x.image, x.lang_tokens, x.lang_len, x.state, x.noise, K=self.K, lang_len_max=x.lang_len_max)
            with torch.cuda.stream(self._side):
                e.verifier_context(x.vf_image, x.vf_tokens, hold_text=self.hold_text)
            traj = format_trajectories(actions, x.past, e.cfg.vf_history, self.n_future, self.p01, self.p99)
            cur.wait_stream(self._side)
            scores, gmean, bidx, bscore = e.verifier_score(None, None, traj, R, self.K,
                                                           recompute_context=False)
        else:
            actions = e.pi0_sample(x.image, x.lang_tokens, x.lang_len, x.state, x.noise, K=self.K, lang_len_max=x.lang_len_max)
            traj = format_trajectories(actions, x.past, e.cfg.vf_history, self.n_future, self.p01, self.p99)
            scores, gmean, bidx, bscore = e.verifier_score(x.vf_image, x.vf_tokens, traj, R, self.K, hold_text=self.hold_text)
        return actions, traj, scores, gmean, bidx, bscore

    def __call__(self, x: CoverInputs, gate_threshold: float = 0.1):
        """Full decision incl. the one D2H read the caller needs: (best_idx, best_score, winner actions [chunk, 7])."""
        actions, traj, scores, gmean, bidx, bscore = self.sample_and_score(x)
        # gate (run_simpler_eval_with_openpi.py:344-363): keep candidate 0 when its score clears the threshold
        use0 = scores[0] >= gate_threshold
        idx = torch.where(use0, torch.zeros_like(bidx[0]), bidx[0])
        score = torch.where(use0, scores[0], bscore[0])
        winner = actions.index_select(0, idx.to(torch.int64).reshape(1))[0, :, :7]
        packed = torch.cat([idx.to(torch.float32).reshape(1), score.reshape(1), winner.reshape(-1)]).cpu()
        return int(packed[0]), float(packed[1]), packed[2:].reshape(-1, 7)

    def decide_and_execute(self, x: CoverInputs, gate_threshold: float = 0.1):
        """__call__ plus the action the reference actually executes (run_simpler_eval_with_openpi.py:368-391): the
        winner's first step in execution format (xyz, axis-angle, gripper) with the gripper voted by its K-sample group,
        computed on the device.  Returns (best_idx, best_score, winner actions [chunk, 7] f32, execute_action [7] f64)."""
        actions, traj, scores, gmean, bidx, bscore = self.sample_and_score(x)
        use0 = scores[0] >= gate_threshold
        idx = torch.where(use0, torch.zeros_like(bidx[0]), bidx[0])
        score = torch.where(use0, scores[0], bscore[0])
        ex, _ = execution_action(actions, idx.to(torch.int32).reshape(1).contiguous(), self.K, 0, self.p01, self.p99)
        winner = actions.index_select(0, idx.to(torch.int64).reshape(1))[0, :, :7]
        packed = torch.cat([idx.to(torch.float64).reshape(1), score.to(torch.float64).reshape(1),
                            winner.reshape(-1).to(torch.float64), ex]).cpu()
        n = winner.numel()
        return int(packed[0]), float(packed[1]), packed[2:2 + n].to(torch.float32).reshape(-1, 7), packed[2 + n:].numpy()


class BatchedCoverStep:
    """B independent decisions per call (BASELINE.json configs[4]): the episode-batched driver
    around run_simpler_eval_with_openpi.py:190-449 steps several environments per tick and hands their observations to
    ONE cvb_cover_step_batch.  Every observation keeps its own image, state, rephrases, noise, verifier context and
    action history; the weights are streamed once for all of them."""

    def __init__(self, engine: Engine, samples_per_rephrase: int, n_future: int | None = None,
                 p01=BRIDGE_ACTION_P01, p99=BRIDGE_ACTION_P99):
        self.engine = engine
        self.K = samples_per_rephrase
        self.n_future = n_future or engine.cfg.chunk_size
        self.p01, self.p99 = p01, p99
        self.hold_text = False  # per-task prompt cache, as CoverStep.hold_text (all B instructions unchanged)

    @staticmethod
    def stack(xs: list[CoverInputs]) -> CoverInputs:
        """Stack B single-observation inputs (same R, same history length) along a new leading dimension."""
        past = None if xs[0].past is None else torch.stack([x.past for x in xs]).contiguous()
        hints = [x.lang_len_max for x in xs]
        return CoverInputs(image=torch.stack([x.image for x in xs]).contiguous(),
                           lang_tokens=torch.stack([x.lang_tokens for x in xs]).contiguous(),
                           lang_len=torch.stack([x.lang_len for x in xs]).contiguous(),
                           state=torch.stack([x.state for x in xs]).contiguous(),
                           noise=torch.stack([x.noise for x in xs]).contiguous(),
                           vf_image=torch.stack([x.vf_image for x in xs]).contiguous(),
                           vf_tokens=torch.stack([x.vf_tokens for x in xs]).contiguous(), past=past,
                           lang_len_max=None if any(h is None for h in hints) else max(hints))

    def sample_and_score(self, xb: CoverInputs):
        """xb: stacked inputs (leading dimension B).  Asynchronous; device tensors (actions [B,N,chunk,A], traj, scores
        [B,N], group_mean [B,R], best_idx [B], best_score [B])."""
        return self.engine.cover_step_batch(xb.image, xb.lang_tokens, xb.lang_len, xb.state, xb.noise, self.K, xb.vf_image,
                                            xb.vf_tokens, self.p01, self.p99, past=xb.past, n_future=self.n_future,
                                            lang_len_max=xb.lang_len_max, hold_text=self.hold_text)

    def __call__(self, xb: CoverInputs, gate_threshold: float = 0.1):
        """B decisions incl. the one D2H read: (best_idx [B], best_score [B], winner actions [B, chunk, 7]) on the host."""
        actions, traj, scores, gmean, bidx, bscore = self.sample_and_score(xb)
        B = actions.shape[0]
        use0 = scores[:, 0] >= gate_threshold  # the 1-candidate gate of run_simpler_eval_with_openpi.py:344-363
        idx = torch.where(use0, torch.zeros_like(bidx), bidx).to(torch.int64)
        score = torch.where(use0, scores[:, 0], bscore)
        winner = actions[torch.arange(B, device=actions.device), idx][:, :, :7]
        packed = torch.cat([idx.to(torch.float32).reshape(B, 1), score.reshape(B, 1), winner.reshape(B, -1)], dim=1).cpu()
        return packed[:, 0].to(torch.int64), packed[:, 1], packed[:, 2:].reshape(B, -1, 7)


# ----------------------------------------------------------------------------------------------------
# rephrase sharding across ranks
# ----------------------------------------------------------------------------------------------------
def rephrase_shard(R: int, world_size: int, rank: int) -> tuple[int, int]:
    """[start, stop) of the rephrases rank owns; contiguous, sizes differ by at most one."""
    base, rem = divmod(R, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_inputs(x: CoverInputs, K: int, world_size: int, rank: int) -> CoverInputs:
    R = x.lang_tokens.shape[0]
    a, b = rephrase_shard(R, world_size, rank)
    return CoverInputs(image=x.image, lang_tokens=x.lang_tokens[a:b].contiguous(), lang_len=x.lang_len[a:b].contiguous(),
                       state=x.state, noise=x.noise[a * K:b * K].contiguous(), vf_image=x.vf_image,
                       vf_tokens=x.vf_tokens, past=x.past, lang_len_max=x.lang_len_max)


def gather_and_select(local_scores, local_actions, R: int, K: int, select_fn, group=None):
    """All-gather the per-rank score / action slices (rephrase-major order is preserved because shards are
    contiguous) and run the same selection on every rank.  Returns (scores [N], actions [N,...], gmean, idx, score)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        scores, actions = local_scores, local_actions
    else:
        sizes = [(rephrase_shard(R, world, r)[1] - rephrase_shard(R, world, r)[0]) * K for r in range(world)]
        if len(set(sizes)) == 1:
            scores = torch.empty(R * K, dtype=local_scores.dtype, device=local_scores.device)
            dist.all_gather_into_tensor(scores, local_scores.contiguous(), group=group)
            actions = torch.empty((R * K, *local_actions.shape[1:]), dtype=local_actions.dtype, device=local_actions.device)
            dist.all_gather_into_tensor(actions, local_actions.contiguous(), group=group)
        else:  # ragged shards (R not divisible by the world size): pad every slice to the largest, gather, compact
            mx = max(sizes)

            def gather_padded(local):
                pad = torch.zeros((mx, *local.shape[1:]), dtype=local.dtype, device=local.device)
                pad[:local.shape[0]] = local
                buf = torch.empty((world * mx, *local.shape[1:]), dtype=local.dtype, device=local.device)
                dist.all_gather_into_tensor(buf, pad, group=group)
                return torch.cat([buf[r * mx:r * mx + s] for r, s in enumerate(sizes)])

            scores = gather_padded(local_scores)
            actions = gather_padded(local_actions)
    gmean, idx, score = select_fn(scores, R, K)
    return scores, actions, gmean, idx, score


class ShardedCoverStep:
    """CoverStep over torch.distributed (one process per GPU, NCCL): each rank samples and scores its own
    rephrase slice; a single small all-gather collects the scores for the global argmax."""

    def __init__(self, engine: Engine, samples_per_rephrase: int, group=None, peer_memory: bool = True, **kw):
        """peer_memory=True (default on CUDA): the exchange is cvb_allgather_select - one kernel per rank storing its slice
        into its peers' HBM over NVLink and selecting (cover_vla_b200/comm.py); False: two NCCL all-gathers + cvb_select."""
        import torch.distributed as dist
        self.step = CoverStep(engine, samples_per_rephrase, **kw)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.peer = None
        if peer_memory and self.world > 1:
            from .comm import PeerGather
            cfg = engine.cfg
            n_max = cfg.max_rephrases * cfg.max_samples  # this rank's workspace bound is also its slice bound
            self.peer = PeerGather(n_max * (1 + cfg.chunk_size * 7), device=engine.device, group=group)

    def __call__(self, x: CoverInputs):
        R, K = x.lang_tokens.shape[0], self.step.K
        mine = shard_inputs(x, K, self.world, self.rank)
        if mine.lang_tokens.shape[0] == 0:
            raise ValueError("more ranks than rephrases: shrink the process group for this decision")
        actions, traj, scores, *_ = self.step.sample_and_score(mine)
        local_actions = actions[:, :, :7].contiguous()
        if self.peer is not None:
            s, a, gmean, idx, score = self.peer(scores, local_actions, R, K)
            return s, a, gmean, idx, score
        return gather_and_select(scores, local_actions, R, K, self.step.engine.select, self.group)
