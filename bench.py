#!/usr/bin/env python
"""bench.py - CoVer-VLA sample-and-verify on B200 (contract: see the task statement / DESIGN.md section 5).

One "step" = one CoVer decision: pi0 samples N = R*K = 8*5 = 40 action chunks for one observation
(SigLIP tower once, PaliGemma prefix once per unique rephrase, 10 denoise steps over all candidates), the
candidates are formatted for the verifier on the device, scored by the 3-member bridge_verifier ensemble
(SigLIP2 ViT-L trunk + text tower + fp32 heads) and the group-mean / argmax rule picks the winner
(BASELINE.json configs[2]: "full CoVer step ... simpler_widowx shapes, 1 B200").  Full-size models,
random-init weights, synthetic inputs.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (weak scaling:
        every rank runs its own observation - episode-parallel, BASELINE.json configs[4] - no data-path collective)
    python bench.py --impl reference ...      (the reference algorithm on the host CPU: oracle port, rank 0 only)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

R_DEFAULT, K_DEFAULT = 8, 5
METRIC = "verified_candidates_per_sec"
UNIT = "candidates/s"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Start of the timed region: only rows read after this call are reported (the sampler is started before the
        warm-up so that nvidia-smi's start-up time does not eat a short timed region)."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        first = getattr(self, "first", 0)
        rows, window = self.rows[first:], "timed region"
        if not rows and self.rows:  # timed region shorter than one sampling period: the last warm-up samples (same load)
            rows, window = self.rows[-3:], "warm-up steps right before the timed region (it was shorter than one sampling period)"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ----------------------------------------------------------------------------------------------------
def algorithmic_flops(R: int, K: int, prefix_rows: int = 328) -> dict:
    """De-duplicated algorithmic FLOPs per decision, counted on the rows the engine PROCESSES:
    the prefix runs on `prefix_rows` = 256 image tokens + the valid language rows (right-padding is skipped, SURVEY.md
    F11), not on all 328.  Linear layers scale with the rows, the attention (4 Tq Tk d per layer) with their square."""
    N = R * K
    att = lambda t: 4.0 * t * t * 2048 * 18
    prefix = (1315.9e9 - att(328)) * prefix_rows / 328.0 + att(prefix_rows)
    return {"vision": 220.2e9, "prefix": prefix * R, "denoise": 33.6e9 * N, "verifier_trunk": 420.6e9,
            "verifier_heads": 3.7e9 + 20.2e9 * N / 40}


EXPERT_WEIGHT_BYTES = 311.4e6 * 2  # bf16 action-expert weights streamed once per denoise step


def ncu_traffic(name: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from THIS round's `ncu --set full` capture as
    summarised in profiles/r2_ncu_traffic.json (written by tools/ncu_traffic.py from the .ncu-rep); None if absent."""
    p = ROOT / "profiles" / "r2_ncu_traffic.json"
    if not p.exists():
        return None, None
    d = json.loads(p.read_text())
    e = d.get(name)
    return (e.get("dram_bytes_per_launch"), "profiles/r2_ncu_traffic.json:" + name) if e else (None, None)


def make_device_inputs(S, d, v, R, K, seed, device):
    import torch
    from cover_vla_b200.cover import CoverInputs
    inp = S.make_inputs(d, R, K, seed=seed)
    vin = S.make_verifier_inputs(v, 1, seed=seed)
    past = torch.tensor([[0.004, -0.011, 0.002, 0.01, -0.02, 0.03, 1.0],
                         [0.001, 0.006, -0.003, 0.0, 0.01, -0.04, 1.0],
                         [-0.002, 0.003, 0.007, 0.02, 0.0, 0.05, 0.0]])
    host = dict(image=inp["image"][0].contiguous(), lang_tokens=inp["tokens"].contiguous(),
                lang_len=inp["lens"].to(torch.int32).contiguous(), state=inp["state"][0].contiguous(),
                noise=inp["noise"].contiguous(), vf_image=vin["image"][0].contiguous(),
                vf_tokens=vin["tokens"][0].contiguous(), past=past)
    host = {k: t.pin_memory() for k, t in host.items()}
    # the host tokenised the prompts, so it knows the longest one without a device round trip
    lmax = int(inp["lens"].max())
    dev = CoverInputs(**{k: t.to(device) for k, t in host.items()}, lang_len_max=lmax)
    return host, dev, lmax


def episode_driver_leg(engB, xsB, R, K, Bo, n_ticks, world, device):
    """BASELINE.json configs[4] through the production surface: Bo environments stepped by EpisodeBatchDriver, timed per
    DECISION tick with the host clock (the D2H read of the results synchronises).  Frames are synthetic 480 x 640 uint8
    (the simulator's camera size) from a small host pool; states are random; the environments never finish."""
    import numpy as np
    import torch
    from cover_vla_b200.episodes import EpisodeBatchDriver, TaskPrompts

    class Env:
        def __init__(self, seed):
            rng = np.random.default_rng(seed)
            self.pool = [rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8) for _ in range(4)]
            self.states = rng.normal(size=(16, 7)).astype(np.float32)
            self.k = 0

        def reset(self, task, seed):
            self.k = 0
            return 0

        def step(self, action):
            self.k += 1
            return self.k, False

        def frame(self, obs):
            return self.pool[obs % 4]

        def state(self, obs):
            return self.states[obs % 16]

    tasks = [TaskPrompts([f"instruction {i}" for i in range(R)], x.lang_tokens, x.lang_len,
                         x.vf_tokens[None, :].repeat(R, 1).contiguous(), x.lang_len_max) for x in xsB]
    drv = EpisodeBatchDriver(engB, [Env(7 + b) for b in range(Bo)], tasks, [(b, 0, b) for b in range(Bo)], R, K,
                             max_steps=10 ** 9)
    n = drv.n_action_steps
    for _ in range(6 * n):  # the history grows 0 -> 4 -> 6 rows: three shapes to run eagerly, capture and replay
        drv.tick()
    torch.cuda.synchronize()
    t_dec = []
    for _ in range(n_ticks * n):
        deciding = drv.slots[0].t % n == 0
        t1 = time.perf_counter()
        drv.tick()
        if deciding:
            t_dec.append((time.perf_counter() - t1) * 1e3)
    ms = statistics.median(t_dec)  # this rank's; the caller takes the max over ranks (no collective in here)
    return {"what": "EpisodeBatchDriver.tick() at a decision tick, host clock, p50: %d host frames (480x640x3 uint8) -> H2D -> "
                    "Lanczos4 224 / bilinear-antialias 256 + bicubic 384 on the device -> cvb_cover_step_batch -> gate, "
                    "execution-format actions + gripper vote, history rows -> one D2H read; per-task prompt cache active "
                    "(the verifier text tower is skipped on ticks where no environment switched instruction)" % Bo,
            "decision_ticks": len(t_dec), "ms_per_decision_tick": round(ms, 3), "ms_per_decision": round(ms / Bo, 3),
            "value": round(world * Bo * R * K / (ms * 1e-3), 2), "unit": UNIT,
            "h2d_bytes_per_decision": 480 * 640 * 3 + 32 * 4 + 6 * 7 * 4,  # frame, state, history tail (noise is drawn on the device)
            "d2h_bytes_per_decision": (2 + 2 * 7 * n) * 8,
            "batched_calls": drv.batched_calls, "decisions": drv.decisions}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from cover_vla_b200 import _lib, synthetic as S
    from cover_vla_b200 import build as cvb_build
    from cover_vla_b200.cover import CoverInputs, CoverStep, ShardedCoverStep, rephrase_shard
    from cover_vla_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if not _lib.LIB_PATH.exists():
        cvb_build.build()
    lib = _lib.load()
    lib.cvb_launch_count.restype = __import__("ctypes").c_int64

    R, K = args.rephrases, args.samples
    N = R * K
    sharded = args.mode == "sharded"
    if sharded and world > R:
        raise SystemExit("--mode sharded needs at least one rephrase per rank")
    d, v = S.FULL, S.VFULL
    t0 = time.time()
    w = S.make_pi0_weights(d, seed=0)
    vw = S.make_verifier_weights(v, seed=0)
    if sharded:
        a_, b_ = rephrase_shard(R, world, rank)
        R_loc = b_ - a_
        R_max = max(rephrase_shard(R, world, r)[1] - rephrase_shard(R, world, r)[0] for r in range(world))
    else:
        R_loc = R_max = R
    eng = S.build_engine(d, w, v, vw, R_max, K, device=device,
                         use_cuda_graph=0 if os.environ.get("CVB_BENCH_EAGER") else 1)  # CVB_BENCH_EAGER=1: diagnostic only
    t_build = time.time() - t0
    # episode mode: every rank has its own observation; sharded mode: all ranks see the SAME observation
    host, x, lmax = make_device_inputs(S, d, v, R, K, seed=100 + (0 if sharded else rank), device=device)
    step = CoverStep(eng, K)
    sstep = ShardedCoverStep(eng, K) if sharded and world > 1 else None

    def one_step():
        if sstep is not None:
            return sstep(x)
        return step.sample_and_score(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: the first call of a shape runs every kernel eagerly (this is where the launches of one decision are
    # counted), the second is captured into the CUDA graph, later ones replay it
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = lib.cvb_launch_count()
    out = one_step()
    torch.cuda.synchronize()
    launches_per_step = int(lib.cvb_launch_count() - l0)
    for _ in range(max(3, args.warmup) - 1):
        out = one_step()
    torch.cuda.synchronize()
    traj0 = out[1] if sstep is None else step.sample_and_score(sstep_inputs(x, K, world, rank))[1]

    # ---- timed region: `steps` decisions, device-resident inputs, CUDA events, max over ranks
    barrier()
    clocks.mark()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        out = one_step()
        ev[i + 1].record()
    barrier()
    clk = clocks.stop()
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        t = torch.tensor([total_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    cands_per_step = N if sharded else N * world
    value = cands_per_step / (ms_per_step / 1e3)

    # ---- end to end through the public API with HOST buffers (H2D + D2H inside the timed region)
    def e2e_step():
        xin = CoverInputs(**{k: t.to(device, non_blocking=True) for k, t in host.items()}, lang_len_max=lmax)
        if sstep is not None:
            scores, actions, gmean, idx, score = sstep(xin)
            winner = actions.index_select(0, idx.to(torch.int64).reshape(1))[0]
            packed = torch.cat([idx.to(torch.float32).reshape(1), score.reshape(1), winner.reshape(-1)]).cpu()
            return int(packed[0]), float(packed[1]), packed[2:].reshape(-1, 7)
        return step(xin)  # returns python (idx, score, winner actions): includes the D2H read

    for _ in range(3):
        e2e_step()
    barrier()
    t_e2e = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        idx, score, winner = e2e_step()
        t_e2e.append((time.perf_counter() - t1) * 1e3)
    barrier()
    e2e_ms = sum(t_e2e) / len(t_e2e)
    e2e_p50 = statistics.median(t_e2e)
    if world > 1:
        t = torch.tensor([e2e_ms, e2e_p50], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_p50 = float(t[0].item()), float(t[1].item())
    h2d = sum(t.numel() * t.element_size() for t in host.values())
    d2h = (2 + d.chunk_size * 7) * 4

    # ---- throughput mode (BASELINE.json configs[4], per-GPU share): B independent configs[2] decisions per step through
    # ONE cvb_cover_step_batch on a batch-capable handle - every weight is streamed once for all B * N candidates
    batched = None
    if not sharded and args.batch_obs > 1:
        from cover_vla_b200.cover import BatchedCoverStep
        Bo = args.batch_obs
        engB = S.build_engine(d, w, v, vw, R, K, device=device, max_observations=Bo)
        xsB = [make_device_inputs(S, d, v, R, K, seed=1000 + rank * Bo + b, device=device)[1] for b in range(Bo)]
        xb = BatchedCoverStep.stack(xsB)
        bstep = BatchedCoverStep(engB, K)
        for _ in range(3):
            bstep.sample_and_score(xb)
        nb = max(3, args.steps // 4)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nb):
            outB = bstep.sample_and_score(xb)
        e1.record()
        barrier()
        msB = e0.elapsed_time(e1) / nb
        if world > 1:
            t = torch.tensor([msB], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            msB = float(t.item())
        batched = {"workload": "BASELINE.json configs[4] per-GPU share: %d independent configs[2] decisions per step in one "
                               "cvb_cover_step_batch (sampler batched, one verifier context per observation)" % Bo,
                   "observations_per_gpu_per_step": Bo, "steps": nb, "ms_per_step": round(msB, 3),
                   "ms_per_decision": round(msB / Bo, 3), "value": round(world * Bo * N / (msB * 1e-3), 2), "unit": UNIT,
                   "vs_single_decision_mode": round((world * Bo * N / (msB * 1e-3)) / value, 3)}
        # ... and the same handle driven by the episode-batched driver (cover_vla_b200/episodes.py) from HOST simulator frames:
        # per decision tick Bo uint8 480 x 640 frames go H2D, both image chains, the batched decision, the gate, the
        # execution-format actions + gripper vote and the history rows run on the device, one D2H read returns them.
        # A sub-measurement: a failure here must not lose the headline.
        try:
            leg = episode_driver_leg(engB, xsB, R, K, Bo, nb, world, device)
        except Exception as e:  # noqa: BLE001
            leg = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world > 1:  # max over ranks OUTSIDE the guarded region: every rank reaches this collective whatever happened above
            t = torch.tensor([leg.get("ms_per_decision_tick", -1.0), 1.0 if "error" in leg else 0.0], device=device,
                             dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if "error" not in leg:
                if float(t[1].item()) > 0:
                    leg = {"error": "the leg failed on another rank"}
                else:
                    ms_all = float(t[0].item())
                    leg.update(ms_per_decision_tick=round(ms_all, 3), ms_per_decision=round(ms_all / Bo, 3),
                               value=round(world * Bo * N / (ms_all * 1e-3), 2))
        batched["episode_driver"] = leg
        engB.close()
        del engB, xsB, xb, outB
        torch.cuda.empty_cache()

    # ---- BASELINE.json configs[3] under the same clock (N > 1 only): ONE observation, 16 rephrases x 16 samples, sharded
    # by rephrase over the ranks; the exchange is cvb_allgather_select (peer-memory stores over NVLink + in-kernel
    # selection).  Every rank also runs the WHOLE decision alone and checks the sharded result against it.
    sharded_info = None
    if not sharded and world > 1 and world <= 16 and not args.no_sharded:
        Rs, Ks = 16, 16
        engS = S.build_engine(d, w, v, vw, Rs, Ks, device=device)
        _, xS, _ = make_device_inputs(S, d, v, Rs, Ks, seed=777, device=device)  # the SAME observation on every rank
        a1, t1, s1, g1, i1, b1 = [t.clone() for t in CoverStep(engS, Ks).sample_and_score(xS)]
        res = {}
        for name, peer in (("peer_memory", True), ("nccl", False)):
            ss = ShardedCoverStep(engS, Ks, peer_memory=peer)
            for _ in range(3):
                so = ss(xS)
            barrier()
            ns = max(3, args.steps // 2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(ns):
                so = ss(xS)
            e1.record()
            barrier()
            msS = e0.elapsed_time(e1) / ns
            scores_s, actions_s, gmean_s, idx_s, score_s = so
            ok_idx = int(idx_s.item()) == int(i1.item())
            bit = bool(torch.equal(scores_s, s1)) and bool(torch.equal(actions_s, a1[:, :, :7]))
            diff = float((scores_s - s1).abs().max().item())
            t = torch.tensor([msS, -float(ok_idx), -float(bit), diff], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max time / worst agreement over the ranks
            res[name] = {"ms_per_decision": round(float(t[0]), 3), "value": round(Rs * Ks / (float(t[0]) * 1e-3), 2),
                         "index_equals_single_gpu_on_every_rank": bool(t[1] == -1.0),
                         "scores_bit_identical_to_single_gpu": bool(t[2] == -1.0),
                         "scores_max_abs_diff_vs_single_gpu": float(t[3])}
            if ss.peer is not None:
                ss.peer.close()
        # (reported, not raised: a sub-measurement must not lose the line.  The winner can only differ from the 1-GPU pass when
        # the top-2 gap is inside the bf16 noise between the two kernel paths named in `note`; the parity tests gate that)
        if not res["peer_memory"]["index_equals_single_gpu_on_every_rank"]:
            res["peer_memory"]["error"] = "the sharded decision picked another candidate than the 1-GPU decision on some rank"
        sharded_info = {"workload": "BASELINE.json configs[3]: ONE observation, 16 rephrases x 16 samples = 256 candidates sharded "
                                    "by rephrase over %d ranks, fused peer-memory all-gather + select" % world,
                        "unit": UNIT, "candidates": Rs * Ks, **res["peer_memory"], "nccl_allgather_variant": res["nccl"],
                        "note": "bit-identical when the per-rank row count takes the same kernel path as the 1-GPU pass "
                                "(> 256 suffix rows: fused-epilogue GEMMs; <= 256: split-K), else equal within bf16 noise"}
        engS.close()
        del engS
        torch.cuda.empty_cache()
    del w, vw

    line = None
    if rank == 0:
        peaks, peak_src = _peaks()

        # ---- phase breakdown (CUDA events) and the roofline of the dominant kernel
        def ev_ms(fn, iters=5, warm=3):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / iters
        phases = {nm: ev_ms(lambda p=p: eng.pi0_run_phase(p, R_loc, K)) for p, nm in enumerate(["vision", "prefix", "denoise"])}
        phases["verifier_context"] = ev_ms(lambda: eng.verifier_context(x.vf_image, x.vf_tokens))
        phases["verifier_trajectories_select"] = ev_ms(lambda: eng.verifier_score(None, None, traj0, R_loc, K, recompute_context=False))
        lang_rows = min(d.max_lang_len, (lmax + 7) // 8 * 8)
        fl = algorithmic_flops(R, K, d.n_img_tokens + lang_rows)
        total_flops = sum(fl.values())

        # dominant kernel: the prefix gate/up GeGLU GEMM (tcgen05), M = R*(256 + language rows), N = 2*16384 packed,
        # K = 2048.  Timed alone, cycling through 6 different weight matrices (6 x 134 MB >> 126 MB L2).
        M_, Kd, I_ = R_loc * (d.n_img_tokens + lang_rows), d.lm_width, d.lm_mlp
        a_ = torch.randn(M_, Kd, device=device, dtype=torch.bfloat16)
        ws = [(torch.randn(2 * I_, Kd, device=device) * 0.02).to(torch.bfloat16) for _ in range(6)]
        o_ = torch.empty(M_, I_, device=device, dtype=torch.bfloat16)
        for wgt in ws:
            ops.gemm_bf16(a_, wgt, epilogue=ops.EPI_GEGLU, n_out=I_, out=o_)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        reps = 5
        a.record()
        for _ in range(reps):
            for wgt in ws:
                ops.gemm_bf16(a_, wgt, epilogue=ops.EPI_GEGLU, n_out=I_, out=o_)
        b.record()
        torch.cuda.synchronize()
        gemm_ms = a.elapsed_time(b) / (reps * len(ws))
        gemm_flops = 2.0 * M_ * (2 * I_) * Kd
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
        peak = float(peaks["bf16_tflops"])
        roofline = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_2sm<EPI_GEGLU> (cta_group::2, 256x256 tiles; prefix gate/up, M=%d N=%d K=%d)" % (M_, 2 * I_, Kd),
                    "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                    "traffic": ncu_traffic("prefix_gateup_gemm")[0], "traffic_source": ncu_traffic("prefix_gateup_gemm")[1],
                    "algorithmic_bytes": int((M_ * Kd + 2 * I_ * Kd + M_ * I_) * 2),
                    "peak_source": peak_src + ", burst figure (kernel timed alone)",
                    "launch_ms": round(gemm_ms, 4), "launches_per_step": d.layers - 1,
                    "step_frac_of_sustained_peak": round(total_flops / (ms_per_step * 1e-3) / 1e12 / float(peaks["bf16_tflops_sustained"]) /
                                                         (world if sharded else 1), 4)}
        del a_, ws, o_
        # second roofline: the denoise loop is an HBM weight stream (the expert's 0.62 GB once per Euler step, M = 5 N rows)
        n_steps = d.num_steps
        den_gbs = EXPERT_WEIGHT_BYTES * n_steps / (phases["denoise"] * 1e-3) / 1e9
        roofline_denoise = {"bound": "hbm", "kernel": "denoise loop (%d Euler steps x %d expert layers; split-K tcgen05 GEMMs + "
                            "tcgen05 decode attention + RMSNorm-reduce)" % (n_steps, d.layers),
                            "achieved": round(den_gbs, 1), "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                            "frac": round(den_gbs / float(peaks["hbm_gbs"]), 4), "traffic": None,
                            "algorithmic_bytes": int(EXPERT_WEIGHT_BYTES * n_steps), "phase_ms": round(phases["denoise"], 3),
                            "peak_source": peak_src + ", STREAM-style copy"}

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_sample()
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("BASELINE.json configs[3]: ONE observation, %d rephrases x %d samples (%d candidates) sharded by "
                                    "rephrase over the ranks, NCCL all-gather of scores/actions, full-size random-init models" % (R, K, N))
                       if sharded else
                       ("BASELINE.json configs[2]: full CoVer step, pi0 %d rephrases x %d samples (%d candidates) "
                        "+ 3-member verifier argmax, simpler_widowx shapes, full-size random-init models" % (R, K, N)),
                       "rephrases": R, "samples_per_rephrase": K, "candidates_per_step": cands_per_step,
                       "parallelism": ("rephrase-sharded over %d ranks + score all-gather" % world) if sharded else
                       ("episode-parallel: 1 observation per GPU per step, no data-path collective" if world > 1 else "1 GPU"),
                       "l2": "no flush needed: every step streams ~8.6 GB of weights from HBM (>> 126 MB L2)",
                       "max_valid_language_tokens": lmax,
                       "cuda_graph": bool(eng.cfg.use_cuda_graph)},
            "p50_ms": round(statistics.median(per_step), 3),
            "p95_ms": round(sorted(per_step)[min(len(per_step) - 1, int(0.95 * len(per_step)))], 3),
            "phases_ms": {k: round(val, 3) for k, val in phases.items()},
            "algorithmic_tflop_per_step": round(total_flops / 1e12, 3),
            "clocks": clk,
            "e2e": {"value": round(cands_per_step / (e2e_ms / 1e3), 2), "unit": UNIT, "ms_per_step": round(e2e_ms, 3),
                    "p50_ms": round(e2e_p50, 3), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roofline,
            "roofline_denoise": roofline_denoise,
            "batched": batched,
            "sharded": sharded_info,
            "cpu_baseline": cpu,
            "build_s": round(t_build, 1),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def sstep_inputs(x, K, world, rank):
    from cover_vla_b200.cover import shard_inputs
    return shard_inputs(x, K, world, rank)


# ----------------------------------------------------------------------------------------------------
# CPU legs: the oracle (bit-exact restatement of the reference) on the host cores
# ----------------------------------------------------------------------------------------------------
_CPU_STATE = {}


def _cpu_setup():
    """Full-size weights + oracle modules (test infrastructure; only the CPU legs import oracle/)."""
    if _CPU_STATE:
        return _CPU_STATE
    import torch
    from cover_vla_b200 import synthetic as S
    from oracle import pi0_oracle as O
    from oracle import verifier_oracle as V
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _CPU_STATE.update(S=S, O=O, V=V, cores=cores, w=S.make_pi0_weights(S.FULL, 0), vw=S.make_verifier_weights(S.VFULL, 0))
    return _CPU_STATE


def cpu_decision(R: int, K: int, seed: int):
    """One CoVer decision for R*K candidates exactly as the reference executes it (batch layout N, no
    de-duplication): sample_actions at B = N, verifier-format, ensemble scores, selection."""
    import numpy as np
    import torch
    st = _cpu_setup()
    S, O, V = st["S"], st["O"], st["V"]
    d, v = S.FULL, S.VFULL
    inp = S.make_inputs(d, R, K, seed=seed)
    b = S.expand_to_batch(inp, K)
    vin = S.make_verifier_inputs(v, 1, seed=seed)
    t0 = time.perf_counter()
    actions = O.sample_actions(st["w"], d, b["image"], b["tokens"], b["masks"], b["state"], b["noise"])
    from cover_vla_b200.cover import BRIDGE_ACTION_P01, BRIDGE_ACTION_P99
    a = actions[:, :, :7].numpy().astype(np.float32)
    fut = np.zeros(a.shape, dtype=np.float64)
    fut[:, :, :6] = (a[:, :, :6] + 1) / 2 * (np.array(BRIDGE_ACTION_P99) - np.array(BRIDGE_ACTION_P01)) + np.array(BRIDGE_ACTION_P01)
    fut[:, :, 6] = np.where(a[:, :, 6] < 0.5, 0, 1)
    best, idx, scores, means = V.compute_max_similarity_scores(st["vw"], v, vin["image"], vin["tokens"],
                                                               [fut[n] for n in range(R * K)], K)
    return time.perf_counter() - t0, idx


CPU_SAMPLE_R, CPU_SAMPLE_K = 8, 5  # one whole configs[2] decision: 40 candidates in the reference's batch layout (B = 40)


def cpu_baseline_sample():
    st = _cpu_setup()
    R, K = CPU_SAMPLE_R, CPU_SAMPLE_K
    cpu_decision(1, 1, seed=1)  # warm-up (thread pools, allocator)
    t, _ = cpu_decision(R, K, seed=2)
    return {"value": round(R * K / t, 4), "unit": UNIT, "cores": st["cores"], "kind": "port",
            "sample": "full-size models, ONE configs[2] decision: all %d candidates (R=%d, K=%d) in the reference's batch "
                      "layout (no de-duplication) incl. verifier; oracle port (bit-exact vs the reference files in the "
                      "authoring container); %.1f s of CPU work" % (R * K, R, K, t)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    st = _cpu_setup()
    R, K = CPU_SAMPLE_R, CPU_SAMPLE_K
    # each step = one whole configs[2] decision: R*K = 40 candidates (B = 40 in the reference's batch layout), ~13 s
    for i in range(max(1, min(args.warmup, 1))):
        cpu_decision(R, K, seed=10 + i)
    times = []
    budget = 900.0  # the driver allows 1800 s for this arm
    t_start = time.perf_counter()
    for i in range(args.steps):
        t, _ = cpu_decision(R, K, seed=20 + i)
        times.append(t)
        if time.perf_counter() - t_start > budget:
            break
    ms = 1e3 * sum(times) / len(times)
    value = R * K / (ms / 1e3)
    sample = ("each step = one full-size configs[2] decision over all %d candidates (R=%d, K=%d, reference batch layout, "
              "no de-duplication, incl. verifier trunk + heads + selection) on %d host threads; %d of %d steps run inside "
              "the %.0f s budget" % (R * K, R, K, st["cores"], len(times), args.steps, budget))
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": round(ms, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[2]: full CoVer step, pi0 %d rephrases x %d samples (%d candidates) "
                                   "+ 3-member verifier argmax, simpler_widowx shapes, full-size random-init models, on the "
                                   "host CPU" % (R, K, R * K),
                       "rephrases": R, "samples_per_rephrase": K, "candidates_per_step": R * K},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": st["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rephrases", type=int, default=R_DEFAULT)
    ap.add_argument("--samples", type=int, default=K_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-obs", type=int, default=8, help="observations per step of the throughput-mode sub-measurement "
                                                             "(configs[4] per-GPU share); 0 / 1 disables it")
    ap.add_argument("--no-sharded", action="store_true", help="skip the configs[3] sub-measurement under torchrun")
    ap.add_argument("--mode", default="episode", choices=["episode", "sharded"],
                    help="episode: every rank decides its own observation (weak scaling, BASELINE configs[2]/[4]); "
                         "sharded: ONE observation, rephrases sharded over the ranks + NCCL score all-gather "
                         "(strong scaling, BASELINE configs[3]; use --rephrases 16 --samples 16)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
