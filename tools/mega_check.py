"""Persistent expert kernel vs the separate-kernel denoise path: same inputs, both engines, actions / v0 compared and
the denoise phase timed.  usage: python tools/mega_check.py [mid] [full]"""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from cover_vla_b200 import synthetic as S  # noqa: E402


def ev(fn, iters=5, warm=2):
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


def run(name, R, K, modes=("0", "1")):
    d = getattr(S, name)
    w = S.make_pi0_weights(d, 0)
    inp = S.make_inputs(d, R, K, seed=5)
    args = (inp["image"][0].cuda().contiguous(), inp["tokens"].cuda(), inp["lens"].to(torch.int32).cuda(),
            inp["state"][0].cuda().contiguous(), inp["noise"].cuda())
    outs, v0s = {}, {}
    for mode in modes:
        os.environ["CVB_DENOISE_MEGA"] = mode
        t = time.time()
        eng = S.build_engine(d, w, None, None, R, K)
        runs = [eng.pi0_sample(*args, K=K).cpu() for _ in range(4)]
        torch.cuda.synchronize()
        same = all(torch.equal(runs[0], r) for r in runs[1:])
        outs[mode] = runs[0]
        v0s[mode] = eng.debug("v0", (R * K, d.chunk_size, d.max_action_dim), torch.float32).cpu()
        ms = ev(lambda: eng.pi0_run_phase(2, R, K))
        tot = ev(lambda: eng.pi0_sample(*args, K=K))
        print(f"{name} R{R}K{K} mega={mode}: denoise {ms:.3f} ms eager, pi0_sample {tot:.3f} ms (graph), replay identical {same}, "
              f"finite {bool(torch.isfinite(runs[0]).all())}, build {time.time() - t:.1f}s", flush=True)
        eng.close()
    if len(modes) == 2:
        a, b = outs[modes[0]], outs[modes[1]]
        print(f"   actions max|mega - legacy| = {(a - b).abs().max().item():.3e}   v0 rel-L2 = "
              f"{((v0s[modes[0]] - v0s[modes[1]]).norm() / v0s[modes[0]].norm()).item():.3e}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["mid"]
    if "mid" in which:
        run("MID", 2, 3)
        run("MID", 8, 5)
    if "full" in which:
        run("FULL", 8, 5)
        run("FULL", 1, 5)
