"""Full-size CoverStep timing (graph + overlap) with phase breakdown."""
import sys, time
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import synthetic as S
from cover_vla_b200.cover import CoverInputs, CoverStep
R, K = 8, 5
d, v = S.FULL, S.VFULL
import os
eng = S.build_engine(d, S.make_pi0_weights(d, 0), v, S.make_verifier_weights(v, 0), R, K, use_cuda_graph=int(os.environ.get("CVB_GRAPH", "1")))
inp = S.make_inputs(d, R, K, seed=3); vin = S.make_verifier_inputs(v, 1, seed=3)
x = CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(), vf_tokens=vin["tokens"][0].cuda(), past=None)
step = CoverStep(eng, K)
def ev(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for ov in (True, False):
    step.overlap_context = ov
    print(f"CoverStep overlap={ov}: {ev(lambda: step.sample_and_score(x)):.3f} ms")
print(f"pi0_sample: {ev(lambda: eng.pi0_sample(x.image, x.lang_tokens, x.lang_len, x.state, x.noise, K=K)):.3f} ms")
for ph, nm in enumerate(["vision", "prefix", "denoise"]):
    print(f"  {nm}: {ev(lambda: eng.pi0_run_phase(ph, R, K), 5, 1):.3f} ms (eager)")
traj = step.sample_and_score(x)[1]
print(f"verifier context: {ev(lambda: eng.verifier_context(x.vf_image, x.vf_tokens)):.3f} ms")
print(f"verifier traj+select: {ev(lambda: eng.verifier_score(None, None, traj, R, K, recompute_context=False)):.3f} ms")
