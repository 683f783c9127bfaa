"""Full-size batched decisions (configs[4] per-GPU share): B observations x (R x K) candidates per cvb_cover_step_batch."""
import os
import sys

import torch

sys.path.insert(0, ".")
from cover_vla_b200 import synthetic as S  # noqa: E402
from cover_vla_b200.cover import BatchedCoverStep, CoverInputs, CoverStep  # noqa: E402

B, R, K = int(os.environ.get("B", 8)), 8, 5
d, v = S.FULL, S.VFULL
eng = S.build_engine(d, S.make_pi0_weights(d, 0), v, S.make_verifier_weights(v, 0), R, K, max_observations=B)
xs = []
for b in range(B):
    inp = S.make_inputs(d, R, K, seed=100 + b)
    vin = S.make_verifier_inputs(v, 1, seed=100 + b)
    xs.append(CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                          lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                          noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(),
                          vf_tokens=vin["tokens"][0].cuda(), past=None, lang_len_max=24))


def ev(fn, iters=5, warm=3):
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


for nb in sorted({1, 2, 4, B}):
    if nb > B:
        continue
    xb = BatchedCoverStep.stack(xs[:nb])
    step = BatchedCoverStep(eng, K)
    ms = ev(lambda: step.sample_and_score(xb))
    print(f"B={nb}: {ms:.2f} ms per step = {ms / nb:.2f} ms per decision, {nb * R * K / ms * 1e3:.0f} candidates/s", flush=True)
    st = [xb.image, xb.lang_tokens, xb.lang_len, xb.state, xb.noise]
    eng.pi0_sample_batch(*st, K=K, lang_len_max=24)
    for ph, nm in enumerate(["vision", "prefix", "denoise"]):
        print(f"    {nm}: {ev(lambda: eng.pi0_run_phase(ph, R, K, nb), 3, 1):.3f} ms (eager)", flush=True)
single = CoverStep(eng, K)
print(f"single decision on the batch handle: {ev(lambda: single.sample_and_score(xs[0])):.2f} ms")
