"""Decision latency with and without the per-task prompt cache (cvb_verifier_hold_text), full size, bench configuration."""
import sys

import torch

sys.path.insert(0, ".")
from cover_vla_b200 import synthetic as S  # noqa: E402
from cover_vla_b200.cover import CoverInputs, CoverStep  # noqa: E402

R, K = 8, 5
d, v = S.FULL, S.VFULL
eng = S.build_engine(d, S.make_pi0_weights(d, 0), v, S.make_verifier_weights(v, 0), R, K)
inp = S.make_inputs(d, R, K, seed=3)
vin = S.make_verifier_inputs(v, 1, seed=3)
x = CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(), vf_tokens=vin["tokens"][0].cuda(),
                past=None, lang_len_max=24)
step = CoverStep(eng, K)


def ev(fn, iters=20, warm=5):
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


for hold in (False, True, False, True):
    step.hold_text = hold
    print(f"hold_text={hold}: {ev(lambda: step.sample_and_score(x)):.3f} ms per decision", flush=True)
