"""Where the verifier context lands under the denoise loop: event timestamps of the unfused CoverStep (pi0 graph on the main
stream, context graph on a side stream) with and without the side work."""
import sys

import torch

sys.path.insert(0, ".")
from cover_vla_b200 import synthetic as S  # noqa: E402
from cover_vla_b200.cover import CoverInputs  # noqa: E402

R, K = 8, 5
d, v = S.FULL, S.VFULL
eng = S.build_engine(d, S.make_pi0_weights(d, 0), v, S.make_verifier_weights(v, 0), R, K)
inp = S.make_inputs(d, R, K, seed=3)
vin = S.make_verifier_inputs(v, 1, seed=3)
x = CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(), vf_tokens=vin["tokens"][0].cuda(),
                past=None, lang_len_max=24)
side = torch.cuda.Stream()
cur = torch.cuda.current_stream()


def run(with_ctx, delay_phase):
    """delay_phase: 0 = context starts with the sampler, 1 = after vision + prefix (run as separate phase calls)."""
    E = lambda: torch.cuda.Event(enable_timing=True)
    t0, t_pre, t_pi0, c0, c1 = E(), E(), E(), E(), E()
    t0.record(cur)
    if delay_phase == 0 and with_ctx:
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            c0.record(side)
            eng.verifier_context(x.vf_image, x.vf_tokens)
            c1.record(side)
    eng.pi0_run_phase(0, R, K)
    eng.pi0_run_phase(1, R, K)
    t_pre.record(cur)
    if delay_phase == 1 and with_ctx:
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            c0.record(side)
            eng.verifier_context(x.vf_image, x.vf_tokens)
            c1.record(side)
    eng.pi0_run_phase(2, R, K)
    t_pi0.record(cur)
    cur.wait_stream(side)
    torch.cuda.synchronize()
    r = {"prefix_done": t0.elapsed_time(t_pre), "denoise_done": t0.elapsed_time(t_pi0)}
    if with_ctx:
        r.update(ctx_start=t0.elapsed_time(c0), ctx_done=t0.elapsed_time(c1))
    return r


eng.pi0_sample(x.image, x.lang_tokens, x.lang_len, x.state, x.noise, K=K, lang_len_max=24)
for with_ctx, ph in ((False, 1), (True, 1), (True, 0), (False, 1), (True, 1)):
    for _ in range(3):
        r = run(with_ctx, ph)
    print(f"ctx={with_ctx} start_after_prefix={ph}: " + "  ".join(f"{k}={v:.2f}" for k, v in r.items()), flush=True)
