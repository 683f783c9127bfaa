"""One eager batched decision (B observations x R x K candidates) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum`.  B from the environment (default 8)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from cover_vla_b200 import synthetic as S  # noqa: E402
from cover_vla_b200.cover import BatchedCoverStep, CoverInputs  # noqa: E402

B, R, K = int(os.environ.get("B", 8)), 8, 5
os.environ.setdefault("CVB_GRAPH", "0")
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
xb = BatchedCoverStep.stack(xs)
step = BatchedCoverStep(eng, K)
step.sample_and_score(xb)   # eager warm-up (the graph cache captures on the second call; only one more call follows)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.sample_and_score(xb)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
