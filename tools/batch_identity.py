"""FULL-size check of the batching invariant: B decisions in one cvb_cover_step_batch == B single decisions on the same
(batch-capable) handle, bit for bit - sampler, trajectories, scores, index (tests/test_batch_gpu.py does this at MID size)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from cover_vla_b200 import synthetic as S  # noqa: E402
from cover_vla_b200.cover import BatchedCoverStep, CoverInputs  # noqa: E402

B, R, K = int(os.environ.get("B", 4)), 8, 5
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
step = BatchedCoverStep(eng, K)
singles = [[t.clone() for t in step.sample_and_score(BatchedCoverStep.stack([x]))] for x in xs]
out = [t.clone() for t in step.sample_and_score(BatchedCoverStep.stack(xs))]
torch.cuda.synchronize()
ok = True
for i, nm in enumerate(["actions", "traj", "scores", "group_mean", "best_idx", "best_score"]):
    for b in range(B):
        eq = torch.equal(out[i][b].reshape(-1), singles[b][i].reshape(-1))
        ok &= eq
        if not eq:
            diff = (out[i][b].reshape(-1).float() - singles[b][i].reshape(-1).float()).abs().max().item()
            print(f"observation {b} {nm}: NOT equal, max|diff| {diff:.3e}")
print(f"FULL size, B={B}: batched == singles bit for bit: {ok}")
