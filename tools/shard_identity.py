"""Which tensors of a rephrase-sharded decision are bit-identical to the whole decision?  One GPU: the shards of every
world size are run one after the other on the same handle (what each rank of ShardedCoverStep computes locally)."""
import sys

import torch

sys.path.insert(0, ".")
from cover_vla_b200 import synthetic as S  # noqa: E402
from cover_vla_b200.cover import CoverInputs, CoverStep, shard_inputs  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "MID"
R, K = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3, 2)
d, v = getattr(S, name), getattr(S, "V" + name)
eng = S.build_engine(d, S.make_pi0_weights(d, 0), v, S.make_verifier_weights(v, 0), R, K)
inp = S.make_inputs(d, R, K, seed=21)
vin = S.make_verifier_inputs(v, 1, seed=21)
x = CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(), vf_tokens=vin["tokens"][0].cuda(),
                past=None, lang_len_max=int(inp["lens"].max()))
step = CoverStep(eng, K)
for fused in (True, False):
    step.fused = fused
    whole = [t.clone() for t in step.sample_and_score(x)]
    for world in (2, 3):
        if world > R:
            continue
        parts = [[t.clone() for t in step.sample_and_score(shard_inputs(x, K, world, r))] for r in range(world)]
        for i, nm in enumerate(["actions", "traj", "scores"]):
            got = torch.cat([p[i] for p in parts])
            diff = (got - whole[i]).abs().max().item()
            print(f"{name} R={R} K={K} fused={fused} world={world} {nm}: equal={torch.equal(got, whole[i])} max|diff|={diff:.3e}")
