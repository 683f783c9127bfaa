"""One eager full-size decision between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import synthetic as S
from cover_vla_b200.cover import CoverInputs, CoverStep

R, K = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 5
d, v = S.FULL, S.VFULL
eng = S.build_engine(d, S.make_pi0_weights(d, 0), v, S.make_verifier_weights(v, 0), R, K, use_cuda_graph=0)
inp = S.make_inputs(d, R, K, seed=3)
vin = S.make_verifier_inputs(v, 1, seed=3)
x = CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(),
                vf_tokens=vin["tokens"][0].cuda(), past=None, lang_len_max=int(inp["lens"].max()))
step = CoverStep(eng, K)
for _ in range(2):
    step.sample_and_score(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.sample_and_score(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
