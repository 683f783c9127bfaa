"""Full-size pi0 (+ verifier) on the GPU: golden check (R=2,K=2) and phase timing at R=8,K=5."""
import sys, time
import torch
sys.path.insert(0, ".")
from oracle import pi0_oracle as O, verifier_oracle as V
from tests.helpers import build_full_engine


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


t = time.time()
d, v = O.FULL, V.VFULL
w = O.make_pi0_weights(d, 0)
vw = V.make_verifier_weights(v, 0)
print("weights generated", time.time() - t, flush=True)
t = time.time()
eng = build_full_engine(d, w, v, vw, 8, 5)
print("engine built", time.time() - t, flush=True)
del w
gold = torch.load("tests/golden/pi0_full_R2K2.pt")
inp = O.make_inputs(d, 2, 2, seed=0)
args = (inp["image"][0].cuda().contiguous(), inp["tokens"].cuda(), inp["lens"].to(torch.int32).cuda(),
        inp["state"][0].cuda().contiguous(), inp["noise"].cuda())
out = eng.pi0_sample(*args, K=2).cpu()
print("FULL R2K2 vs reference golden: actions max-abs", (out - gold["actions"]).abs().max().item(),
      "v0 rel", ((eng.debug("v0", (4, 4, 32), torch.float32).cpu() - gold["v0"]).norm() / gold["v0"].norm()).item(), flush=True)
R, K = 8, 5
inp = O.make_inputs(d, R, K, seed=3)
args = (inp["image"][0].cuda().contiguous(), inp["tokens"].cuda(), inp["lens"].to(torch.int32).cuda(),
        inp["state"][0].cuda().contiguous(), inp["noise"].cuda())
for _ in range(3):
    eng.pi0_sample(*args, K=K)
torch.cuda.synchronize()
print("pi0_sample R8K5 (graph) ms:", ev_time(lambda: eng.pi0_sample(*args, K=K)))
for ph, nm in enumerate(["vision", "prefix", "denoise"]):
    print(f"  phase {nm}: {ev_time(lambda: eng.pi0_run_phase(ph, R, K)):.3f} ms (eager)")
vin = V.make_inputs(v, R * K, seed=3)
traj = V.pad_histories(vin["histories"], v.history).cuda()
img, tok = vin["image"][0].cuda().contiguous(), vin["tokens"][0].cuda()
print("verifier full (context+traj) ms:", ev_time(lambda: eng.verifier_score(img, tok, traj, R, K)))
print("verifier traj only ms:", ev_time(lambda: eng.verifier_score(None, None, traj, R, K, recompute_context=False)))
