"""Timeline of the persistent expert kernel: the last 16 launches of a denoise loop (globaltimer stamps per CTA)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
from cover_vla_b200 import _lib, synthetic as S  # noqa: E402

lib = _lib.load()
lib.cvb_debug_set_timestamps.argtypes = [C.c_void_p]
R, K = int(os.environ.get("R", 8)), int(os.environ.get("K", 5))
d = S.FULL
eng = S.build_engine(d, S.make_pi0_weights(d, 0), None, None, R, K, use_cuda_graph=0)
inp = S.make_inputs(d, R, K, seed=5)
args = (inp["image"][0].cuda().contiguous(), inp["tokens"].cuda(), inp["lens"].to(torch.int32).cuda(),
        inp["state"][0].cuda().contiguous(), inp["noise"].cuda())
for _ in range(2):
    eng.pi0_sample(*args, K=K)
torch.cuda.synchronize()
G = 148
ts = torch.zeros(8192 + 16 * G * 64, dtype=torch.int64, device="cuda")
lib.cvb_debug_set_timestamps(C.c_void_p(ts.data_ptr()))
eng.pi0_run_phase(2, R, K)
torch.cuda.synchronize()
lib.cvb_debug_set_timestamps(C.c_void_p(0))
t = ts[8192:].view(16, G, 64).cpu()
att = ts[:8192].view(-1, 8).cpu()
att = att[att[:, 0] > 0]
starts = sorted((int(t[i, :, 0].min()), i) for i in range(16) if int(t[i, :, 0].max()) > 0)
print("attention (last launch): start %d, stored max +%.2f us" % (int(att[:, 0].min()), (att[:, 7].max() - att[:, 0].min()) / 1e3))
prev_end = None
names = {5: ["O", "N1", "GU", "D", "N2"], 6: ["O", "N1", "GU", "D", "N2", "QKV"], 2: ["N0", "QKV"]}
for st, i in starts:
    x = t[i]
    nph = int(((x[:, 2::3] > 0).sum(1) > 0).sum()) if False else int((x[0, 2::3] > 0).sum())
    t0 = int(x[:, 0].min())
    end = int(x[:, 2:2 + 3 * nph:3].max())
    gap = "" if prev_end is None else f" gap since previous mega end {(t0 - prev_end) / 1e3:.2f} us"
    print(f"launch slot {i}: {nph} phases, start spread {(int(x[:, 0].max()) - t0) / 1e3:.2f} us, total {(end - t0) / 1e3:.2f} us{gap}")
    nm = names.get(nph, [str(k) for k in range(nph)])
    for p in range(nph):
        mma = x[:, 1 + 3 * p]
        mma = mma[mma > 0]
        work = x[:, 2 + 3 * p]
        bar = x[:, 3 + 3 * p]
        bar = bar[bar > 0]
        s = f"   {nm[p]:4s} work done min/med/max {(work.min() - t0) / 1e3:6.2f} {(work.median() - t0) / 1e3:6.2f} {(work.max() - t0) / 1e3:6.2f}"
        if mma.numel():
            s += f" | first acc ready min/med/max {(mma.min() - t0) / 1e3:6.2f} {(mma.median() - t0) / 1e3:6.2f} {(mma.max() - t0) / 1e3:6.2f} ({mma.numel()} CTAs)"
        if bar.numel():
            s += f" | barrier passed med {(bar.median() - t0) / 1e3:6.2f} max {(bar.max() - t0) / 1e3:6.2f}"
        print(s)
    prev_end = end
