"""Phase timeline of one skinny GEMM launch (globaltimer stamps per CTA)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops, _lib
M, N, K, split = (int(x) for x in sys.argv[1:5])
lib = _lib.load()
lib.cvb_debug_set_timestamps.argtypes = [C.c_void_p]
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16) for _ in range(12)]
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for w in ws[:4]:
    ops.gemm_bf16(a, w, out=out, force_bn=split)
torch.cuda.synchronize()
ts = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
lib.cvb_debug_set_timestamps(C.c_void_p(ts.data_ptr()))
for rep in range(3):
    ts.zero_()
    ops.gemm_bf16(a, ws[6 + rep], out=out, force_bn=split)
    torch.cuda.synchronize()
    t = ts.view(-1, 8).cpu()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t[:, :7] - t0).float() / 1e3
    names = ["start", "setup", "mma_done", "bar1", "sent", "bar2", "final"]
    print(f"M={M} N={N} K={K} split={split}: {t.shape[0]} CTAs; us since first CTA start (min / median / max)")
    for i, n in enumerate(names):
        c = rel[:, i]
        print(f"  {n:9s} {c.min():7.2f} {c.median():7.2f} {c.max():7.2f}")
lib.cvb_debug_set_timestamps(C.c_void_p(0))
