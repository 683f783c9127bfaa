"""In-graph (warm, back-to-back) time of the per-Euler-step fp32 kernels around the expert layers."""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops

def graph_time(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

for M in (160, 200, 1280):
    x = torch.randn(M, 32, device="cuda"); wc = torch.randn(1024, 32, device="cuda") * 0.1
    b = torch.randn(1024, device="cuda"); tv = torch.randn(1024, device="cuda")
    a2 = torch.empty(M, 1024, device="cuda"); wo = torch.randn(1024, 1024, device="cuda") * 0.03
    out = torch.empty(M, 1024, device="cuda")
    t2 = graph_time(lambda: ops.sgemm_f32(x, wc, bias=b, row_bias=tv, act=ops.SACT_SILU, out=a2))
    t3 = graph_time(lambda: ops.sgemm_f32(a2, wo, bias=b, out=out))
    print(f"M={M}: action_time_mlp_in (K=32, SiLU) {t2:.1f} us, action_time_mlp_out (K=1024) {t3:.1f} us")
# verifier trajectory encoder shapes (400 rows)
for (M, N, K) in [(400, 1536, 512), (400, 512, 512), (400, 1024, 512), (400, 512, 1024)]:
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.03; o = torch.empty(M, N, device="cuda")
    print(f"sgemm {M}x{N}x{K}: {graph_time(lambda: ops.sgemm_f32(a, w, out=o)):.1f} us")
