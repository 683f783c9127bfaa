"""GPU diagnostic: o_proj / down_proj of the denoise step as (split-K partials + rmsnorm_reduce) vs (fused-epilogue GEMM
+ rmsnorm), timed as pairs inside a CUDA graph with the weights cycled through > L2."""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops


def time_graph(fn, copies, iters=5):
    for i in range(copies):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(copies):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * copies) * 1e3


if __name__ == "__main__":
    copies = 24
    for name, M, N, K in [("o_e", 200, 1024, 2048), ("down_e", 200, 1024, 4096)]:
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16) for _ in range(copies)]
        resid = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
        gw = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
        h = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        row = [f"{name} M={M} N={N} K={K}:"]
        us = time_graph(lambda i: ops.gemm_bf16(a, ws[i], epilogue=ops.EPI_RESID, resid=resid, out=h), copies)
        row.append(f"fused gemm alone {us:5.1f}us")
        for S in (4, 8, 12, 16):
            us_g = time_graph(lambda i: ops.gemm_splitk_partial(a, ws[i], S), copies)
            p = ops.gemm_splitk_partial(a, ws[0], S)
            us_r = time_graph(lambda i: ops.rmsnorm_reduce(p, resid, gw), copies)
            us_pair = time_graph(lambda i: ops.rmsnorm_reduce(ops.gemm_splitk_partial(a, ws[i], S), resid, gw), copies)
            row.append(f"S{S}: gemm {us_g:5.1f} reduce {us_r:5.1f} pair {us_pair:5.1f}us")
        print(" | ".join(row), flush=True)
