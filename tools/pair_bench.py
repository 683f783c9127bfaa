"""GPU diagnostic: 1-CTA vs CTA-pair (cta_group::2) GEMM at the prefix shapes (weights cycled through > L2)."""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops


def timeit(M, N, K, bn, epi=ops.EPI_STORE, copies=6, iters=5):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    ws = [(torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16) for _ in range(copies)]
    kw = dict(n_out=N // 2) if epi == ops.EPI_GEGLU else {}
    out = torch.empty(M, N // 2 if kw else N, device="cuda", dtype=torch.bfloat16)
    for w in ws:
        ops.gemm_bf16(a, w, out=out, force_bn=bn, epilogue=epi, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        for w in ws:
            ops.gemm_bf16(a, w, out=out, force_bn=bn, epilogue=epi, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (iters * copies) * 1e3
    return us, 2.0 * M * N * K / us / 1e6


for name, M, N, K, epi in [("gateup", 2240, 32768, 2048, ops.EPI_GEGLU), ("down", 2240, 2048, 16384, ops.EPI_STORE),
                           ("qkv", 2240, 2560, 2048, ops.EPI_STORE), ("o", 2240, 2048, 2048, ops.EPI_STORE)]:
    r = [f"{name:7s} M={M} N={N} K={K}:"]
    for bn in (256, 512):
        us, tf = timeit(M, N, K, bn, epi)
        r.append(f"{'pair' if bn == 512 else '1cta'} {us:7.1f}us {tf:7.1f}TF/s")
    print(" | ".join(r), flush=True)
