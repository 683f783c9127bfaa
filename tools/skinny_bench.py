"""GPU diagnostic: skinny vs general GEMM on the denoise / vision / text-tower shapes (weights cycled through > L2)."""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops


def timeit(M, N, K, force, epi=ops.EPI_STORE, copies=24, iters=5):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16) for _ in range(copies)]
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    kw = dict(n_out=N // 2) if epi in (ops.EPI_GEGLU, ops.EPI_GEGLU64) else {}
    if kw:
        out = torch.empty(M, N // 2, device="cuda", dtype=torch.bfloat16)
    for w in ws:
        ops.gemm_bf16(a, w, out=out, force_bn=force, epilogue=epi, **kw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for w in ws:
            ops.gemm_bf16(a, w, out=out, force_bn=force, epilogue=epi, **kw)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (iters * copies) * 1e3
    gb = N * K * 2 / us / 1e3
    return us, gb


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    shapes = [("qkv_e", 200, 2560, 1024, 0), ("o_e", 200, 1024, 2048, 0), ("gateup_e", 200, 8192, 1024, 1), ("down_e", 200, 1024, 4096, 0),
              ("qkv_v", 256, 3456, 1152, 0), ("o_v", 256, 1152, 1152, 0), ("fc1_v", 256, 4304, 1152, 0), ("fc2_v", 256, 1152, 4304, 0),
              ("qkv_t", 64, 3072, 1024, 0), ("fc1_t", 64, 4096, 1024, 0), ("fc2_t", 64, 1024, 4096, 0)]
    for name, M, N, K, geglu in shapes:
        row = [f"{name:9s} M={M} N={N} K={K}:"]
        if geglu:
            us, gb = timeit(M, N, K, 0, ops.EPI_GEGLU)
            row.append(f"general {us:6.1f}us {gb:5.0f}GB/s")
            for s in (-100, -1, -2, -4):
                us, gb = timeit(M, N, K, s, ops.EPI_GEGLU64)
                row.append(f"S{-s if s != -100 else 'auto'} {us:6.1f}us {gb:5.0f}GB/s")
        else:
            for f in (64, 128):
                us, gb = timeit(M, N, K, f)
                row.append(f"bn{f} {us:6.1f}us")
            for s in (-100, -2, -4, -6, -8, -16):
                try:
                    us, gb = timeit(M, N, K, s)
                    row.append(f"S{-s if s != -100 else 'auto'} {us:6.1f}us {gb:5.0f}GB/s")
                except Exception as e:
                    row.append(f"S{-s} ERR {str(e)[:40]}")
        print(" | ".join(row), flush=True)
