"""GPU diagnostic for the tcgen05 GEMM: error pattern + quick timing. Not a test; prints a report."""
import sys, time
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops


def report(M, N, K, bn):
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    y = ops.gemm_bf16(a, w, force_bn=bn)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    err = (y.float() - ref).abs()
    rel = ((y.float() - ref).norm() / ref.norm()).item()
    bad = err > 0.05
    msg = f"M={M} N={N} K={K} bn={bn}: rel={rel:.3e} max={err.max().item():.3e} bad={bad.float().mean().item():.4f}"
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()[:16].tolist()
        cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        msg += f" bad_rows={rows} bad_cols={cols}"
        msg += f" y[0,:4]={y[0,:4].tolist()} ref[0,:4]={ref[0,:4].tolist()}"
    print(msg, flush=True)


def timeit(M, N, K, bn=0, epi=ops.EPI_STORE, iters=20):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm_bf16(a, w, out=out, force_bn=bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        ops.gemm_bf16(a, w, out=out, force_bn=bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    gb = (M * K + N * K + M * N) * 2 / ms / 1e6
    # cuBLAS for comparison (diagnostic only)
    for _ in range(3):
        torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"time M={M} N={N} K={K} bn={bn}: {ms*1e3:.1f} us  {tf:.1f} TF/s  {gb:.0f} GB/s | cublas {ms2*1e3:.1f} us {2.0*M*N*K/ms2/1e9:.1f} TF/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for (M, N, K) in [(128, 64, 64), (128, 256, 64), (128, 128, 256), (256, 512, 512), (200, 2560, 1024), (2624, 2560, 2048), (256, 1152, 4304)]:
        for bn in (64, 128, 256):
            try:
                report(M, N, K, bn)
            except Exception as e:
                print("EXC", M, N, K, bn, repr(e), flush=True)
    for (M, N, K) in [(2624, 2560, 2048), (2624, 32768, 2048), (2624, 2048, 16384), (200, 2560, 1024), (200, 8192, 1024), (200, 1024, 4096), (8192, 8192, 8192)]:
        for bn in (64, 128, 256):
            timeit(M, N, K, bn)
