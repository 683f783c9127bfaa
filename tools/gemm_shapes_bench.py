"""Time the bf16 GEMM kernel variants on given shapes with cold weights (18 different weight matrices in rotation, as in
the denoise loop).  usage: python tools/gemm_shapes_bench.py M,N,K[,epi] ...   (epi: store | resid | geglu64)"""
import sys

import torch

sys.path.insert(0, ".")
from cover_vla_b200 import ops  # noqa: E402

VARIANTS = {"auto": 0, "1sm-64": 64, "1sm-128": 128, "1sm-256": 256, "pair-256": 512, "pair-128": 384}


def bench(M, N, K, epi):
    L = 18
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16) for _ in range(L)]
    r = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N if epi != "geglu64" else N // 2, device="cuda", dtype=torch.bfloat16)
    res = {}
    for name, bn in VARIANTS.items():
        if epi == "geglu64" and bn not in (128, 384):
            continue

        def run(w):
            if epi == "store":
                ops.gemm_bf16(a, w, out=out, force_bn=bn)
            elif epi == "resid":
                ops.gemm_bf16(a, w, epilogue=ops.EPI_RESID, resid=r, out=out, force_bn=bn)
            else:
                ops.gemm_bf16(a, w, epilogue=ops.EPI_GEGLU64, n_out=N // 2, out=out, force_bn=bn)
        try:
            for w in ws:
                run(w)
        except Exception as e:  # noqa: BLE001
            res[name] = f"n/a ({str(e)[:40]})"
            continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            for w in ws:
                run(w)
        e1.record()
        torch.cuda.synchronize()
        res[name] = round(e0.elapsed_time(e1) * 1000 / (5 * L), 2)
    gf = 2.0 * M * N * K / 1e9
    print(f"M={M} N={N} K={K} {epi} ({gf:.1f} GF): " + "  ".join(f"{k}={v}" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    for arg in sys.argv[1:]:
        parts = arg.split(",")
        bench(int(parts[0]), int(parts[1]), int(parts[2]), parts[3] if len(parts) > 3 else "store")
