"""A few launches of the roofline kernel (prefix gate/up GeGLU GEMM, M=2240 N=32768 K=2048) for `ncu --set full`."""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops

M, K, I = 2240, 2048, 16384
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
ws = [(torch.randn(2 * I, K, device="cuda") * 0.02).to(torch.bfloat16) for _ in range(3)]
o = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
for w in ws:
    ops.gemm_bf16(a, w, epilogue=ops.EPI_GEGLU, n_out=I, out=o)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for w in ws:
    ops.gemm_bf16(a, w, epilogue=ops.EPI_GEGLU, n_out=I, out=o)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
