"""A handful of skinny GEMM launches for an ncu capture: python tools/skinny_one.py M N K split"""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops
M, N, K, split = (int(x) for x in sys.argv[1:5])
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16) for _ in range(12)]
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for w in ws:
    ops.gemm_bf16(a, w, out=out, force_bn=split)
torch.cuda.synchronize()
