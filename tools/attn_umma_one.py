"""A few launches of the tcgen05 attention kernels at the bench shapes for `ncu --set full`:
prefix (8 prompts x 280 tokens x 8 heads x 256), denoise (40 candidates x 5 tokens vs 280 + 5 keys), SigLIP (16 x 72, 256)."""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops
heads, hd = 8, 256
B, T = 8, 280
q = torch.randn(B, T, heads * hd, device="cuda").to(torch.bfloat16)
k = torch.randn(B, T, hd, device="cuda").to(torch.bfloat16)
v = torch.randn(B, T, hd, device="cuda").to(torch.bfloat16)
lens = torch.randint(264, 281, (B,), device="cuda", dtype=torch.int32)
R, K, S = 8, 5, 5
N = R * K
qd = torch.randn(N, S, heads * hd, device="cuda").to(torch.bfloat16)
k1 = torch.randn(N, S, hd, device="cuda").to(torch.bfloat16)
v1 = torch.randn(N, S, hd, device="cuda").to(torch.bfloat16)
tab = torch.randn(R, S, hd // 2, 2, device="cuda")
vt = ops.transpose_values(v)
qs = torch.randn(1, 256, 16 * 72, device="cuda").to(torch.bfloat16)
ks = torch.randn(1, 256, 16 * 72, device="cuda").to(torch.bfloat16)
vs = torch.randn(1, 256, 16 * 72, device="cuda").to(torch.bfloat16)


def run():
    ops.attention_umma(q, k, v, lens=lens)
    ops.attention(qd, k, v, heads=heads, kv_heads=1, head_dim=hd, kv0_len_dev=lens, q_per_kv_batch=K, k1=k1, v1=v1,
                  suffix_mask=True, rope=tab, vt0=vt, algo=3)
    ops.attention(qs, ks, vs, heads=16, kv_heads=16, head_dim=72, kv0_len=256)


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
