"""Prefix-shaped self attention launches (8 prompts x 280 tokens, 8 Q heads : 1 KV head, head_dim 256) for ncu."""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops
B, T, heads, hd = 8, 280, 8, 256
q = torch.randn(B, T, heads * hd, device="cuda").to(torch.bfloat16)
k = torch.randn(B, T, hd, device="cuda").to(torch.bfloat16)
v = torch.randn(B, T, hd, device="cuda").to(torch.bfloat16)
lens = torch.randint(264, 281, (B,), device="cuda", dtype=torch.int32)
for _ in range(4):
    ops.attention(q, k, v, heads=heads, kv_heads=1, head_dim=hd, kv0_len_dev=lens)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(20):
    ops.attention(q, k, v, heads=heads, kv_heads=1, head_dim=hd, kv0_len_dev=lens)
e1.record(); torch.cuda.synchronize()
print("prefix attention us:", e0.elapsed_time(e1) / 20 * 1e3)
