"""GPU diagnostic: tcgen05 prefix attention vs the mma.sync kernel at the bench shape (8 prompts x 280 tokens)."""
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops
B, T, heads, hd = 8, 280, 8, 256
q = torch.randn(B, T, heads * hd, device="cuda").to(torch.bfloat16)
k = torch.randn(B, T, hd, device="cuda").to(torch.bfloat16)
v = torch.randn(B, T, hd, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
for name, fn in (("umma", lambda: ops.attention_umma(q, k, v, lens=lens)),
                 ("mma.sync", lambda: ops.attention(q, k, v, heads=heads, kv_heads=1, head_dim=hd, kv0_len_dev=lens))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(name, "us per call (incl. wrapper overhead):", e0.elapsed_time(e1) / 20 * 1e3)
