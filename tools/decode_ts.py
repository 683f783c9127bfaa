"""Phase timeline of one cluster decode attention launch at the denoise shape."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from cover_vla_b200 import ops, _lib
lib = _lib.load()
lib.cvb_debug_set_timestamps.argtypes = [C.c_void_p]
R, K, S, P, heads, hd = 8, 5, 5, int(__import__('os').environ.get('P', 280)), 8, 256
N = R * K
q = torch.randn(N, S, heads * hd, device="cuda").to(torch.bfloat16)
k0 = torch.randn(R, P, hd, device="cuda").to(torch.bfloat16)
v0 = torch.randn(R, P, hd, device="cuda").to(torch.bfloat16)
k1 = torch.randn(N, S, hd, device="cuda").to(torch.bfloat16)
v1 = torch.randn(N, S, hd, device="cuda").to(torch.bfloat16)
lens = torch.randint(264, min(P, 280) + 1, (R,), device="cuda", dtype=torch.int32)
tab = torch.randn(R, S, hd // 2, 2, device="cuda")
kw = dict(heads=heads, kv_heads=1, head_dim=hd, kv0_len_dev=lens, q_per_kv_batch=K, k1=k1, v1=v1, suffix_mask=True, rope=tab)
for _ in range(3):
    ops.attention(q, k0, v0, **kw)
torch.cuda.synchronize()
import os
kw["vt0"] = ops.transpose_values(v0)
if os.environ.get("TS"):
  kw["algo"] = int(os.environ["TS"])
  ts = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
  lib.cvb_debug_set_timestamps(C.c_void_p(ts.data_ptr()))
  for rep in range(2):
    ts.zero_()
    ops.attention(q, k0, v0, **kw)
    torch.cuda.synchronize()
    t = ts.view(-1, 8).cpu()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t[:, :7] - t0).float() / 1e3
    print(f"{t.shape[0]} CTAs; us since first CTA start (min / median / max)")
    names = (["start", "staged", "qk+stats", "bar1", "pv+sent", "bar2", "final"] if kw["algo"] == 2 else
             ["start", "dep_ok", "staged", "s_full", "sums", "p_ready", "o_full", "stored"])
    rel = (t[:, :len(names)] - t0).float() / 1e3
    for i, n in enumerate(names):
        c = rel[:, i]
        print(f"  {n:9s} {c.min():7.2f} {c.median():7.2f} {c.max():7.2f}")
  lib.cvb_debug_set_timestamps(C.c_void_p(0))
kw["vt0"] = ops.transpose_values(v0)
for algo in (1, 2, 3):
  kw["algo"] = algo
  e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
      for _ in range(20):
          ops.attention(q, k0, v0, **kw)
  g.replay(); torch.cuda.synchronize()
  e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
  print("algo", algo, "graph avg us per launch:", e0.elapsed_time(e1) / 20 * 1e3)
  