#!/usr/bin/env python
"""Per-launch timing of back-to-back attention launches (looks for rare long stalls): python tools/attn_stall.py T S heads n"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stable-diffusion-xl-burn_b200")):
    sys.path.insert(0, p)
import torch
import sdxl_b200
T, S, nh, n = (int(a) for a in sys.argv[1:5])
ctx = sdxl_b200.Context(0)
B, C = 2, nh * 64
g = torch.Generator(device="cuda").manual_seed(1)
q, k, vv = (torch.randn(B, L, C, device="cuda", generator=g).half() for L in (T, S, S))
o = torch.empty_like(q)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ctx.enter()
ev[0].record(ctx.stream)
for i in range(n):
    assert ctx.lib.sdxl_qkv_attention(ctx.h, q.data_ptr(), k.data_ptr(), vv.data_ptr(), None, B, T, S, C, nh, o.data_ptr()) == 0
    ev[i + 1].record(ctx.stream)
ctx.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(n)]
print(f"T{T} S{S}: min {min(ts):.1f} us median {sorted(ts)[n//2]:.1f} max {max(ts):.1f}; >3x median: {[ (i, round(t)) for i, t in enumerate(ts) if t > 3 * sorted(ts)[n//2]]}")
