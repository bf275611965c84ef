#!/usr/bin/env python
"""In-kernel timeline of the attention kernel (CTA 0): per key block, when S became ready, when the exp phase ended, when P was
handed to the MMA warp, and when the MMA warp saw P / finished issuing. Clocks relative to the first stamp.
    python tools/attn_timeline.py [T S heads]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stable-diffusion-xl-burn_b200")):
    sys.path.insert(0, p)
import sdxl_b200  # noqa: E402

T, S, nh = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 1024, 20)
ctx = sdxl_b200.Context(0)
buf = (C.c_longlong * (3 * 1024))()
ctx.check(ctx.lib.sdxl_dbg_attention_timeline(ctx.h, 2, T, S, nh, buf), "timeline")
st = [[[buf[r * 1024 + j * 4 + k] for k in range(4)] for j in range(256)] for r in range(3)]
t0 = min(v for r in st for j in r for v in j if v > 0)
print(f"T={T} S={S} heads={nh}; clocks since the first stamp")
print(" blk |   A: S ready  exp done  P handed  (item out) |   B: S ready  exp done  P handed  (item out) | MMA: P_A seen  A issued  P_B seen  B issued")
for j in range(256):
    if not any(st[r][j][0] for r in range(3)):
        break
    f = lambda v: f"{v - t0:9d}" if v > 0 else "        -"  # noqa: E731
    print(f"{j:4d} | " + " ".join(f(st[0][j][k]) for k in range(4)) + " | " + " ".join(f(st[1][j][k]) for k in range(4)) + " | " + " ".join(f(st[2][j][k]) for k in range(4)))

