#!/usr/bin/env python
"""Where the time between back-to-back GEMM launches goes (GPU only): every CTA of the 2-CTA GEMM kernel stamps %globaltimer at
its entry, at the end of its prologue, when its programmatic dependency is resolved and at its exit; n launches of the same GEMM
are issued back to back with programmatic dependent launch, as the step's graph does.
Stamps are taken by thread 0 (producer warp) except "epilogue done" (first epilogue warp). A stamp that directly follows a CTA
barrier marks thread 0's ARRIVAL there (BAR.SYNC.DEFER_BLOCKING lets the timer read issue before the barrier completes): "barriers
init", "prologue done" and "producer at final barrier" are arrival times; "cluster sync", "deps resolved", "epilogue done" and
"exit" follow blocking waits and are completion times.

    python tools/igemm_gaps.py [M K N residual]     (default: the UNet's 2048 x 1280 x 1280 out-projection with residual)
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-xl-burn_b200"))
import sdxl_b200  # noqa: E402

ctx = sdxl_b200.Context(0)
shapes = [tuple(int(a) for a in sys.argv[1:5])] if len(sys.argv) >= 5 else [(2048, 1280, 1280, 1), (2048, 1280, 3840, 0), (8192, 640, 640, 1), (2048, 5120, 1280, 1)]
n = 6
NAMES = ["entry", "barriers init", "cluster sync", "prologue done", "deps resolved", "epilogue done", "producer at final barrier", "exit"]
for M, K, N, r in shapes:
    out = (C.c_int64 * (n * 16 + 1))()
    ctx.check(ctx.lib.sdxl_dbg_igemm_gaps(ctx.h, M, K, N, r, n, out), "gaps")
    v = list(out)
    print(f"M={M} K={K} N={N} residual={r}: grid {v[n * 16]} CTAs; ns, 'first..last' over the CTAs, relative to the first CTA entry of launch 0")
    for i in range(n):
        row = v[i * 16:i * 16 + 16]
        prev_exit = v[(i - 1) * 16 + 15] if i else 0
        print(f"  launch {i}: " + " | ".join(f"{NAMES[k]} {row[2 * k] - prev_exit}..{row[2 * k + 1] - prev_exit}" for k in range(8)) + (f" || period {row[15] - prev_exit}" if i else ""))
