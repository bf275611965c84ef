"""Per-K-block main-loop time of one GEMM shape vs the number of active CTAs (M) — separates a per-SM ingest limit from a
chip-wide L2 limit. usage: python tools/igemm_timeline2.py K N M1 M2 ..."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-xl-burn_b200"))
import sdxl_b200  # noqa: E402

ctx = sdxl_b200.Context(0)
K, N = int(sys.argv[1]), int(sys.argv[2])
for M in map(int, sys.argv[3:]):
    st = (C.c_uint64 * 16)()
    ctx.check(ctx.lib.sdxl_dbg_igemm_timeline(ctx.h, M, K, N, 0, 1, st), "timeline")
    s = list(st)
    cfg = s[8]
    BN, pair, nst = cfg & 0xFFFF, (cfg >> 16) & 0xF, (cfg >> 28) & 0xF
    kb = K // 64
    print(f"M={M} K={K} N={N} BN={BN} pair={pair} nst={nst}: first_data {s[2]-s[1]} ns, main loop {s[3]-s[2]} ns = {(s[3]-s[2])/kb:.1f} ns/kblock, event {s[7]} ns")
