"""Prints the in-kernel timeline of the implicit-GEMM kernel for the UNet's main GEMM shapes (GPU only)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-xl-burn_b200"))
import sdxl_b200  # noqa: E402

ctx = sdxl_b200.Context(0)
shapes = [(2048, 1280, 1280, 0, 1), (2048, 5120, 1280, 0, 1), (2048, 1280, 3840, 0, 0), (2048, 1280, 10240, 1, 0),
          (8192, 640, 640, 0, 1), (8192, 640, 5120, 1, 0), (2048, 11520, 1280, 0, 1), (154, 2048, 2560, 0, 0)]
print("M K N geglu res | BN pair CMxCN nst | prologue->deps deps->first_data first_data->acc0 acc0->epi0 | producer_done(from prologue) producer_done(from deps) event_ns | MMA-ideal_ns")
for M, K, N, g, r in shapes:
    st = (C.c_uint64 * 9)()
    ctx.check(ctx.lib.sdxl_dbg_igemm_timeline(ctx.h, M, K, N, g, r, st), "timeline")
    s = list(st)
    cfg = s[8]
    BN, pair, CM, CN, nst = cfg & 0xFFFF, (cfg >> 16) & 0xF, (cfg >> 20) & 0xF, (cfg >> 24) & 0xF, (cfg >> 28) & 0xF
    ideal = 2.0 * M * K * N / 148 / (8192 * 1.9)  # ns at 8192 flop/clk/SM, 1.9 GHz, all SMs
    print(f"{M} {K} {N} {g} {r} | {BN} {pair} {CM}x{CN} {nst} | {s[1]-s[0]:6d} {s[2]-s[1]:6d} {s[3]-s[2]:7d} {s[4]-s[3]:6d} | "
          f"{s[6]-s[0]:7d} {s[6]-s[1]:7d} {s[7]:7d} | {ideal:7.0f}")
