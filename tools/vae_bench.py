"""Times LatentDecoder.decode_latent at 1024^2 (latent 128^2) on one GPU: CUDA-graph replay time, algorithmic TFLOP/s,
per-kind profile, optional per-op CSV. Analysis aid (the headline bench is bench.py)."""
import json
import sys
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-xl-burn_b200"))
import sdxl_b200 as S  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dump = sys.argv[2] if len(sys.argv) > 2 else None
    ctx = S.Context(0)
    d = S.LatentDecoder(ctx, S.SDXL_VAE, S.synth_weights(S.SDXL_VAE, seed=7))
    lat = torch.randn(B, 4, 128, 128, device="cuda") * S.SDXL_VAE.scale_factor
    for _ in range(3):
        d.decode_latent(lat)
    ctx.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 10
    with torch.cuda.stream(ctx.stream):
        ev[0].record()
        for _ in range(n):
            d.decode_latent(lat)
        ev[1].record()
    ctx.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    prof = d.profile_plan()
    if dump:
        d.profile_dump(dump)
    print(json.dumps({"workload": f"vae_decode_1024x1024_bs{B}", "ms": ms, "images_per_s": B * 1e3 / ms,
                      "tflops": d.plan_flops / ms / 1e9, "by_kernel_ms": {k: round(v["ms"], 3) for k, v in prof.items()}}))


if __name__ == "__main__":
    main()
