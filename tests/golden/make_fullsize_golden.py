#!/usr/bin/env python
"""Generates the full-size golden latents under tests/golden/ with the CPU f32 oracle (oracle/unet_oracle.py).

    python tests/golden/make_fullsize_golden.py [--threads N] [--only fwd,config1,config2,refiner,inpaint]

Run once, offline, on host cores (no GPU): SDXL-base / refiner sized synthetic weights (2.57 B / 2.26 B parameters,
seeded on the CPU generator), inputs from tests/fullsize_cases.py. 1 + 8 + 62 + 10 + 20 UNet forwards; about an
hour on 8 cores. The outputs are the fixtures tests/test_fullsize_parity_gpu.py compares the CUDA path against:

    base_fwd_1024.npz      one UNet::forward at 1024^2 (latent 128x128), t = 999
    base_config1.npz       BASELINE config 1: 256^2, 4 steps, cfg 1.0 and 7.5 (final latents)
    base_config2.npz       BASELINE config 2: 1024^2, n = 30 (31 iterations), cfg 7.5 (final latent + checkpoints)
    refiner_10step.npz     refine_latent(step_start 800, n 50) = 10 refiner iterations at 1024^2
    base_inpaint10.npz     10-iteration inpainting run at 1024^2 (mask = top 25 latent rows), cfg 7.5
    vae_1024.npz           LatentDecoder::decode_latent of a 128x128 latent (digests of the 1024^2 image, fullsize_cases.vae_image_digest)
                           and LatentDecoder::image_to_latent of a 1024^2 u8 image (full latent)

PARITY UNPINNED: these come from the restated oracle, not from the reference binary (which cannot be built here,
see DESIGN.md); they pin the CUDA path to the oracle at BASELINE.json's own sizes.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "stable-diffusion-xl-burn_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import fullsize_cases as FC  # noqa: E402
import sdxl_b200  # noqa: E402  (config + synthetic weights only; the .so is never loaded here)
from oracle import unet_oracle as O  # noqa: E402


def log(msg):
    print(f"[{time.strftime('%H:%M:%S')}] {msg}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="fwd,config1,config2,refiner,inpaint,vae")
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    only = set(args.only.split(","))
    alphas = sdxl_b200.alphas_cumprod()
    t_all = time.time()

    with torch.no_grad():
        if only & {"fwd", "config1", "config2", "inpaint"}:
            cfg = sdxl_b200.SDXL_BASE
            t0 = time.time()
            w = O.to_f32(sdxl_b200.synth_weights(cfg, seed=FC.BASE_WEIGHT_SEED, device="cpu"))
            log(f"base weights generated in {time.time() - t0:.0f} s")

            if "fwd" in only:
                x, ctx, y = FC.fwd_1024_inputs()
                t0 = time.time()
                out = O.unet_forward(cfg, w, x, torch.tensor([FC.FWD_1024_T]), ctx, y)
                log(f"1024^2 forward: {time.time() - t0:.1f} s")
                np.savez(os.path.join(HERE, "base_fwd_1024.npz"), out=out.numpy(), t=FC.FWD_1024_T)

            if "config1" in only:
                c = FC.CONFIG1
                cond = O.OracleConditioning(**FC.base_conditioning(c["res"]))
                outs = {}
                for g in c["guidances"]:
                    t0 = time.time()
                    outs[f"out_cfg{g}"] = O.sample_latent(cfg, w, alphas, FC.base_noise(c["res"]), cond, g, c["n_steps"]).numpy()
                    log(f"config 1 cfg {g}: {time.time() - t0:.1f} s")
                np.savez(os.path.join(HERE, "base_config1.npz"), **outs)

            if "config2" in only:
                c = FC.CONFIG2
                cond = O.OracleConditioning(**FC.base_conditioning(c["res"]))
                trace = {}
                t0 = time.time()

                def tr(it, lat):
                    log(f"config 2 iteration {it}: {time.time() - t0:.0f} s, |x| = {float(lat.norm()):.4f}")
                    if it in c["checkpoints"]:
                        trace[f"it{it}"] = lat.numpy().copy()
                        np.savez(os.path.join(HERE, "base_config2.partial.npz"), **trace)   # resumable evidence if interrupted
                out = O.sample_latent(cfg, w, alphas, FC.base_noise(c["res"]), cond, c["guidance"], c["n_steps"], trace=tr)
                np.savez(os.path.join(HERE, "base_config2.npz"), out=out.numpy(), **trace)
                os.remove(os.path.join(HERE, "base_config2.partial.npz"))

            if "inpaint" in only:
                c = FC.INPAINT
                cond = O.OracleConditioning(**FC.base_conditioning(c["res"]))
                ref, mask, init, step_noise = FC.inpaint_inputs()
                t0 = time.time()
                out = O.sample_latent_with_inpainting(cfg, w, alphas, init, cond, c["guidance"], c["n_steps"], ref, mask, list(step_noise))
                log(f"inpaint 10 iterations: {time.time() - t0:.0f} s")
                np.savez(os.path.join(HERE, "base_inpaint10.npz"), out=out.numpy())
            del w

        if "refiner" in only:
            cfg = sdxl_b200.SDXL_REFINER
            t0 = time.time()
            w = O.to_f32(sdxl_b200.synth_weights(cfg, seed=FC.REFINER_WEIGHT_SEED, device="cpu"))
            log(f"refiner weights generated in {time.time() - t0:.0f} s")
            c = FC.REFINER
            lat, noise, cond = FC.refiner_inputs()
            t0 = time.time()
            out = O.refine_latent(cfg, w, alphas, lat, O.OracleConditioning(**cond), c["guidance"], c["step_start"], c["n_steps"], noise)
            log(f"refiner 10 iterations: {time.time() - t0:.0f} s")
            np.savez(os.path.join(HERE, "refiner_10step.npz"), out=out.numpy())
        if "vae" in only:
            from oracle import vae_oracle as VO
            cfg = sdxl_b200.SDXL_VAE
            w = O.to_f32(sdxl_b200.synth_weights(cfg, seed=FC.VAE_WEIGHT_SEED, device="cpu"))
            lat, rgb = FC.vae_1024_inputs(cfg.scale_factor)
            t0 = time.time()
            img = VO.decode_latent(cfg, w, lat)
            log(f"VAE decode 1024^2: {time.time() - t0:.0f} s")
            pool, samp = FC.vae_image_digest(img)
            t0 = time.time()
            enc = VO.image_to_latent(cfg, w, rgb)
            log(f"VAE encode 1024^2: {time.time() - t0:.0f} s")
            np.savez(os.path.join(HERE, "vae_1024.npz"), pool=pool.numpy(), samp=samp.numpy(), norm=float(img.double().norm()),
                     latent=enc.numpy())
    log(f"done in {time.time() - t_all:.0f} s")


if __name__ == "__main__":
    main()
