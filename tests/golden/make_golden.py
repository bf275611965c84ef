"""Generates the committed golden fixtures from the ORACLE (parity unpinned: the reference cannot run
here and ships no numeric vectors for this path — see oracle/unet_oracle.py header).

Method = the reference's own tiny-model probe (src/bin/test/main.rs:51-54,128-140): deterministic
arb_tensor(dims) = sin(arange(prod(dims))) inputs through a tiny UNet. Weights are the seeded synthetic
set (sdxl_b200.synth_weights(TINY, seed=0), CPU generator => identical on every machine of this image).

    python tests/golden/make_golden.py          # everything
    python tests/golden/make_golden.py vae      # only the latent-decoder fixture
    python tests/golden/make_golden.py clip     # only the text-encoder fixture
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-xl-burn_b200"))
from oracle import unet_oracle as O  # noqa: E402
from oracle import vae_oracle as VO  # noqa: E402
from oracle import clip_oracle as CO  # noqa: E402
from sdxl_b200.config import TINY, TINY_CLIP, TINY_OPEN_CLIP, TINY_VAE  # noqa: E402
from sdxl_b200.weights import alphas_cumprod, synth_weights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def arb(*dims):
    return torch.sin(torch.arange(int(np.prod(dims)), dtype=torch.float32)).reshape(*dims)


def h16f(t):
    return t.to(torch.float16).float()


def vae():
    """Latent decoder (oracle/vae_oracle.py): TINY_VAE, latent = 0.13025 * 2 * sin(arange) at 8x8, B=2 -> image f32 + u8."""
    w = O.to_f32(synth_weights(TINY_VAE, seed=0))
    latent = arb(2, 4, 8, 8) * (2 * TINY_VAE.scale_factor)
    img = VO.decode_latent(TINY_VAE, w, latent)
    u8 = VO.latent_to_image(TINY_VAE, w, latent)
    np.savez(os.path.join(HERE, "tiny_vae_decode.npz"), latent=latent.numpy(), image=img.numpy(), u8=u8.numpy())
    print("tiny_vae_decode", img.shape, float(img.abs().mean()), float(u8.float().mean()))
    # encoder half: u8 image (a deterministic pattern) -> latent
    rgb = ((arb(1, 64, 96, 3) * 0.5 + 0.5) * 255.0).to(torch.uint8)
    lat = VO.image_to_latent(TINY_VAE, w, rgb)
    np.savez(os.path.join(HERE, "tiny_vae_encode.npz"), rgb=rgb.numpy(), latent=lat.numpy())
    print("tiny_vae_encode", lat.shape, float(lat.abs().mean()))


def clip():
    """Text encoders (oracle/clip_oracle.py): TINY_CLIP forward_hidden (penultimate) and TINY_OPEN_CLIP forward_hidden_pooled on
    two fixed token rows (CLIP-style end-of-text padding and OpenCLIP-style zero padding)."""
    rows = [[49406, 320, 1125, 539, 320, 2368, 49407], [49406, 17, 4, 256, 300, 301, 302, 303, 9, 49407]]
    t1 = torch.tensor([r + [49407] * (77 - len(r)) for r in rows])
    t2 = torch.tensor([r + [0] * (77 - len(r)) for r in rows])
    w1, w2 = O.to_f32(synth_weights(TINY_CLIP, seed=1)), O.to_f32(synth_weights(TINY_OPEN_CLIP, seed=2))
    h1 = CO.forward_hidden(TINY_CLIP, w1, t1, TINY_CLIP.n_layer - 1)
    h2, p2 = CO.forward_hidden_pooled(TINY_OPEN_CLIP, w2, t2, TINY_OPEN_CLIP.n_layer - 1)
    np.savez(os.path.join(HERE, "tiny_clip.npz"), tokens_clip=t1.numpy().astype(np.int32), tokens_open_clip=t2.numpy().astype(np.int32),
             hidden_clip=h1.numpy(), hidden_open_clip=h2.numpy(), pooled_open_clip=p2.numpy())
    print("tiny_clip", h1.shape, float(h1.abs().mean()), h2.shape, float(p2.abs().mean()))


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if "vae" in sys.argv[1:]:
        return vae()
    if "clip" in sys.argv[1:]:
        return clip()
    vae()
    clip()
    w = O.to_f32(synth_weights(TINY, seed=0))
    # 1) the reference's probe shapes: x[1,4,4,4], context[1,1,ctx], y[1,adm], t=[1]
    for tag, (B, h, wd, n_ctx, t) in {"": (1, 4, 4, 1, 1), "_16": (2, 16, 16, 77, 749)}.items():
        x, ctx, y = arb(B, 4, h, wd), h16f(arb(B, n_ctx, TINY.context_dim)), h16f(arb(B, TINY.adm_in_channels))
        out = O.unet_forward(TINY, w, x, torch.tensor([t]), ctx, y)
        np.savez(os.path.join(HERE, f"tiny_unet_forward{tag}.npz"), x=x.numpy(), context=ctx.numpy(), y=y.numpy(),
                 t=np.int32(t), out=out.numpy())
        print("tiny_unet_forward" + tag, out.shape, float(out.abs().mean()))
    # 2) sampler: 4 steps (t=999,749,499,249), cfg 7.5, injected noise
    B, n_ctx = 1, 5
    c = dict(context_full=h16f(arb(B, n_ctx, 24) * 0.9), unconditional_context_full=h16f(arb(n_ctx, 24).cos()),
             channel_context=h16f(arb(B, 8)), unconditional_channel_context=h16f(arb(8).cos()), resolution=(64, 64))
    noise = torch.randn(B, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    out = O.sample_latent(TINY, w, alphas_cumprod(), noise, O.OracleConditioning(**c), 7.5, 4)
    np.savez(os.path.join(HERE, "tiny_sample_latent.npz"), noise=noise.numpy(), out=out.numpy(), guidance=7.5, n_steps=4,
             **{k: v.numpy() for k, v in c.items() if k != "resolution"})
    print("tiny_sample_latent", float(out.abs().mean()))
    # 3) primitive KATs (inputs are sin(arange)); consumed by tests/test_oracle.py
    x = arb(2, 64, 4, 4)
    gam, bet = 1 + 0.1 * arb(64), 0.1 * arb(64).cos()
    q, k, v = arb(1, 6, 128), arb(1, 3, 128).cos(), arb(1, 3, 128) * 0.5
    np.savez(os.path.join(HERE, "primitives.npz"),
             gn=O.group_norm(x, gam, bet).numpy(), ln=O.layer_norm(arb(5, 64), gam, bet).numpy(),
             attn=O.qkv_attention(q, k, v, None, 2).numpy(), temb=O.timestep_embedding(torch.tensor([1, 999]), 64).numpy(),
             gelu=O.gelu_erf(arb(16)).numpy(), silu=O.silu(arb(16)).numpy())


if __name__ == "__main__":
    main()
