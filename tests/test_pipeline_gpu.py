"""End-to-end `sample` flow on the GPU with tiny models (reference src/bin/sample/main.rs:225-285): text -> Embedder ->
Conditioning -> Diffuser::sample_latent -> LatentDecoder::latent_to_image, plus the inpainting branch
(image_to_latent -> sample_latent_with_inpainting), each stage checked against the oracle chain."""
import os

import numpy as np
import pytest
import torch

from sdxl_b200 import (TINY, TINY_CLIP, TINY_OPEN_CLIP, TINY_VAE, ClipConfig, ClipTextEncoder, Diffuser, Embedder, LatentDecoder,
                       OpenClipTokenizer, UNetConfig, synth_weights)
from oracle import clip_oracle as CO
from oracle import tokenizer_oracle as TO
from oracle import unet_oracle as O
from oracle import vae_oracle as VO

pytestmark = pytest.mark.gpu
MINI = os.path.join(os.path.dirname(__file__), "golden", "mini_bpe")

# text encoders whose widths add up to the tiny UNet's context_dim (24 is not reachable with head dim 64): use a UNet
# config sized for them instead — context 128+192, label 64 + 6*256
CLIP_A = TINY_CLIP
CLIP_B = TINY_OPEN_CLIP
UNET = UNetConfig(adm_in_channels=CLIP_B.embed_dim + 6 * 256, model_channels=64, channel_mults=(1, 2, 4), transformer_depths=(0, 1, 1),
                  context_dim=CLIP_A.n_state + CLIP_B.n_state)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_text_to_image_and_inpaint(ctx):
    wa, wb, wu, wv = (synth_weights(c, seed=s) for c, s in ((CLIP_A, 1), (CLIP_B, 2), (UNET, 3), (TINY_VAE, 0)))
    ea, eb = ClipTextEncoder(ctx, CLIP_A, wa), ClipTextEncoder(ctx, CLIP_B, wb)
    tok = OpenClipTokenizer(os.path.join(MINI, "mini_merges.txt"), os.path.join(MINI, "mini_vocab.txt"))
    emb = Embedder(ctx, ea, eb, tok, tok)
    dif = Diffuser(ctx, UNET, wu)
    vae = LatentDecoder(ctx, TINY_VAE, wv)
    text, res = "a photo of a cat", (64, 64)

    cond = emb.text_to_conditioning(text, res, (0, 0), res)
    noise = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    latent = dif.sample_latent(cond, 5.0, 6, noise=noise)
    rgb = vae.latent_to_image(latent)
    assert rgb.shape == (1, 32, 32, 3) and rgb.dtype == torch.uint8

    # oracle chain on the same inputs
    otok = TO.OpenClipTokenizer(os.path.join(MINI, "mini_merges.txt"), os.path.join(MINI, "mini_vocab.txt"))
    oc = CO.text_to_conditioning(CLIP_A, O.to_f32(wa), CLIP_B, O.to_f32(wb), otok, otok, TO.tokenize_text, text, res, (0, 0), res)
    h16 = lambda t: t.to(torch.float16).float()  # Conditioning::convert
    ocond = O.OracleConditioning(context_full=h16(oc["context_full"]), unconditional_context_full=h16(oc["unconditional_context_full"]),
                                 channel_context=h16(oc["channel_context"]),
                                 unconditional_channel_context=h16(oc["unconditional_channel_context"]), resolution=res)
    from sdxl_b200 import alphas_cumprod
    olat = O.sample_latent(UNET, O.to_f32(wu), alphas_cumprod(), noise, ocond, 5.0, 6)
    e = rel_err(latent, olat)
    print("pipeline latent rel err", e)
    assert e <= 5e-3
    oimg = VO.latent_to_image(TINY_VAE, O.to_f32(wv), olat).numpy().astype(np.int32)
    diff = np.abs(rgb.cpu().numpy().astype(np.int32) - oimg)
    print("pipeline image max diff", diff.max(), "equal fraction", (diff == 0).mean())
    assert diff.max() <= 3 and (diff <= 1).mean() >= 0.99

    # inpainting branch: reference image -> latent -> sample_latent_with_inpainting (mask true = keep generated)
    ref_rgb = torch.randint(0, 256, (1, 32, 32, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
    ref_lat = vae.image_to_latent(ref_rgb)
    assert ref_lat.shape == (1, 4, 8, 8)
    mask = torch.zeros(1, 4, 8, 8, dtype=torch.bool)
    mask[:, :, :3] = True
    out = dif.sample_latent_with_inpainting(cond, 5.0, 6, ref_lat, mask, seed=7)
    assert out.shape == (1, 4, 8, 8) and torch.isfinite(out).all()
    for o in (ea, eb, dif, vae):
        o.close()
