"""CPU checks of the latent-decoder oracle (oracle/vae_oracle.py): independent PyTorch implementations of the same
published ops, the committed golden fixture, and the FLOP total SURVEY.md §8(f) derives from the reference."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import unet_oracle as O
from oracle import vae_oracle as VO
from sdxl_b200 import SDXL_VAE, TINY_VAE, synth_weights, vae_decoder_tensor_specs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_decoder_flops_match_survey():
    # 10.47 TFLOP at 1024^2; 49.49 M decoder parameters follow from the hard-coded widths
    assert abs(VO.decoder_flops(SDXL_VAE, 128, 128) / 1e12 - 10.47) < 0.01
    n = sum(int(np.prod(s[1])) for s in vae_decoder_tensor_specs(SDXL_VAE))
    assert n == 49_490_199


def test_attention_block_vs_torch_sdpa():
    """ConvSelfAttentionBlock == GN -> 1x1 convs -> F.scaled_dot_product_attention (single head) -> 1x1 + x."""
    torch.manual_seed(0)
    w = O.to_f32(synth_weights(TINY_VAE, seed=0))
    p = "decoder/mid/attn"
    x = torch.randn(2, 128, 8, 8)
    got = VO.conv_self_attention_block(x, w, p)
    h = F.group_norm(x, 32, w[f"{p}/norm/weight"], w[f"{p}/norm/bias"], eps=1e-5)
    q, k, v = (F.conv2d(h, w[f"{p}/{n}/weight"], w[f"{p}/{n}/bias"]).flatten(2).transpose(1, 2) for n in "qkv")
    a = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)
    a = a.transpose(1, 2).reshape(2, 128, 8, 8)
    want = x + F.conv2d(a, w[f"{p}/proj_out/weight"], w[f"{p}/proj_out/bias"])
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)


def test_decoder_block_upsample_is_nearest():
    w = O.to_f32(synth_weights(TINY_VAE, seed=0))
    p = "decoder/blocks/0"
    x = torch.randn(1, 128, 4, 4, generator=torch.Generator().manual_seed(1))
    got = VO.decoder_block(x, w, p)
    y = x
    for r in ("res1", "res2", "res3"):
        y = VO.resnet_block(y, w, f"{p}/{r}")
    want = F.conv2d(F.interpolate(y, scale_factor=2, mode="nearest"), w[f"{p}/upsampler/weight"], w[f"{p}/upsampler/bias"], padding=1)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    # last block has no upsampler
    assert VO.decoder_block(torch.randn(1, 64, 4, 4), w, "decoder/blocks/2").shape == (1, 64, 4, 4)


def test_latent_to_image_truncates_and_clamps():
    w = O.to_f32(synth_weights(TINY_VAE, seed=0))
    lat = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(2)) * 3.0  # large: forces clamping
    img = VO.decode_latent(TINY_VAE, w, lat)
    u8 = VO.latent_to_image(TINY_VAE, w, lat)
    ref = np.floor(np.clip(((img.permute(0, 2, 3, 1).numpy().astype(np.float32) + 1.0) / 2.0) * 255.0, 0, 255)).astype(np.uint8)
    assert u8.shape == (1, 32, 32, 3) and (u8.numpy() == ref).all()
    assert (u8 == 0).any() or (u8 == 255).any()


def test_vae_golden_reproduces():
    g = np.load(os.path.join(GOLD, "tiny_vae_decode.npz"))
    w = O.to_f32(synth_weights(TINY_VAE, seed=0))
    img = VO.decode_latent(TINY_VAE, w, torch.from_numpy(g["latent"]))
    assert np.allclose(img.numpy(), g["image"], atol=2e-5, rtol=1e-5)
