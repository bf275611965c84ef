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


def test_encoder_flops_match_survey():
    assert abs(VO.encoder_flops(SDXL_VAE, 1024, 1024) / 1e12 - 4.88) < 0.01


def test_padded_conv_is_bottom_right_padding():
    """PaddedConv2d(3, stride 2, pad (0,1,0,1)) == zero-pad one row/column at the bottom/right, then a valid stride-2 conv."""
    w = O.to_f32(synth_weights(TINY_VAE, seed=0))
    p = "encoder/blocks/0/downsampler"
    for hw in ((10, 12), (8, 8), (6, 14)):
        x = torch.randn(2, 64, *hw, generator=torch.Generator().manual_seed(hw[0]))
        want = F.conv2d(F.pad(x, (0, 1, 0, 1)), w[f"{p}/conv/weight"], w[f"{p}/conv/bias"], stride=2)
        got = VO.padded_conv2d(x, w, p)
        assert got.shape == want.shape == (2, 64, hw[0] // 2, hw[1] // 2) and torch.allclose(got, want, atol=1e-6)


def test_encode_keeps_mean_channels_and_scales():
    w = O.to_f32(synth_weights(TINY_VAE, seed=0))
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(4))
    z = F.conv2d(VO.encoder_forward(TINY_VAE, w, x), w["quant_conv/weight"], w["quant_conv/bias"])
    assert z.shape == (1, 8, 16, 16)
    assert torch.allclose(VO.encode_image(TINY_VAE, w, x), z[:, :4] * TINY_VAE.scale_factor)
    rgb = torch.randint(0, 256, (1, 64, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
    want = VO.encode_image(TINY_VAE, w, (rgb.float() / 255.0).permute(0, 3, 1, 2) * 2.0 - 1.0)
    assert torch.equal(VO.image_to_latent(TINY_VAE, w, rgb), want)


def test_vae_encode_golden_reproduces():
    g = np.load(os.path.join(GOLD, "tiny_vae_encode.npz"))
    w = O.to_f32(synth_weights(TINY_VAE, seed=0))
    lat = VO.image_to_latent(TINY_VAE, w, torch.from_numpy(g["rgb"]))
    assert np.allclose(lat.numpy(), g["latent"], atol=2e-5, rtol=1e-5)
