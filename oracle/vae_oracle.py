"""ORACLE — test infrastructure only. Never imported by the product path (sdxl_b200 / libsdxl_b200.so).

CPU f32 restatement (PyTorch tensor ops) of the reference's autoencoder, written line-by-line from
/root/reference (Gadersd/stable-diffusion-xl-burn): src/model/autoencoder/mod.rs (Autoencoder::{decode_latent,
encode_image}, Decoder, Encoder, Mid, ResnetBlock, ConvSelfAttentionBlock, DecoderBlock, EncoderBlock, PaddedConv2d) and
src/model/stablediffusion/mod.rs:199-266 (LatentDecoder::{decode_latent, latent_to_image, encode_image, image_to_latent}).

    *** PARITY UNPINNED *** — the reference cannot be built here (no cargo/rustc, un-vendored burn/tch crates) and
    ships no numeric golden for the autoencoder; goldens under tests/golden/ come from THIS file
    (tests/golden/make_golden.py). tests/test_oracle.py cross-checks its primitives against independent PyTorch
    implementations of the same published ops (F.group_norm, F.scaled_dot_product_attention, F.interpolate).

The reference runs this module in f32 (`LatentDecoder<Backend>` with `Backend = LibTorch<f32>`,
src/bin/sample/main.rs), which is what this file does. Weight names/layouts: autoencoder/load.rs dump tree,
conv weights OIHW (python/save.py:56-72).
"""
from __future__ import annotations

from typing import Dict

import torch

from .unet_oracle import conv2d, group_norm, qkv_attention, silu

W = Dict[str, torch.Tensor]


def resnet_block(x: torch.Tensor, w: W, p: str) -> torch.Tensor:
    """ResnetBlock::forward, autoencoder/mod.rs:507-524."""
    h = conv2d(silu(group_norm(x, w[f"{p}/norm1/weight"], w[f"{p}/norm1/bias"])), w, f"{p}/conv1")
    h = conv2d(silu(group_norm(h, w[f"{p}/norm2/weight"], w[f"{p}/norm2/bias"])), w, f"{p}/conv2")
    if f"{p}/nin_shortcut/weight" in w:
        return conv2d(x, w, f"{p}/nin_shortcut", padding=0) + h
    return x + h


def conv_self_attention_block(x: torch.Tensor, w: W, p: str) -> torch.Tensor:
    """ConvSelfAttentionBlock::forward, autoencoder/mod.rs:548-586: GN -> 1x1 q/k/v -> single-head attention over
    the h*w positions (qkv_attention with n_head = 1, no mask) -> 1x1 proj_out, plus the input."""
    n_batch, n_channel, height, width = x.shape
    h = group_norm(x, w[f"{p}/norm/weight"], w[f"{p}/norm/bias"])

    def tok(name: str) -> torch.Tensor:
        return conv2d(h, w, f"{p}/{name}", padding=0).reshape(n_batch, n_channel, height * width).transpose(1, 2)

    wv = qkv_attention(tok("q"), tok("k"), tok("v"), None, 1)
    wv = wv.transpose(1, 2).reshape(n_batch, n_channel, height, width)
    return x + conv2d(wv, w, f"{p}/proj_out", padding=0)


def mid(x: torch.Tensor, w: W, p: str) -> torch.Tensor:
    """Mid::forward, autoencoder/mod.rs:445-452."""
    x = resnet_block(x, w, f"{p}/block_1")
    x = conv_self_attention_block(x, w, f"{p}/attn")
    return resnet_block(x, w, f"{p}/block_2")


def decoder_block(x: torch.Tensor, w: W, p: str) -> torch.Tensor:
    """DecoderBlock::forward, autoencoder/mod.rs:306-324: three ResnetBlocks, then (if present) nearest-2x by
    reshape/repeat followed by a 3x3 conv."""
    for r in ("res1", "res2", "res3"):
        x = resnet_block(x, w, f"{p}/{r}")
    if f"{p}/upsampler/weight" in w:
        n_batch, n_channel, height, width = x.shape
        x = x.reshape(n_batch, n_channel, height, 1, width, 1).repeat(1, 1, 1, 2, 1, 2).reshape(
            n_batch, n_channel, 2 * height, 2 * width)
        x = conv2d(x, w, f"{p}/upsampler")
    return x


def decoder_forward(cfg, w: W, x: torch.Tensor) -> torch.Tensor:
    """Decoder::forward, autoencoder/mod.rs:202-216."""
    x = conv2d(x, w, "decoder/conv_in")
    x = mid(x, w, "decoder/mid")
    for i in range(len(cfg.block_channels)):
        x = decoder_block(x, w, f"decoder/blocks/{i}")
    x = silu(group_norm(x, w["decoder/norm_out/weight"], w["decoder/norm_out/bias"]))
    return conv2d(x, w, "decoder/conv_out")


def decode_latent(cfg, w: W, latent: torch.Tensor) -> torch.Tensor:
    """LatentDecoder::decode_latent (stablediffusion/mod.rs:263-266) over Autoencoder::decode_latent
    (autoencoder/mod.rs:66-69): post_quant_conv(latent * (1 / scale_factor)) -> Decoder."""
    x = latent * (1.0 / cfg.scale_factor)
    x = conv2d(x, w, "post_quant_conv", padding=0)
    return decoder_forward(cfg, w, x)


def latent_to_image(cfg, w: W, latent: torch.Tensor) -> torch.Tensor:
    """LatentDecoder::latent_to_image (stablediffusion/mod.rs:200-237): ((image + 1) / 2) to [B,H,W,3], x255,
    clamp to [0,255] in f64, truncate to u8. Returns u8 [B, H, W, 3]."""
    image = decode_latent(cfg, w, latent)
    image = (image + 1.0) / 2.0
    image = image.permute(0, 2, 3, 1) * 255.0
    return image.to(torch.float64).clamp(0.0, 255.0).to(torch.uint8)


def padded_conv2d(x: torch.Tensor, w: W, p: str, kernel_size: int = 3, stride: int = 2, pad=(0, 1, 0, 1)) -> torch.Tensor:
    """PaddedConv2d::{init, forward}, autoencoder/mod.rs:326-407 (pad = left, right, top, bottom): a symmetric conv with
    padding calc_padding(..) followed by a slice that drops the leading outputs."""
    pad_left, pad_right, pad_top, pad_bottom = pad

    def calc_padding(p_left, p_right):
        n = 0 if p_left >= p_right else (p_right - p_left + stride - 1) // stride
        return n * stride + p_left
    pv, ph = calc_padding(pad_top, pad_bottom), calc_padding(pad_left, pad_right)
    n_batch, n_channel, height, width = x.shape
    desired_h = (pad_top + pad_bottom + height - kernel_size) // stride + 1
    desired_w = (pad_left + pad_right + width - kernel_size) // stride + 1
    skip_v, skip_h = (pv - pad_top) // stride, (ph - pad_left) // stride
    import torch.nn.functional as F
    y = F.conv2d(x, w[f"{p}/conv/weight"], w.get(f"{p}/conv/bias"), stride=stride, padding=(pv, ph))
    return y[:, :, skip_v:skip_v + desired_h, skip_h:skip_h + desired_w]


def encoder_block(x: torch.Tensor, w: W, p: str) -> torch.Tensor:
    """EncoderBlock::forward, autoencoder/mod.rs:284-295."""
    x = resnet_block(x, w, f"{p}/res1")
    x = resnet_block(x, w, f"{p}/res2")
    if f"{p}/downsampler/conv/weight" in w:
        x = padded_conv2d(x, w, f"{p}/downsampler")
    return x


def encoder_forward(cfg, w: W, x: torch.Tensor) -> torch.Tensor:
    """Encoder::forward, autoencoder/mod.rs:128-144."""
    x = conv2d(x, w, "encoder/conv_in")
    for i in range(len(cfg.enc_block_channels)):
        x = encoder_block(x, w, f"encoder/blocks/{i}")
    x = mid(x, w, "encoder/mid")
    x = silu(group_norm(x, w["encoder/norm_out/weight"], w["encoder/norm_out/bias"]))
    return conv2d(x, w, "encoder/conv_out")


def encode_image(cfg, w: W, image: torch.Tensor) -> torch.Tensor:
    """LatentDecoder::encode_image (stablediffusion/mod.rs:258-261) over Autoencoder::encode_image (autoencoder/mod.rs:58-64):
    quant_conv(encoder(x))[:, 0:4] * scale_factor (the mean channels; the reference does not sample)."""
    latent = conv2d(encoder_forward(cfg, w, image), w, "quant_conv", padding=0)
    return latent[:, 0:cfg.latent_channels] * cfg.scale_factor


def image_to_latent(cfg, w: W, rgb_u8: torch.Tensor) -> torch.Tensor:
    """LatentDecoder::image_to_latent (stablediffusion/mod.rs:239-256): u8 [B,H,W,3] / 255 -> NCHW -> * 2 - 1 -> encode."""
    x = (rgb_u8.to(torch.float32) / 255.0).permute(0, 3, 1, 2) * 2.0 - 1.0
    return encode_image(cfg, w, x)


def encoder_flops(cfg, H: int, Wd: int, batch: int = 1) -> float:
    """Algorithmic FLOPs of encode_image at image H x Wd (quant_conv counted on all z channels as the reference computes it)."""
    c0, ce, cz = cfg.enc_block_channels[0][0], cfg.enc_block_channels[-1][1], cfg.enc_z_channels

    def res(hw, ci, co):
        return 2.0 * hw * 9 * (ci * co + co * co) + (2.0 * hw * ci * co if ci != co else 0.0)
    hw = H * Wd
    f = 2.0 * hw * 27 * c0
    for i, (ci, co) in enumerate(cfg.enc_block_channels):
        f += res(hw, ci, co) + res(hw, co, co)
        if i != len(cfg.enc_block_channels) - 1:
            hw //= 4
            f += 2.0 * hw * 9 * co * co
    f += 2 * res(hw, ce, ce) + 4 * 2.0 * hw * ce * ce + 2 * 2.0 * hw * hw * ce
    f += 2.0 * hw * 9 * ce * cz + 2.0 * hw * cz * cz
    return f * batch


def decoder_flops(cfg, h: int, wd: int, batch: int = 1) -> float:
    """Algorithmic FLOPs (2*MAC over conv / 1x1 / QK^T / PV) of decode_latent at latent h x wd."""
    cl, c0 = cfg.latent_channels, cfg.block_channels[0][0]
    f = 2.0 * h * wd * cl * cl + 2.0 * h * wd * 9 * cl * c0

    def res(hw, ci, co):
        return 2.0 * hw * 9 * (ci * co + co * co) + (2.0 * hw * ci * co if ci != co else 0.0)

    hw = h * wd
    f += 2 * res(hw, c0, c0)
    f += 4 * 2.0 * hw * c0 * c0 + 2 * 2.0 * hw * hw * c0
    for i, (ci, co) in enumerate(cfg.block_channels):
        f += res(hw, ci, co) + 2 * res(hw, co, co)
        if i != len(cfg.block_channels) - 1:
            hw *= 4
            f += 2.0 * hw * 9 * co * co
    f += 2.0 * hw * 9 * cfg.block_channels[-1][1] * 3
    return f * batch
