"""ORACLE — test infrastructure only. Never imported by the product path (sdxl_b200 / libsdxl_b200.so).

CPU f32 restatement (PyTorch tensor ops) of the reference's latent decoder, written line-by-line from
/root/reference (Gadersd/stable-diffusion-xl-burn): src/model/autoencoder/mod.rs (Autoencoder::decode_latent,
Decoder, Mid, ResnetBlock, ConvSelfAttentionBlock, DecoderBlock) and src/model/stablediffusion/mod.rs:199-237,
263-266 (LatentDecoder::{decode_latent, latent_to_image}).

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
