"""ORACLE — test infrastructure only. Never imported by the product path (sdxl_b200 / libsdxl_b200.so).

CPU f32 restatement (PyTorch tensor ops) of the reference's text encoder and Embedder glue, line by line from
/root/reference: src/model/clip/mod.rs (CLIP::forward_hidden / forward_hidden_pooled, ResidualDecoderAttentionBlock,
MultiHeadSelfAttention, MLP, QuickGELU), src/backend.rs:21,88-128 (attn_decoder_mask, generic qkv_attention) and
src/model/stablediffusion/mod.rs:654-776 (Embedder::text_to_conditioning and helpers).

    *** PARITY UNPINNED *** for the encoder numerics — the reference cannot be built here and ships no numeric golden for
    CLIP; goldens come from this file. (The tokenizers feeding it ARE pinned: oracle/tokenizer_oracle.py.)
tests/test_clip_oracle.py cross-checks the block against torch.nn.functional primitives (F.layer_norm,
F.scaled_dot_product_attention(is_causal=True), F.gelu) and the whole encoder against HuggingFace transformers'
CLIPTextModelWithProjection loaded with the same weights (hidden_states[n_layer-1] and text_embeds agree to 1e-4): the
architecture is pinned against its canonical implementation — the very model the reference's dump script reads its CLIP-L
weights from (python/clip.py:7-48 walks a HuggingFace CLIPTextModel); the reference's Rust port of it is what stays unverifiable.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from .unet_oracle import conditioning_embedding, gelu_erf, layer_norm, linear, qkv_attention

W = Dict[str, torch.Tensor]


def attn_decoder_mask(seq_length: int) -> torch.Tensor:
    """Backend::attn_decoder_mask (src/backend.rs:21, 130-140): -inf strictly above the diagonal, 0 elsewhere."""
    m = torch.zeros(seq_length, seq_length)
    return m.masked_fill(torch.ones(seq_length, seq_length, dtype=torch.bool).triu(1), float("-inf"))


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    """QuickGELU::forward, clip/mod.rs:316-318."""
    return x * torch.sigmoid(x * 1.702)


def mlp(x: torch.Tensor, w: W, p: str, quick: bool) -> torch.Tensor:
    """MLP::forward, clip/mod.rs:296-304."""
    x = linear(x, w, f"{p}/fc1")
    x = quick_gelu(x) if quick else gelu_erf(x)
    return linear(x, w, f"{p}/fc2")


def self_attention(x: torch.Tensor, mask: torch.Tensor, w: W, p: str, n_head: int) -> torch.Tensor:
    """MultiHeadSelfAttention::forward, clip/mod.rs:228-245."""
    q, k, v = linear(x, w, f"{p}/query"), linear(x, w, f"{p}/key"), linear(x, w, f"{p}/value")
    return linear(qkv_attention(q, k, v, mask, n_head), w, f"{p}/out")


def block(x: torch.Tensor, mask: torch.Tensor, w: W, p: str, n_head: int, quick: bool) -> torch.Tensor:
    """ResidualDecoderAttentionBlock::forward, clip/mod.rs:176-182."""
    x = x + self_attention(layer_norm(x, w[f"{p}/attn_ln/weight"], w[f"{p}/attn_ln/bias"]), mask, w, f"{p}/attn", n_head)
    return x + mlp(layer_norm(x, w[f"{p}/mlp_ln/weight"], w[f"{p}/mlp_ln/bias"]), w, f"{p}/mlp", quick)


def _embed(cfg, w: W, tokens: torch.Tensor) -> torch.Tensor:
    seq_len = tokens.shape[1]
    return w["token_embedding/weight"][tokens.long()] + w["position_embedding/weight"][:seq_len].unsqueeze(0)


def forward_hidden(cfg, w: W, tokens: torch.Tensor, hidden_idx: int) -> torch.Tensor:
    """CLIP::forward_hidden, clip/mod.rs:82-100."""
    mask = attn_decoder_mask(tokens.shape[1])
    x = _embed(cfg, w, tokens)
    for i in range(hidden_idx):
        x = block(x, mask, w, f"blocks/{i}", cfg.n_head, cfg.quick_gelu)
    return x


def forward_hidden_pooled(cfg, w: W, tokens: torch.Tensor, hidden_idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """CLIP::forward_hidden_pooled, clip/mod.rs:102-143."""
    mask = attn_decoder_mask(tokens.shape[1])
    x = _embed(cfg, w, tokens)
    h_out = None
    for i in range(cfg.n_layer):
        if i == hidden_idx:
            h_out = x.clone()
        x = block(x, mask, w, f"blocks/{i}", cfg.n_head, cfg.quick_gelu)
    eot_indices = tokens.argmax(1)
    normed = layer_norm(x.reshape(-1, x.shape[-1]), w["layer_norm/weight"], w["layer_norm/bias"]).reshape(x.shape)
    o = normed[torch.arange(tokens.shape[0]), eot_indices]
    pooled = o.matmul(w["text_projection"]) if "text_projection" in w else o
    return h_out, pooled


def text_to_conditioning(clip_cfg, clip_w: W, oc_cfg, oc_w: W, clip_tok, oc_tok, tokenize_text, text: str, size: Sequence[int],
                         crop: Sequence[int], ar: Sequence[int]) -> Dict[str, torch.Tensor]:
    """Embedder::text_to_conditioning, stablediffusion/mod.rs:654-759 (n_batch = 1). clip_tok / oc_tok are oracle
    tokenizers, tokenize_text is oracle.tokenizer_oracle.tokenize_text."""
    def context(t: str):
        t1 = torch.tensor([tokenize_text(t, clip_tok, clip_cfg.n_ctx)])
        clip_context = forward_hidden(clip_cfg, clip_w, t1, clip_cfg.n_layer - 1)
        t2 = torch.tensor([tokenize_text(t, oc_tok, oc_cfg.n_ctx)])
        oc_context, pooled = forward_hidden_pooled(oc_cfg, oc_w, t2, oc_cfg.n_layer - 1)
        sz, cr, a = torch.tensor([list(size)]), torch.tensor([list(crop)]), torch.tensor([list(ar)])
        aes = torch.tensor([[6]])
        return (torch.cat([clip_context, oc_context], 2), oc_context, conditioning_embedding(pooled, 256, sz, cr, a),
                conditioning_embedding(pooled, 256, sz, cr, aes))
    u = context("")
    c = context(text)
    return dict(context_full=c[0], context_open_clip=c[1], channel_context=c[2], channel_context_refiner=c[3],
                unconditional_context_full=u[0].squeeze(0), unconditional_context_open_clip=u[1].squeeze(0),
                unconditional_channel_context=u[2].squeeze(0), unconditional_channel_context_refiner=u[3].squeeze(0))
