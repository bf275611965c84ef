"""CPU checks of the text-encoder oracle (oracle/clip_oracle.py) against independent PyTorch implementations of the same
published ops, and of the committed fixture."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import clip_oracle as CO
from oracle import unet_oracle as O
from sdxl_b200 import SDXL_CLIP_L, SDXL_OPEN_CLIP_G, TINY_CLIP, TINY_OPEN_CLIP, clip_tensor_specs, synth_weights

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _torch_block(x, w, p, n_head, quick):
    """The same block written with torch.nn.functional (PyTorch Linear stores [out, in])."""
    def lin(t, q):
        return F.linear(t, w[f"{q}/weight"].t(), w[f"{q}/bias"])
    B, T, C = x.shape
    h = F.layer_norm(x, (C,), w[f"{p}/attn_ln/weight"], w[f"{p}/attn_ln/bias"], eps=1e-5)
    q, k, v = (lin(h, f"{p}/attn/{n}").reshape(B, T, n_head, C // n_head).transpose(1, 2) for n in ("query", "key", "value"))
    a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, C)
    x = x + lin(a, f"{p}/attn/out")
    h = lin(F.layer_norm(x, (C,), w[f"{p}/mlp_ln/weight"], w[f"{p}/mlp_ln/bias"], eps=1e-5), f"{p}/mlp/fc1")
    h = h * torch.sigmoid(1.702 * h) if quick else F.gelu(h)
    return x + lin(h, f"{p}/mlp/fc2")


def test_block_vs_torch_functional():
    for cfg, seed in ((TINY_CLIP, 1), (TINY_OPEN_CLIP, 2)):
        w = O.to_f32(synth_weights(cfg, seed=seed))
        x = torch.randn(2, 77, cfg.n_state, generator=torch.Generator().manual_seed(0))
        got = CO.block(x, CO.attn_decoder_mask(77), w, "blocks/1", cfg.n_head, cfg.quick_gelu)
        want = _torch_block(x, w, "blocks/1", cfg.n_head, cfg.quick_gelu)
        assert torch.allclose(got, want, atol=3e-5, rtol=1e-5)


def test_forward_hidden_pooled_structure():
    cfg = TINY_OPEN_CLIP
    w = O.to_f32(synth_weights(cfg, seed=2))
    tok = torch.tensor([[49406, 7, 8, 49407] + [0] * 73, [49406, 9, 49407] + [0] * 74])
    h, pooled = CO.forward_hidden_pooled(cfg, w, tok, cfg.n_layer - 1)
    assert torch.equal(h, CO.forward_hidden(cfg, w, tok, cfg.n_layer - 1))   # h_out == stream entering the last block
    x = CO.block(h, CO.attn_decoder_mask(77), w, f"blocks/{cfg.n_layer - 1}", cfg.n_head, cfg.quick_gelu)
    n = F.layer_norm(x, (cfg.n_state,), w["layer_norm/weight"], w["layer_norm/bias"], eps=1e-5)
    want = torch.stack([n[0, 3], n[1, 2]]) @ w["text_projection"]               # end-of-text positions (argmax of ids)
    assert torch.allclose(pooled, want, atol=2e-5, rtol=1e-5)
    # causal: changing a later token must not change earlier positions
    tok2 = tok.clone()
    tok2[0, 3] = 11
    assert torch.equal(CO.forward_hidden(cfg, w, tok2, 2)[0, :3], CO.forward_hidden(cfg, w, tok, 2)[0, :3])


def test_param_counts():
    # CLIP ViT-L/14 text tower 123.06 M (+0.59 M text_projection); OpenCLIP bigG text tower 694.7 M
    n1 = sum(int(np.prod(s[1])) for s in clip_tensor_specs(SDXL_CLIP_L))
    n2 = sum(int(np.prod(s[1])) for s in clip_tensor_specs(SDXL_OPEN_CLIP_G))
    assert n1 == 123_650_304 and n2 == 694_659_840


def test_clip_golden_reproduces():
    g = np.load(os.path.join(GOLD, "tiny_clip.npz"))
    w1, w2 = O.to_f32(synth_weights(TINY_CLIP, seed=1)), O.to_f32(synth_weights(TINY_OPEN_CLIP, seed=2))
    h1 = CO.forward_hidden(TINY_CLIP, w1, torch.from_numpy(g["tokens_clip"]), TINY_CLIP.n_layer - 1)
    h2, p2 = CO.forward_hidden_pooled(TINY_OPEN_CLIP, w2, torch.from_numpy(g["tokens_open_clip"]), TINY_OPEN_CLIP.n_layer - 1)
    assert np.allclose(h1.numpy(), g["hidden_clip"], atol=2e-5, rtol=1e-5)
    assert np.allclose(h2.numpy(), g["hidden_open_clip"], atol=2e-5, rtol=1e-5)
    assert np.allclose(p2.numpy(), g["pooled_open_clip"], atol=2e-5, rtol=1e-5)
