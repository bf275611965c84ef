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


def _hf_text_model(cfg, w):
    """The same weights loaded into HuggingFace transformers' CLIPTextModelWithProjection (an independent implementation of the
    architecture the reference's dump scripts port from); Linear weights are [in, out] in the dump tree, [out, in] in PyTorch."""
    transformers = __import__("pytest").importorskip("transformers")
    hc = transformers.CLIPTextConfig(vocab_size=cfg.n_vocab, hidden_size=cfg.n_state, intermediate_size=4 * cfg.n_state,
                                     num_hidden_layers=cfg.n_layer, num_attention_heads=cfg.n_head, max_position_embeddings=cfg.n_ctx,
                                     hidden_act="quick_gelu" if cfg.quick_gelu else "gelu", projection_dim=cfg.embed_dim,
                                     eos_token_id=49407, bos_token_id=49406, pad_token_id=0, layer_norm_eps=1e-5, attention_dropout=0.0)
    m = transformers.CLIPTextModelWithProjection(hc).eval()
    sd = {"text_model.embeddings.token_embedding.weight": w["token_embedding/weight"],
          "text_model.embeddings.position_embedding.weight": w["position_embedding/weight"],
          "text_model.final_layer_norm.weight": w["layer_norm/weight"], "text_model.final_layer_norm.bias": w["layer_norm/bias"],
          "text_projection.weight": w["text_projection"].t().contiguous()}
    for i in range(cfg.n_layer):
        b, h = f"blocks/{i}", f"text_model.encoder.layers.{i}"
        for ours, theirs in (("attn/query", "self_attn.q_proj"), ("attn/key", "self_attn.k_proj"), ("attn/value", "self_attn.v_proj"),
                             ("attn/out", "self_attn.out_proj"), ("mlp/fc1", "mlp.fc1"), ("mlp/fc2", "mlp.fc2")):
            sd[f"{h}.{theirs}.weight"] = w[f"{b}/{ours}/weight"].t().contiguous()
            sd[f"{h}.{theirs}.bias"] = w[f"{b}/{ours}/bias"]
        for ours, theirs in (("attn_ln", "layer_norm1"), ("mlp_ln", "layer_norm2")):
            sd[f"{h}.{theirs}.weight"] = w[f"{b}/{ours}/weight"]
            sd[f"{h}.{theirs}.bias"] = w[f"{b}/{ours}/bias"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m


def test_oracle_vs_huggingface_clip_text_model():
    """Pins the text-encoder oracle against HuggingFace transformers' CLIP text tower on the same weights: the stream entering
    the last block (== hidden_states[n_layer-1], SDXL's "penultimate layer") and the projected end-of-text feature."""
    for cfg, seed, pad in ((TINY_CLIP, 1, 49407), (TINY_OPEN_CLIP, 2, 0)):
        w = O.to_f32(synth_weights(cfg, seed=seed))
        m = _hf_text_model(cfg, w)
        rows = [[49406, 320, 1125, 539, 320, 2368, 49407], [49406, 17, 4, 256, 300, 301, 302, 303, 9, 49407]]
        tok = torch.tensor([r + [pad] * (77 - len(r)) for r in rows])
        with torch.no_grad():
            out = m(input_ids=tok, output_hidden_states=True)
        h, pooled = CO.forward_hidden_pooled(cfg, w, tok, cfg.n_layer - 1)
        assert torch.allclose(h, out.hidden_states[cfg.n_layer - 1], atol=5e-5, rtol=1e-4)
        assert torch.allclose(CO.forward_hidden(cfg, w, tok, 1), out.hidden_states[1], atol=5e-5, rtol=1e-4)
        assert torch.allclose(pooled, out.text_embeds, atol=1e-4, rtol=1e-4)
