"""The `sample` binary's flow (reference src/bin/sample/main.rs:128-285) over the library: crop window -> inpainting mask,
text -> conditioning -> base sampling (plain or inpainting) -> optional refiner hand-off -> latent -> image."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

REFINER_STEP_START = 800   # main.rs:263: the refiner re-noises to t0 = 1000 - 800 and runs the remaining iterations


def make_inpaint_mask(img_hw: Tuple[int, int], latent_hw: Tuple[int, int], crop_left: Optional[int] = None, crop_right: Optional[int] = None,
                      crop_top: Optional[int] = None, crop_bottom: Optional[int] = None, crop_out: bool = False, n_channels: int = 4) -> torch.Tensor:
    """== main.rs:144-190 (sdxl_make_inpaint_mask). Returns Bool [1, n_channels, h, w]; True keeps the generated latent."""
    lib = _lib.load()
    (ih, iw), (lh, lw) = img_hw, latent_hw
    out = np.empty((n_channels, lh, lw), dtype=np.uint8)
    opt = lambda v: -1 if v is None else int(v)  # noqa: E731
    rc = lib.sdxl_make_inpaint_mask(iw, ih, lw, lh, opt(crop_left), opt(crop_right), opt(crop_top), opt(crop_bottom), int(crop_out), n_channels, out.ctypes.data)
    if rc != 0:
        raise _lib.SdxlError(f"sdxl_make_inpaint_mask failed with {rc}: invalid crop parameters")
    return torch.from_numpy(out).bool().unsqueeze(0)


def sample(embedder, diffuser, decoder, prompt: str, guidance: float = 7.5, n_steps: int = 30, refiner=None,
           reference_rgb: Optional[torch.Tensor] = None, crop: Sequence[Optional[int]] = (None, None, None, None), crop_out: bool = False,
           resolution: Tuple[int, int] = (1024, 1024), seed: int = 0, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One image, like `sample --prompt ... [--reference-img ... --crop-* ...] [--use-refiner]`.
    reference_rgb: uint8 [1, H, W, 3] (the reference image: switches to inpainting, main.rs:131-197); crop = (left, right, top,
    bottom) in pixels. Returns uint8 [1, H, W, 3]."""
    if reference_rgb is not None:
        resolution = (int(reference_rgb.shape[1]), int(reference_rgb.shape[2]))        # main.rs:225-229: orig_dims
    size = [int(resolution[0]), int(resolution[1])]
    cond = embedder.text_to_conditioning(prompt, size, [0, 0], size)                     # main.rs:231-235: size, crop = 0, ar = size
    if reference_rgb is not None:
        ref_latent = decoder.image_to_latent(reference_rgb)                              # main.rs:158
        lh, lw = int(ref_latent.shape[2]), int(ref_latent.shape[3])
        cond.resolution = (8 * lh, 8 * lw)   # the sampler's latent extent follows the encoded reference (== the image size for the x8 SDXL VAE)
        mask = make_inpaint_mask(resolution, (lh, lw), *crop, crop_out=crop_out)
        latent = diffuser.sample_latent_with_inpainting(cond, guidance, n_steps, ref_latent, mask, init_noise=noise, seed=seed)   # main.rs:246
    else:
        latent = diffuser.sample_latent(cond, guidance, n_steps, noise=noise, seed=seed)                                      # main.rs:249
    if refiner is not None:
        latent = refiner.refine_latent(latent, cond, guidance, REFINER_STEP_START, n_steps, seed=seed + 1)                    # main.rs:258-265
    return decoder.latent_to_image(latent)                                                                                   # main.rs:277


def load_models(ctx, model_dir: str, use_refiner: bool = False, tokenizer_dir: str = "tokenizer", tokenizers=None):
    """The four loads of the reference's `sample` (src/bin/sample/main.rs:156, 220, 242, 255, 274): `<model_dir>/embedder`,
    `/diffuser`, `/refiner` (optional) and `/latent_decoder`, each a burn record `<name>.mpk` + `<name>.cfg`
    (burn_record.read_model_dir); the tokenizers read `<tokenizer_dir>/clip/bpe_simple_vocab_16e6.txt` and
    `<tokenizer_dir>/open_clip/{merges,vocab}.txt` like the reference (src/token/clip.rs:97, open_clip.rs:88-89) unless a
    (clip, open_clip) pair is passed. Returns (embedder, diffuser, refiner | None, decoder), ready for `sample()`."""
    import os
    from . import burn_record as BR
    from .embedder import ClipTextEncoder, Embedder
    from .engine import Diffuser, LatentDecoder
    from .tokenizer import ClipTokenizer, OpenClipTokenizer
    files = BR.read_model_dir(model_dir, use_refiner)
    if tokenizers is None:
        tokenizers = (ClipTokenizer(os.path.join(tokenizer_dir, "clip", "bpe_simple_vocab_16e6.txt")),
                      OpenClipTokenizer(os.path.join(tokenizer_dir, "open_clip", "merges.txt"), os.path.join(tokenizer_dir, "open_clip", "vocab.txt")))
    ca, wa, cb, wb = files["embedder"]
    emb = Embedder(ctx, ClipTextEncoder(ctx, ca, wa), ClipTextEncoder(ctx, cb, wb), tokenizers[0], tokenizers[1])
    dif = Diffuser(ctx, *files["diffuser"])
    ref = Diffuser(ctx, *files["refiner"]) if files["refiner"] is not None else None
    dec = LatentDecoder(ctx, *files["latent_decoder"])
    return emb, dif, ref, dec
