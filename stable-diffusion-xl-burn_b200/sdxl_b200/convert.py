"""Weight-format reader: the reference's npy dump tree -> the flat pack sdxl_*_load takes.

The reference's Python dump scripts (python/save.py:10-16) store every tensor as a 1-D float32 .npy whose first `ndim` values
are the shape and the rest the flattened data (scalars are `[1.0, value]`); src/model/load.rs:15-44 reads them back with the
rank known from context. File paths are `<root>/<tensor name>.npy` with exactly the names this package's `*_tensor_specs`
list (they were taken from the reference's loaders). `convert` in the reference (src/bin/convert/main.rs) turns that tree
into burn's `.mpk`; here it is turned into the pack, stored f16 like the shipped half-precision records.

The burn `.mpk` (NamedMpk + HalfPrecisionSettings) reader is not built: no sample file exists offline to validate against.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from .config import ClipConfig, UNetConfig, VaeConfig
from .weights import build_pack, clip_tensor_specs, unet_tensor_specs, vae_tensor_specs


class NpyTreeError(RuntimeError):
    pass


def read_tensor(root: str, name: str, ndim: int) -> np.ndarray:
    """== load_tensor::<B, D> (src/model/load.rs:27-44) + numpy_to_tensor (:15-25)."""
    path = os.path.join(root, name + ".npy")
    if not os.path.exists(path):
        raise NpyTreeError(f"missing tensor file {path}")
    v = np.load(path)
    if v.dtype != np.float32 or v.ndim != 1 or v.size < ndim:
        raise NpyTreeError(f"{path}: expected a 1-D float32 array with a {ndim}-value shape prefix")
    shape = [int(x) for x in v[:ndim]]
    if any(s < 0 or float(s) != float(x) for s, x in zip(shape, v[:ndim])) or int(np.prod(shape)) != v.size - ndim:
        raise NpyTreeError(f"{path}: shape prefix {v[:ndim].tolist()} does not match {v.size - ndim} values")
    return v[ndim:].reshape(shape)


def read_scalar(root: str, name: str) -> float:
    """== load_f32 / load_usize (src/model/load.rs:46-62): stored as [1.0, value]."""
    return float(read_tensor(root, name, 1)[0])


def write_tensor(root: str, name: str, t) -> None:
    """== save_tensor (python/save.py:10-16); used by tests and by anyone re-dumping weights."""
    a = np.asarray(t.detach().cpu().float().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)
    path = os.path.join(root, name + ".npy")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.save(path, np.concatenate((np.array(a.shape, dtype=np.float32), a.flatten())).astype(np.float32))


def _specs(cfg) -> Iterable[Tuple[str, Tuple[int, ...]]]:
    if isinstance(cfg, UNetConfig):
        return [(s[0], s[1]) for s in unet_tensor_specs(cfg)]
    if isinstance(cfg, VaeConfig):
        return [(s[0], s[1]) for s in vae_tensor_specs(cfg)]
    if isinstance(cfg, ClipConfig):
        return [(s[0], s[1]) for s in clip_tensor_specs(cfg)]
    raise TypeError(f"unsupported config {type(cfg).__name__}")


def load_npy_tree(root: str, cfg, alphas_root: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Reads every tensor the model needs from `<root>/<name>.npy`, checks shapes against the config, returns f16 tensors.
    UNet: `alphas_cumprod.npy` is looked up in alphas_root (default: the parent of root, where load_diffuser reads it,
    src/model/stablediffusion/load.rs:56-57)."""
    out: Dict[str, torch.Tensor] = {}
    for name, shape in _specs(cfg):
        if name == "text_projection" and not os.path.exists(os.path.join(root, name + ".npy")):
            continue  # optional (clip/load.rs:102-104)
        a = read_tensor(root, name, len(shape))
        if tuple(a.shape) != tuple(shape):
            raise NpyTreeError(f"{os.path.join(root, name)}.npy has shape {tuple(a.shape)}, the config needs {tuple(shape)}")
        out[name] = torch.from_numpy(np.ascontiguousarray(a)).to(torch.float16)
    if isinstance(cfg, UNetConfig):
        ar = alphas_root if alphas_root is not None else os.path.dirname(os.path.abspath(root))
        a = read_tensor(ar, "alphas_cumprod", 1)
        if a.shape[0] != cfg.n_steps:
            raise NpyTreeError(f"alphas_cumprod has {a.shape[0]} entries, expected {cfg.n_steps}")
        out["alphas_cumprod"] = torch.from_numpy(np.ascontiguousarray(a)).to(torch.float16)
    return out


def pack_from_npy_tree(root: str, cfg, alphas_root: Optional[str] = None) -> torch.Tensor:
    """npy dump tree -> flat pack (uint8 tensor) for sdxl_unet_load / sdxl_vae_load / sdxl_clip_load."""
    return build_pack(load_npy_tree(root, cfg, alphas_root))


def write_npy_tree(weights: Dict[str, torch.Tensor], root: str, alphas_root: Optional[str] = None) -> None:
    """Inverse of load_npy_tree (the layout python/save.py produces)."""
    for name, t in weights.items():
        if name == "alphas_cumprod":
            write_tensor(alphas_root if alphas_root is not None else os.path.dirname(os.path.abspath(root)), name, t)
        else:
            write_tensor(root, name, t)


def _main(argv=None) -> int:
    """`python -m sdxl_b200.convert <model> <npy tree root> <out.pack> [--to-mpk STEM | --from-mpk STEM]` — the role of the reference's
    `convert` binary (src/bin/convert/main.rs) for this library's pack format, plus both directions of its burn record. model: unet_base | unet_refiner | vae | clip_l | open_clip_g."""
    import argparse
    from .config import SDXL_BASE, SDXL_CLIP_L, SDXL_OPEN_CLIP_G, SDXL_REFINER, SDXL_VAE
    models = {"unet_base": SDXL_BASE, "unet_refiner": SDXL_REFINER, "vae": SDXL_VAE, "clip_l": SDXL_CLIP_L, "open_clip_g": SDXL_OPEN_CLIP_G}
    ap = argparse.ArgumentParser(prog="python -m sdxl_b200.convert", description=_main.__doc__)
    ap.add_argument("model", choices=sorted(models))
    ap.add_argument("root", help="directory of the npy dump tree for this model (e.g. params/diffuser_base)")
    ap.add_argument("out", help="output pack file")
    ap.add_argument("--alphas-root", default=None, help="directory holding alphas_cumprod.npy (UNet only; default: parent of root)")
    ap.add_argument("--from-mpk", metavar="STEM", default=None,
                    help="UNet models: read the burn record <STEM>.mpk + <STEM>.cfg (the files the reference ships and loads, "
                         "src/bin/sample/main.rs:28-51) instead of an npy tree; pass '-' as root")
    ap.add_argument("--to-mpk", metavar="STEM", default=None,
                    help="UNet models: also write <STEM>.mpk + <STEM>.cfg, i.e. what the reference's convert binary produces "
                         "(src/bin/convert/main.rs:65-70)")
    a = ap.parse_args(argv)
    cfg = models[a.model]
    is_unet = a.model.startswith("unet")
    if (a.from_mpk or a.to_mpk) and not is_unet:
        ap.error("--from-mpk / --to-mpk apply to the UNet (Diffuser) models")
    if a.from_mpk:
        from . import burn_record
        from .weights import build_pack
        rec_cfg, weights = burn_record.load_diffuser(a.from_mpk)
        if rec_cfg != cfg:
            ap.error(f"{a.from_mpk}.cfg describes {rec_cfg}, not {a.model}")
        order = [name for name, _ in _specs(cfg)]   # the npy tree's order: the pack is then byte-identical to the one built from the tree
        weights = {**{n: weights[n] for n in order if n in weights}, **{n: t for n, t in weights.items() if n not in order}}
        pack = build_pack(weights)
    else:
        weights = load_npy_tree(a.root, cfg, a.alphas_root) if a.to_mpk else None
        pack = pack_from_npy_tree(a.root, cfg, a.alphas_root)
    if a.to_mpk:
        from . import burn_record
        burn_record.save_diffuser(a.to_mpk, cfg, weights)
        print(f"wrote {a.to_mpk}.mpk, {a.to_mpk}.cfg")
    with open(a.out, "wb") as f:
        f.write(pack.numpy().tobytes())
    print(f"wrote {a.out}: {pack.numel()} bytes")
    return 0


if __name__ == "__main__":
    raise SystemExit(_main())
