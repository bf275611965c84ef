"""Weight-format reader: the reference's npy dump tree (python/save.py:10-16, src/model/load.rs:15-62) -> flat pack."""
import os

import numpy as np
import pytest
import torch

from sdxl_b200 import TINY, TINY_CLIP, TINY_VAE, build_pack, synth_weights
from sdxl_b200.convert import NpyTreeError, load_npy_tree, pack_from_npy_tree, read_scalar, read_tensor, write_npy_tree, write_tensor


def test_file_format_is_shape_prefixed_f32(tmp_path):
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    write_tensor(str(tmp_path), "a/b/weight", t)
    raw = np.load(tmp_path / "a" / "b" / "weight.npy")
    assert raw.dtype == np.float32 and raw.ndim == 1 and raw[:3].tolist() == [2, 3, 4] and raw.size == 27   # save.py:10-16
    assert np.array_equal(read_tensor(str(tmp_path), "a/b/weight", 3), t.numpy())
    np.save(tmp_path / "n_steps.npy", np.array([1.0, 1000.0], dtype=np.float32))                          # save_scalar
    assert read_scalar(str(tmp_path), "n_steps") == 1000.0


@pytest.mark.parametrize("cfg,seed", [(TINY, 0), (TINY_VAE, 3), (TINY_CLIP, 1)])
def test_tree_roundtrip_gives_identical_pack(tmp_path, cfg, seed):
    w = synth_weights(cfg, seed=seed)
    root = str(tmp_path / "params" / "model")
    write_npy_tree(w, root)
    back = load_npy_tree(root, cfg)
    assert set(back) == set(w)
    for k in w:
        assert back[k].dtype == torch.float16 and torch.equal(back[k], w[k]), k
    assert torch.equal(pack_from_npy_tree(root, cfg), build_pack(w))


def test_errors(tmp_path):
    w = synth_weights(TINY_VAE, seed=3)
    root = str(tmp_path / "vae")
    write_npy_tree(w, root)
    os.remove(os.path.join(root, "decoder", "conv_in", "bias.npy"))
    with pytest.raises(NpyTreeError, match="missing tensor file"):
        load_npy_tree(root, TINY_VAE)
    write_tensor(root, "decoder/conv_in/bias", torch.zeros(7))
    with pytest.raises(NpyTreeError, match="config needs"):
        load_npy_tree(root, TINY_VAE)
    np.save(os.path.join(root, "decoder", "conv_in", "bias.npy"), np.array([5.0, 1.0, 2.0], dtype=np.float32))
    with pytest.raises(NpyTreeError, match="shape prefix"):
        load_npy_tree(root, TINY_VAE)


def test_convert_cli_writes_the_same_pack(tmp_path, monkeypatch):
    from sdxl_b200 import convert, config
    monkeypatch.setattr(config, "SDXL_VAE", TINY_VAE)   # the CLI's model table, shrunk for the test
    w = synth_weights(TINY_VAE, seed=3)
    root = str(tmp_path / "vae")
    write_npy_tree(w, root)
    out = str(tmp_path / "vae.pack")
    assert convert._main(["vae", root, out]) == 0
    assert open(out, "rb").read() == build_pack(w).numpy().tobytes()


def test_convert_cli_burn_record_both_ways(tmp_path, monkeypatch):
    """npy tree -> (--to-mpk) <stem>.mpk + .cfg, as the reference's convert binary does; then --from-mpk -> the same pack."""
    from sdxl_b200 import TINY, convert, config
    monkeypatch.setattr(config, "SDXL_BASE", TINY)
    w = synth_weights(TINY, seed=4)
    root, stem = str(tmp_path / "diffuser"), str(tmp_path / "rec" / "tiny")
    os.makedirs(os.path.dirname(stem))
    write_npy_tree(w, root)
    out1, out2 = str(tmp_path / "a.pack"), str(tmp_path / "b.pack")
    assert convert._main(["unet_base", root, out1, "--to-mpk", stem]) == 0
    assert os.path.exists(stem + ".mpk") and os.path.exists(stem + ".cfg")
    assert convert._main(["unet_base", "-", out2, "--from-mpk", stem]) == 0
    assert open(out1, "rb").read() == open(out2, "rb").read() == build_pack(w).numpy().tobytes()
    with pytest.raises(SystemExit):
        convert._main(["vae", root, out1, "--to-mpk", stem])
