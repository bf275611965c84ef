"""The reference's shipped weight format (burn 0.13 NamedMpkFileRecorder<HalfPrecisionSettings> `.mpk` + `.cfg`, reference
src/bin/convert/main.rs:65-70, src/bin/sample/main.rs:28-51) and the `sample` front-end's inpainting mask
(src/bin/sample/main.rs:144-190). CPU only: the library's host-side helpers, no CUDA call."""
import json
import os

import msgpack
import numpy as np
import pytest
import torch

import sdxl_b200
from sdxl_b200 import TINY, TINY_REFINER, burn_record as BR, synth_weights
from oracle import unet_oracle as O


def test_value_array_codec_matches_msgpack():
    """The native u16 array codec against the `msgpack` package on every encoding class (fixint, uint8, uint16)."""
    import ctypes as C
    lib = sdxl_b200.load()
    vals = np.array([0, 1, 127, 128, 200, 255, 256, 0x3C00, 0xFFFF, 0x7BFF, 5] + list(np.random.default_rng(0).integers(0, 65536, 5000)), dtype=np.uint16)
    buf = np.empty(3 * vals.size, dtype=np.uint8)
    n = lib.sdxl_mpk_encode_u16(vals.ctypes.data, vals.size, buf.ctypes.data)
    ref = b"".join(msgpack.packb(int(v)) for v in vals)
    assert buf[:n].tobytes() == ref
    out = np.empty_like(vals)
    used = C.c_size_t(0)
    assert lib.sdxl_mpk_decode_u16(buf.ctypes.data, n, vals.size, out.ctypes.data, C.byref(used)) == 0
    assert used.value == n and np.array_equal(out, vals)
    assert lib.sdxl_mpk_decode_u16(buf.ctypes.data, n - 1, vals.size, out.ctypes.data, C.byref(used)) == 7001   # truncated
    bad = np.frombuffer(msgpack.packb(-3), dtype=np.uint8).copy()
    assert lib.sdxl_mpk_decode_u16(bad.ctypes.data, bad.size, 1, out.ctypes.data, C.byref(used)) == 7003          # not unsigned


@pytest.mark.parametrize("cfg,seed", [(TINY, 0), (TINY_REFINER, 1)])
def test_diffuser_record_round_trip(tmp_path, cfg, seed):
    """weights -> <name>.mpk + <name>.cfg -> weights: bit-identical tensors, identical config; the file is a MessagePack document
    the `msgpack` package reads, with the record layout burn's NamedMpk recorder writes (metadata + item, fields by name, enum
    blocks as one-entry maps, constants as nil, f16 bit patterns as unsigned integers)."""
    w = synth_weights(cfg, seed=seed)
    path = str(tmp_path / "diffuser")
    BR.save_diffuser(path, cfg, w)
    cfg2, w2 = BR.load_diffuser(path)
    assert cfg2 == cfg
    assert set(w2) == set(w)
    for k in w:
        assert w2[k].dtype == torch.float16 and torch.equal(w2[k], w[k]), k
    # independent reader: the msgpack package sees the same tree
    with open(path + ".mpk", "rb") as fh:
        doc = msgpack.unpackb(fh.read(), raw=False, strict_map_key=True)
    assert doc["metadata"]["float"] == "f16" and doc["metadata"]["version"].startswith("0.13")
    item = doc["item"]
    assert set(item) == {"n_steps", "alpha_cumulative_products", "diffusion", "is_refiner"} and item["n_steps"] is None
    blk0 = item["diffusion"]["input_blocks"][0]
    assert list(blk0) == ["Conv"] and blk0["Conv"]["weight"]["param"]["shape"] == [cfg.model_channels, cfg.in_channels, 3, 3]
    lin = item["diffusion"]["lin1_time_embed"]["weight"]["param"]
    assert lin["shape"] == [cfg.model_channels, 4 * cfg.model_channels]                       # burn Linear weight is [d_input, d_output]
    got = np.array(lin["value"], dtype=np.uint16).view(np.float16).reshape(lin["shape"])
    assert np.array_equal(got, w["lin1_time_embed/weight"].numpy())
    kinds = [list(b)[0] for b in item["diffusion"]["output_blocks"]]
    assert kinds[2] in ("ResTU", "ResU") and set(kinds) <= {"Res", "ResT", "ResTU", "ResU"}
    # the .cfg is the DiffuserConfig JSON (stablediffusion/mod.rs:269-278)
    d = json.load(open(path + ".cfg"))
    assert d["num_head_channels"] == 64 and d["is_refiner"] == cfg.is_refiner and d["channel_mults"] == list(cfg.channel_mults)


def test_record_reader_rejects_damage(tmp_path):
    w = synth_weights(TINY, seed=0)
    path = str(tmp_path / "d")
    BR.save_diffuser(path, TINY, w)
    raw = open(path + ".mpk", "rb").read()
    open(path + ".mpk", "wb").write(raw[: len(raw) // 2])
    with pytest.raises(BR.BurnRecordError):
        BR.load_diffuser(path)
    open(path + ".mpk", "wb").write(msgpack.packb({"hello": 1}))
    with pytest.raises(BR.BurnRecordError):
        BR.load_diffuser(path)
    os.remove(path + ".cfg")
    with pytest.raises(FileNotFoundError):
        BR.load_diffuser(path)


@pytest.mark.parametrize("img,crop,crop_out", [((1024, 1024), (None, None, None, 200), False), ((1024, 1024), (100, 731, 37, 999), False),
                                               ((1024, 1024), (100, 731, 37, 999), True), ((768, 1344), (8, 1344, 0, 768), False),
                                               ((1024, 1024), (0, 7, 0, 1024), False), ((1152, 896), (3, 893, 5, 1150), True)])
def test_inpaint_mask_matches_oracle(img, crop, crop_out):
    """sdxl_make_inpaint_mask against the restated reference (src/bin/sample/main.rs:144-190), bit for bit."""
    h, w = img
    lat = (h // 8, w // 8)
    ref = O.make_inpaint_mask(w, h, lat[1], lat[0], crop[0], crop[1], crop[2], crop[3], crop_out)
    got = sdxl_b200.make_inpaint_mask(img, lat, *crop, crop_out=crop_out)
    assert got.dtype == torch.bool and got.shape == (1, 4, lat[0], lat[1])
    assert torch.equal(got, ref)
    if crop == (None, None, None, 200) and not crop_out:
        assert got[0, 0, :25].all() and not got[0, 0, 25:].any()          # BASELINE config 5: rows 0..25 (200 px / 8)


def test_inpaint_mask_rejects_bad_windows():
    for crop in ((10, 5, 0, 100), (0, 2000, 0, 100), (0, 100, 50, 50)):
        with pytest.raises(sdxl_b200.SdxlError):
            sdxl_b200.make_inpaint_mask((1024, 1024), (128, 128), *crop)


def test_latent_decoder_and_embedder_records_round_trip(tmp_path):
    """The other two files `sample` loads (src/bin/sample/main.rs:28-33, 43-51): LatentDecoder and Embedder records + their .cfg."""
    from sdxl_b200 import TINY_CLIP, TINY_OPEN_CLIP, TINY_VAE
    wv = synth_weights(TINY_VAE, seed=5)
    BR.save_latent_decoder(str(tmp_path / "latent_decoder"), TINY_VAE, wv)
    cfg, w2 = BR.load_latent_decoder(str(tmp_path / "latent_decoder"))
    assert cfg == TINY_VAE                                  # widths are read off the tensors, scale_factor from the .cfg
    assert set(w2) == set(wv) and all(torch.equal(w2[k], wv[k]) for k in wv)
    wa, wb = synth_weights(TINY_CLIP, seed=6), synth_weights(TINY_OPEN_CLIP, seed=7)
    BR.save_embedder(str(tmp_path / "embedder"), TINY_CLIP, wa, TINY_OPEN_CLIP, wb)
    ca, ra, cb, rb = BR.load_embedder(str(tmp_path / "embedder"))
    assert (ca, cb) == (TINY_CLIP, TINY_OPEN_CLIP)
    for got, want in ((ra, wa), (rb, wb)):
        assert set(got) == set(want) and all(torch.equal(got[k], want[k]) for k in want)
    # a damaged config is an error, not a silent default
    (tmp_path / "embedder.cfg").write_text('{"clip_config": {"n_vocab": 1}}')
    with pytest.raises(BR.BurnRecordError):
        BR.load_embedder(str(tmp_path / "embedder"))


def test_read_model_dir_uses_the_sample_binarys_file_names(tmp_path):
    from sdxl_b200 import TINY_CLIP, TINY_OPEN_CLIP, TINY_VAE
    d = str(tmp_path)
    BR.save_embedder(os.path.join(d, "embedder"), TINY_CLIP, synth_weights(TINY_CLIP, seed=1), TINY_OPEN_CLIP, synth_weights(TINY_OPEN_CLIP, seed=2))
    BR.save_diffuser(os.path.join(d, "diffuser"), TINY, synth_weights(TINY, seed=3))
    BR.save_latent_decoder(os.path.join(d, "latent_decoder"), TINY_VAE, synth_weights(TINY_VAE, seed=4))
    files = BR.read_model_dir(d)
    assert files["refiner"] is None and files["diffuser"][0] == TINY and files["latent_decoder"][0] == TINY_VAE
    assert files["embedder"][0] == TINY_CLIP and files["embedder"][2] == TINY_OPEN_CLIP
    with pytest.raises(FileNotFoundError):
        BR.read_model_dir(d, use_refiner=True)             # <dir>/refiner.{cfg,mpk} absent
    BR.save_diffuser(os.path.join(d, "refiner"), TINY_REFINER, synth_weights(TINY_REFINER, seed=5))
    assert BR.read_model_dir(d, use_refiner=True)["refiner"][0] == TINY_REFINER
