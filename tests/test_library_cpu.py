"""CPU tests of the boundary: the C-ABI library loads without a GPU and exports every symbol the header
declares; the product path refuses to run without CUDA (no CPU fallback); the weight pack round-trips."""
import ctypes
import os
import re
import struct

import pytest
import torch

import sdxl_b200
from sdxl_b200 import _lib
from sdxl_b200.config import SDXL_BASE, TINY, block_program
from sdxl_b200.weights import build_pack, synth_weights, unet_tensor_specs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sdxl_b200.h")).read()
    return sorted(set(re.findall(r"SDXL_API\s+[\w\s\*]+?\b(sdxl_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sdxl_b200.h but not exported"
    assert set(syms) == set(_lib.PROTOTYPES), "python prototypes and header disagree"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.sdxl_ctx_create(0, None, ctypes.byref(h)) != 0 and not h.value
    with pytest.raises(sdxl_b200.SdxlError):
        sdxl_b200.Context(0)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsdxl_b200.so")
    with pytest.raises(sdxl_b200.SdxlLibraryMissing):
        _lib.load()


def test_pack_roundtrip():
    w = synth_weights(TINY, seed=0)
    pack = build_pack(w).numpy().tobytes()
    magic, n, _, data_off = struct.unpack_from("<8sIIQ", pack, 0)
    assert magic == b"SDXLPK01" and n == len(w)
    names = []
    for i in range(n):
        name, dtype, ndim, s0, s1, s2, s3, off, nbytes = struct.unpack_from("<120sII4QQQ", pack, 24 + 176 * i)
        name = name.rstrip(b"\0").decode()
        t = w[name]
        assert dtype == 0 and ndim == t.dim() and off % 256 == 0 and off >= data_off
        assert [s0, s1, s2, s3][:ndim] == list(t.shape) and nbytes == t.numel() * 2
        assert pack[off:off + nbytes] == t.contiguous().numpy().tobytes()
        names.append(name)
    assert names == [s[0] for s in unet_tensor_specs(TINY)] + ["alphas_cumprod"]


def test_block_program_base():
    ins, mid, outs = block_program(SDXL_BASE)
    assert [(b.c_in, b.c_out) for b in outs] == [(2560, 1280), (2560, 1280), (1920, 1280), (1920, 640), (1280, 640), (960, 640),
                                                 (960, 320), (640, 320), (640, 320)]
    assert mid.depth == 10 and mid.n_head == 20
    assert [b.kind for b in ins].count("downsample") == 2


def test_header_is_plain_c_and_links(tmp_path):
    """A C99 program including include/sdxl_b200.h compiles with gcc, links against libsdxl_b200.so (every referenced entry
    point resolves) and drives the CPU-only tokenizer entry points — the binding a cgo / Rust `extern "C"` shim would make."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    lib_dir = os.path.join(root, "stable-diffusion-xl-burn_b200", "sdxl_b200")
    exe = str(tmp_path / "abi_check")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "c_abi", "abi_check.c"), "-L", lib_dir, "-lsdxl_b200", "-Wl,-rpath," + lib_dir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mini = os.path.join(root, "tests", "golden", "mini_bpe")
    r = subprocess.run([exe, os.path.join(mini, "mini_merges.txt"), os.path.join(mini, "mini_vocab.txt")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith("abi_check ok")
