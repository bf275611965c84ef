"""Host-side mirror of the reference's `Tokenizer` trait (src/token/mod.rs) over the C ABI tokenizers in
libsdxl_b200.so (csrc/tokenizer.cpp). No oracle, no Python BPE here: every call goes through the library."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

from . import _lib
from ._lib import SdxlError


class _Tokenizer:
    def __init__(self, handle):
        self._lib = _lib.load()
        self.h = handle

    @staticmethod
    def _check(lib, rc: int, what: str) -> None:
        if rc != 0:
            raise SdxlError(f"{what} failed ({rc}): {lib.sdxl_tokenizer_last_error().decode(errors='replace')}")

    def close(self) -> None:
        if getattr(self, "h", None):
            self._lib.sdxl_tokenizer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, text: str, add_sot: bool, add_eot: bool) -> List[int]:
        """== Tokenizer::encode (src/token/clip.rs:182-205)."""
        raw = text.encode("utf-8")
        if b"\0" in raw:
            raise ValueError("text must not contain NUL")
        n = C.c_int(0)
        self._check(self._lib, self._lib.sdxl_tokenizer_encode(self.h, raw, int(add_sot), int(add_eot), None, 0, C.byref(n)),
                    "sdxl_tokenizer_encode")
        buf = (C.c_uint32 * max(1, n.value))()
        self._check(self._lib, self._lib.sdxl_tokenizer_encode(self.h, raw, int(add_sot), int(add_eot), buf, n.value, C.byref(n)),
                    "sdxl_tokenizer_encode")
        return list(buf[:n.value])

    def decode(self, tokens: Sequence[int]) -> str:
        """== Tokenizer::decode (src/token/clip.rs:207-213)."""
        ids = (C.c_uint32 * max(1, len(tokens)))(*tokens)
        n = C.c_int(0)
        self._check(self._lib, self._lib.sdxl_tokenizer_decode(self.h, ids, len(tokens), None, 0, C.byref(n)), "sdxl_tokenizer_decode")
        buf = C.create_string_buffer(n.value + 1)
        self._check(self._lib, self._lib.sdxl_tokenizer_decode(self.h, ids, len(tokens), buf, n.value + 1, C.byref(n)),
                    "sdxl_tokenizer_decode")
        return buf.raw[:n.value].decode("utf-8", errors="replace")

    def tokenize_text(self, text: str, seq_len: int = 77) -> List[int]:
        """== tokenize_text (src/model/stablediffusion/mod.rs:778-793)."""
        out = (C.c_int32 * max(1, seq_len))()
        self._check(self._lib, self._lib.sdxl_tokenize_text(self.h, text.encode("utf-8"), seq_len, out), "sdxl_tokenize_text")
        return list(out[:seq_len])

    def _special(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._lib, self._lib.sdxl_tokenizer_special(self.h, C.byref(a), C.byref(b), C.byref(c)), "sdxl_tokenizer_special")
        return a.value, b.value, c.value

    def start_of_text_token(self) -> int:
        return self._special()[0]

    def end_of_text_token(self) -> int:
        return self._special()[1]

    def padding_token(self) -> int:
        return self._special()[2]


class ClipTokenizer(_Tokenizer):
    """== ClipTokenizer::new (src/token/clip.rs:91-122); the reference reads tokenizer/clip/bpe_simple_vocab_16e6.txt."""

    def __init__(self, merges_path: str = os.path.join("tokenizer", "clip", "bpe_simple_vocab_16e6.txt")):
        lib = _lib.load()
        h = C.c_void_p()
        self._check(lib, lib.sdxl_tokenizer_create_clip(merges_path.encode(), C.byref(h)), "sdxl_tokenizer_create_clip")
        super().__init__(h)


class OpenClipTokenizer(_Tokenizer):
    """== OpenClipTokenizer::new (src/token/open_clip.rs:82-113); tokenizer/open_clip/{merges,vocab}.txt."""

    def __init__(self, merges_path: str = os.path.join("tokenizer", "open_clip", "merges.txt"),
                 vocab_path: str = os.path.join("tokenizer", "open_clip", "vocab.txt")):
        lib = _lib.load()
        h = C.c_void_p()
        self._check(lib, lib.sdxl_tokenizer_create_open_clip(merges_path.encode(), vocab_path.encode(), C.byref(h)),
                    "sdxl_tokenizer_create_open_clip")
        super().__init__(h)
