"""ORACLE — test infrastructure only. Never imported by the product path (sdxl_b200 / libsdxl_b200.so).

Pure-Python restatement of the reference's two BPE tokenizers, line by line:
  ClipTokenizer      /root/reference/src/token/clip.rs
  OpenClipTokenizer  /root/reference/src/token/open_clip.rs
  tokenize_text      /root/reference/src/model/stablediffusion/mod.rs:778-793

PARITY PINNED: tests/test_tokenizer.py checks this file against the reference's own known-answer vector
(src/token/clip.rs:232-249: "Hello world! <|startoftext|>asdf<|startoftext|>" -> [3306, 1002, 256, 49406, 587, 10468,
49406], decode -> "hello world ! <|startoftext|>asdf <|startoftext|>") using the reference's vocabulary file. It is the
only numeric golden the reference ships.

Rust semantics restated with their Python equivalents: `regex` crate pattern -> the `regex` module with the same
pattern text (Unicode classes, (?i), leftmost-first alternation); `str::to_lowercase` -> `str.lower()` (both apply the
full Unicode lower-case mapping incl. the final-sigma rule); `split_whitespace`/`trim` -> split on \\p{White_Space};
`HashMap: FromIterator` -> dict (later duplicates win); `String::from_utf8_lossy` -> bytes.decode(errors="replace").
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import regex

# The reference's pattern (clip.rs:111) is
#   (?i)<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|\p{L}+|\p{N}|[^\s\p{L}\p{N}]+
# Under (?i) Rust's regex closes every class under simple case folding; for \p{L} that adds exactly U+0345 (Mn, folds to the
# letter U+03B9) and nothing for \p{N} / \s (scan in tests/test_tokenizer.py::test_case_fold_closure_of_letter_class).
# Python's `regex` does not fold property classes (it already keeps U+0345 out of the negated class), so the letter
# alternative spells the closure out.
PAT = r"(?i)<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}\u0345]+|\p{N}|[^\s\p{L}\p{N}]+"
_WS = regex.compile(r"\p{White_Space}+")


def bytes_to_unicode() -> List[Tuple[int, str]]:
    """clip.rs:11-32"""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    cs = [chr(b) for b in bs]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(chr(256 + n))
            n += 1
    return list(zip(bs, cs))


def get_pairs(word: List[str]) -> List[Tuple[str, str]]:
    """clip.rs:34-39"""
    return list(zip(word, word[1:]))


def whitespace_clean(text: str) -> str:
    """clip.rs:41-43"""
    return " ".join(w for w in _WS.split(text) if w)


def load_merges(path: str) -> List[Tuple[str, str]]:
    """clip.rs:45-61"""
    merges = []
    with open(path, encoding="utf-8", newline="\n") as f:
        for line in f:
            words = [w for w in _WS.split(line) if w]
            if len(words) >= 2:
                merges.append((words[0], words[1]))
    return merges


class _Base:
    byte_encoder: Dict[int, str]
    byte_decoder: Dict[str, int]
    encoder: Dict[str, int]
    decoder: Dict[int, str]
    bpe_ranks: Dict[Tuple[str, str], int]
    cache: Dict[str, str]

    def bpe(self, token: str) -> str:
        """clip.rs:125-178"""
        if token in self.cache:
            return self.cache[token]
        word = list(token)
        if word:
            word[-1] += "</w>"
        pairs = get_pairs(word)
        if not pairs:
            return token + "</w>"
        while True:
            cand = [p for p in pairs if p in self.bpe_ranks]
            if not cand:
                break
            first, second = min(cand, key=lambda p: self.bpe_ranks[p])  # first minimum, like Iterator::min_by_key
            new_word: List[str] = []
            i = 0
            while i < len(word):
                try:
                    j = word.index(first, i)
                except ValueError:
                    new_word.extend(word[i:])
                    break
                new_word.extend(word[i:j])
                i = j
                if word[i] == first and i < len(word) - 1 and word[i + 1] == second:
                    new_word.append(first + second)
                    i += 2
                else:
                    new_word.append(word[i])
                    i += 1
            word = new_word
            if len(word) == 1:
                break
            pairs = get_pairs(word)
        return " ".join(word)

    def encode(self, text: str, add_sot: bool, add_eot: bool) -> List[int]:
        """clip.rs:182-205"""
        cleaned_text = whitespace_clean(text).lower()  # trim() is subsumed by split_whitespace
        bpe_tokens: List[int] = []
        if add_sot:
            bpe_tokens.append(self.start_of_text_token())
        for m in self.pat.finditer(cleaned_text):
            token = "".join(self.byte_encoder[b] for b in m.group(0).encode("utf-8"))
            bpe_tokens.extend(self.encoder[t] for t in self.bpe(token).split(" "))
        if add_eot:
            bpe_tokens.append(self.end_of_text_token())
        return bpe_tokens

    def decode(self, tokens: List[int]) -> str:
        """clip.rs:207-213"""
        text = "".join(self.decoder[t] for t in tokens)
        decoded = bytes(self.byte_decoder[c] for c in text)
        return decoded.decode("utf-8", errors="replace").replace("</w>", " ")

    def start_of_text_token(self) -> int:
        return 49406

    def end_of_text_token(self) -> int:
        return 49407


class ClipTokenizer(_Base):
    def __init__(self, merges_path: str):
        """ClipTokenizer::new, clip.rs:91-122"""
        bu = bytes_to_unicode()
        self.byte_encoder = dict(bu)
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        merges = load_merges(merges_path)
        merges = merges[1:49152 - 256 - 2 + 1]
        chars = [u for _, u in bu]
        vocab = chars + [c + "</w>" for c in chars] + [a + b for a, b in merges] + ["<|startoftext|>", "<|endoftext|>"]  # clip.rs:63-77
        self.encoder = {s: i for i, s in enumerate(vocab)}
        self.decoder = {v: k for k, v in self.encoder.items()}
        self.bpe_ranks = {m: i for i, m in enumerate(merges)}
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}
        self.pat = regex.compile(PAT)

    def padding_token(self) -> int:
        return self.end_of_text_token()


class OpenClipTokenizer(_Base):
    def __init__(self, merges_path: str, vocab_path: str):
        """OpenClipTokenizer::new, open_clip.rs:82-113"""
        bu = bytes_to_unicode()
        self.byte_encoder = dict(bu)
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        merges = load_merges(merges_path)
        with open(vocab_path, encoding="utf-8", newline="\n") as f:
            vocab = [ln[:-1] if ln.endswith("\n") else ln for ln in f]
        vocab = [v[:-1] if v.endswith("\r") else v for v in vocab]
        self.encoder = {s: i for i, s in enumerate(vocab)}
        self.decoder = {v: k for k, v in self.encoder.items()}
        self.bpe_ranks = {m: i for i, m in enumerate(merges)}
        self.cache = {}
        self.pat = regex.compile(PAT)

    def padding_token(self) -> int:
        return 0


def tokenize_text(text: str, tokenizer: _Base, seq_len: int) -> List[int]:
    """stablediffusion/mod.rs:778-793: encode(text, true, true) then Vec::resize(seq_len, padding_token)."""
    t = tokenizer.encode(text, True, True)
    return (t + [tokenizer.padding_token()] * seq_len)[:seq_len]
