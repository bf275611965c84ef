// BPE tokenizers of the Embedder: CPU host code behind the C ABI, integer/byte work only.
//   ClipTokenizer      reference src/token/clip.rs:11-230      (merges file -> vocab; pads with <|endoftext|> = 49407)
//   OpenClipTokenizer  reference src/token/open_clip.rs:63-221 (merges.txt + vocab.txt; empty cache; pads with 0)
//   tokenize_text      reference src/model/stablediffusion/mod.rs:778-793
// The reference leans on Rust's `regex` (Unicode classes, (?i)) and `str::to_lowercase`; both are restated here over
// the generated tables in unicode_tables.h. Results are bit-exact token ids: tests/test_tokenizer.py pins them with the
// reference's own known-answer vector (clip.rs:232-249) and a line-by-line Python restatement (oracle/tokenizer_oracle.py).
#include "../../include/sdxl_b200.h"
#include "unicode_tables.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_tok_err;
int tfail(int code, const std::string& msg) {
  g_tok_err = msg;
  return code;
}

// ---- UTF-8 ---------------------------------------------------------------------------------------------------
void put_utf8(std::string& s, uint32_t cp) {
  if (cp < 0x80) s.push_back((char)cp);
  else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 0x3F))); }
  else if (cp < 0x10000) { s.push_back((char)(0xE0 | (cp >> 12))); s.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s.push_back((char)(0x80 | (cp & 0x3F))); }
  else { s.push_back((char)(0xF0 | (cp >> 18))); s.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); s.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s.push_back((char)(0x80 | (cp & 0x3F))); }
}
// Lossy decode (String::from_utf8_lossy semantics: each maximal invalid subpart becomes one U+FFFD).
std::vector<uint32_t> decode_utf8_lossy(const unsigned char* p, size_t n) {
  std::vector<uint32_t> out;
  size_t i = 0;
  while (i < n) {
    const unsigned char b = p[i];
    if (b < 0x80) { out.push_back(b); ++i; continue; }
    int need = 0;
    uint32_t cp = 0;
    unsigned char lo = 0x80, hi = 0xBF;
    if (b >= 0xC2 && b <= 0xDF) { need = 1; cp = b & 0x1F; }
    else if (b >= 0xE0 && b <= 0xEF) { need = 2; cp = b & 0x0F; if (b == 0xE0) lo = 0xA0; if (b == 0xED) hi = 0x9F; }
    else if (b >= 0xF0 && b <= 0xF4) { need = 3; cp = b & 0x07; if (b == 0xF0) lo = 0x90; if (b == 0xF4) hi = 0x8F; }
    else { out.push_back(0xFFFD); ++i; continue; }
    size_t j = i + 1;
    int got = 0;
    while (got < need && j < n) {
      const unsigned char c = p[j];
      const unsigned char l = got == 0 ? lo : 0x80, h = got == 0 ? hi : 0xBF;
      if (c < l || c > h) break;
      cp = (cp << 6) | (c & 0x3F);
      ++j;
      ++got;
    }
    if (got == need) out.push_back(cp);
    else out.push_back(0xFFFD);
    i = j;
  }
  return out;
}

// ---- Unicode properties --------------------------------------------------------------------------------------
bool in_ranges(const uint32_t (*r)[2], int n, uint32_t cp) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < r[mid][0]) hi = mid - 1;
    else if (cp > r[mid][1]) lo = mid + 1;
    else return true;
  }
  return false;
}
// \p{L} under (?i): Rust's regex closes Unicode classes under simple case folding, which adds exactly one code point to
// L — U+0345 COMBINING GREEK YPOGEGRAMMENI (Mn), whose fold is the letter U+03B9 (tools/gen_unicode_tables.py scan).
bool is_letter(uint32_t cp) { return cp == 0x345 || in_ranges(kUniLetter, kUniLetter_n, cp); }
bool is_number(uint32_t cp) { return in_ranges(kUniNumber, kUniNumber_n, cp); }
// White_Space property (char::is_whitespace, regex \s)
bool is_space(uint32_t cp) {
  return (cp >= 0x9 && cp <= 0xD) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) ||
         cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}
bool is_cased(uint32_t cp) { return in_ranges(kUniCased, kUniCased_n, cp); }
bool is_case_ignorable(uint32_t cp) { return in_ranges(kUniCaseIgnorable, kUniCaseIgnorable_n, cp); }
const uint32_t* lower_entry(uint32_t cp) {
  int lo = 0, hi = kUniLower_n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < kUniLower[mid][0]) hi = mid - 1;
    else if (cp > kUniLower[mid][0]) lo = mid + 1;
    else return kUniLower[mid];
  }
  return nullptr;
}
// str::to_lowercase: per-char full lower-case mapping, with the final-sigma rule for U+03A3.
std::vector<uint32_t> to_lowercase(const std::vector<uint32_t>& s) {
  std::vector<uint32_t> out;
  out.reserve(s.size());
  for (size_t i = 0; i < s.size(); ++i) {
    const uint32_t c = s[i];
    if (c == 0x3A3) {
      // final sigma: preceded by a cased letter (skipping case-ignorables) and not followed by one
      bool before = false, after = false;
      for (size_t j = i; j-- > 0;) {
        if (is_case_ignorable(s[j])) continue;
        before = is_cased(s[j]);
        break;
      }
      for (size_t j = i + 1; j < s.size(); ++j) {
        if (is_case_ignorable(s[j])) continue;
        after = is_cased(s[j]);
        break;
      }
      out.push_back(before && !after ? 0x3C2 : 0x3C3);
      continue;
    }
    if (c < 0x80) { out.push_back((c >= 'A' && c <= 'Z') ? c + 32 : c); continue; }
    const uint32_t* e = lower_entry(c);
    if (!e) { out.push_back(c); continue; }
    out.push_back(e[1]);
    if (e[2]) out.push_back(e[2]);
  }
  return out;
}

// ---- the reference's pre-tokeniser regex, hand-compiled ------------------------------------------------------------
// (?i)<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|\p{L}+|\p{N}|[^\s\p{L}\p{N}]+      (clip.rs:111)
// leftmost-first alternation; text is already lower-cased, so (?i) only adds the simple case folds of ASCII letters
// that survive lower-casing: U+017F (long s) ~ 's'. (U+212A KELVIN SIGN lower-cases to 'k' before we get here.)
bool ci_eq(uint32_t c, char lit) {
  if (c == (uint32_t)(unsigned char)lit) return true;
  if (lit >= 'a' && lit <= 'z' && c == (uint32_t)(lit - 32)) return true;
  return lit == 's' && c == 0x17F;
}
size_t match_lit(const std::vector<uint32_t>& s, size_t i, const char* lit) {
  size_t k = 0;
  for (; lit[k]; ++k)
    if (i + k >= s.size() || !ci_eq(s[i + k], lit[k])) return 0;
  return k;
}
// returns the match length at i (0 = no match starting here)
size_t match_at(const std::vector<uint32_t>& s, size_t i) {
  static const char* const lits[] = {"<|startoftext|>", "<|endoftext|>", "'s", "'t", "'re", "'ve", "'m", "'ll", "'d"};
  for (const char* l : lits) {
    const size_t k = match_lit(s, i, l);
    if (k) return k;
  }
  const uint32_t c = s[i];
  if (is_letter(c)) {
    size_t j = i + 1;
    while (j < s.size() && is_letter(s[j])) ++j;
    return j - i;
  }
  if (is_number(c)) return 1;
  if (!is_space(c)) {
    size_t j = i + 1;
    while (j < s.size() && !is_space(s[j]) && !is_letter(s[j]) && !is_number(s[j])) ++j;
    return j - i;
  }
  return 0;
}

}  // namespace

struct sdxl_tokenizer {
  bool open_clip = false;
  uint32_t byte_encoder[256];                         // byte -> code point (bytes_to_unicode, clip.rs:11-32)
  std::unordered_map<uint32_t, uint8_t> byte_decoder;
  std::unordered_map<std::string, uint32_t> encoder;  // vocab string -> id
  std::vector<std::string> decoder;                   // id -> vocab string
  std::unordered_map<std::string, uint32_t> bpe_ranks;  // "first second" -> rank
  std::unordered_map<std::string, std::string> cache;
  uint32_t sot = 49406, eot = 49407, pad = 49407;
};

namespace {

void build_bytes_to_unicode(sdxl_tokenizer* t, std::vector<uint32_t>& order) {
  std::vector<int> bs;
  for (int b = '!'; b <= '~'; ++b) bs.push_back(b);
  for (int b = 0xA1; b <= 0xAC; ++b) bs.push_back(b);
  for (int b = 0xAE; b <= 0xFF; ++b) bs.push_back(b);
  std::vector<uint32_t> cs(bs.begin(), bs.end());
  uint32_t n = 0;
  for (int b = 0; b < 256; ++b)
    if (std::find(bs.begin(), bs.end(), b) == bs.end()) {
      bs.push_back(b);
      cs.push_back(256 + n++);
    }
  for (size_t i = 0; i < bs.size(); ++i) {
    t->byte_encoder[bs[i]] = cs[i];
    t->byte_decoder[cs[i]] = (uint8_t)bs[i];
  }
  order = cs;
}

// load_merges (clip.rs:44-61): first two whitespace-separated words of every line that has at least two
int load_merges(const char* path, std::vector<std::pair<std::string, std::string>>& merges) {
  std::ifstream f(path);
  if (!f) return tfail(7001, std::string("cannot open merges file '") + path + "'");
  std::string line;
  while (std::getline(f, line)) {
    const std::vector<uint32_t> cps = decode_utf8_lossy((const unsigned char*)line.data(), line.size());
    std::vector<std::string> words;
    std::string cur;
    for (uint32_t c : cps) {
      if (is_space(c)) {
        if (!cur.empty()) { words.push_back(cur); cur.clear(); if (words.size() == 2) break; }
      } else {
        put_utf8(cur, c);
      }
    }
    if (!cur.empty() && words.size() < 2) words.push_back(cur);
    if (words.size() >= 2) merges.emplace_back(words[0], words[1]);
  }
  return 0;
}

void finish(sdxl_tokenizer* t, const std::vector<std::string>& vocab, const std::vector<std::pair<std::string, std::string>>& merges) {
  t->decoder = vocab;
  for (size_t i = 0; i < vocab.size(); ++i) t->encoder[vocab[i]] = (uint32_t)i;  // later duplicates win (HashMap collect)
  // decoder is built from the encoder in the reference (clip.rs:104): a duplicated string decodes from its LAST index only;
  // earlier indices of a duplicate are absent there (lookup would panic). Keep id -> string total here; see decode().
  for (size_t i = 0; i < merges.size(); ++i) t->bpe_ranks[merges[i].first + " " + merges[i].second] = (uint32_t)i;
}

// ClipTokenizer::bpe / OpenClipTokenizer::bpe (clip.rs:125-178); word pieces are UTF-8 strings
std::vector<std::string> bpe(const sdxl_tokenizer* t, const std::vector<uint32_t>& token_cps, const std::string& token) {
  auto it = t->cache.find(token);
  if (it != t->cache.end()) return {it->second};
  std::vector<std::string> word;
  for (uint32_t c : token_cps) {
    std::string s;
    put_utf8(s, c);
    word.push_back(s);
  }
  if (word.empty()) return {std::string("</w>")};  // format!("{}{}", token, "</w>") with an empty token
  word.back() += "</w>";
  if (word.size() < 2) return {token + "</w>"};
  for (;;) {
    uint32_t best = 0xFFFFFFFFu;
    size_t best_i = 0;
    for (size_t i = 0; i + 1 < word.size(); ++i) {
      auto r = t->bpe_ranks.find(word[i] + " " + word[i + 1]);
      if (r != t->bpe_ranks.end() && r->second < best) { best = r->second; best_i = i; }
    }
    if (best == 0xFFFFFFFFu) break;
    const std::string first = word[best_i], second = word[best_i + 1];
    std::vector<std::string> nw;
    size_t i = 0;
    while (i < word.size()) {
      size_t j = i;
      while (j < word.size() && word[j] != first) ++j;
      if (j == word.size()) { nw.insert(nw.end(), word.begin() + i, word.end()); break; }
      nw.insert(nw.end(), word.begin() + i, word.begin() + j);
      i = j;
      if (word[i] == first && i < word.size() - 1 && word[i + 1] == second) { nw.push_back(first + second); i += 2; }
      else { nw.push_back(word[i]); i += 1; }
    }
    word.swap(nw);
    if (word.size() == 1) break;
  }
  return word;
}

int encode_impl(const sdxl_tokenizer* t, const char* text, int add_sot, int add_eot, std::vector<uint32_t>& ids) {
  // whitespace_clean(text.trim()).to_lowercase()   (clip.rs:183)
  const std::vector<uint32_t> raw = decode_utf8_lossy((const unsigned char*)text, strlen(text));
  std::vector<uint32_t> cleaned;
  bool pending_space = false;
  for (uint32_t c : raw) {
    if (is_space(c)) { pending_space = !cleaned.empty(); continue; }
    if (pending_space) { cleaned.push_back(' '); pending_space = false; }
    cleaned.push_back(c);
  }
  const std::vector<uint32_t> s = to_lowercase(cleaned);
  if (add_sot) ids.push_back(t->sot);
  size_t i = 0;
  while (i < s.size()) {
    const size_t k = match_at(s, i);
    if (!k) { ++i; continue; }
    // token bytes -> byte_encoder chars (clip.rs:193-198)
    std::string utf8;
    for (size_t j = i; j < i + k; ++j) put_utf8(utf8, s[j]);
    std::vector<uint32_t> enc;
    std::string enc_s;
    for (unsigned char b : utf8) { enc.push_back(t->byte_encoder[b]); put_utf8(enc_s, t->byte_encoder[b]); }
    for (const std::string& piece : bpe(t, enc, enc_s)) {
      // the cached special tokens come back as one string; everything else is already split
      auto e = t->encoder.find(piece);
      if (e == t->encoder.end()) return tfail(7010, "token piece '" + piece + "' is not in the vocabulary (the reference panics here)");
      ids.push_back(e->second);
    }
    i += k;
  }
  if (add_eot) ids.push_back(t->eot);
  return 0;
}

}  // namespace

extern "C" const char* sdxl_tokenizer_last_error(void) { return g_tok_err.c_str(); }

extern "C" int sdxl_tokenizer_create_clip(const char* merges_path, sdxl_tokenizer** out) {
  if (!merges_path || !out) return tfail(-1, "sdxl_tokenizer_create_clip: null argument");
  *out = nullptr;
  std::unique_ptr<sdxl_tokenizer> t(new sdxl_tokenizer());
  std::vector<uint32_t> order;
  build_bytes_to_unicode(t.get(), order);
  std::vector<std::pair<std::string, std::string>> all;
  int r = load_merges(merges_path, all);
  if (r) return r;
  // merges[1..49152 - 256 - 2 + 1]   (clip.rs:98)
  const size_t lo = 1, hi = 49152 - 256 - 2 + 1;
  if (all.size() < hi) return tfail(7002, "merges file has " + std::to_string(all.size()) + " entries, the CLIP tokenizer needs " + std::to_string(hi));
  std::vector<std::pair<std::string, std::string>> merges(all.begin() + lo, all.begin() + hi);
  // construct_vocab (clip.rs:63-77)
  std::vector<std::string> vocab;
  for (uint32_t c : order) { std::string s; put_utf8(s, c); vocab.push_back(s); }
  for (uint32_t c : order) { std::string s; put_utf8(s, c); vocab.push_back(s + "</w>"); }
  for (auto& m : merges) vocab.push_back(m.first + m.second);
  vocab.push_back("<|startoftext|>");
  vocab.push_back("<|endoftext|>");
  finish(t.get(), vocab, merges);
  t->cache["<|startoftext|>"] = "<|startoftext|>";
  t->cache["<|endoftext|>"] = "<|endoftext|>";
  t->pad = t->eot;  // clip.rs:227-229
  *out = t.release();
  return 0;
}

extern "C" int sdxl_tokenizer_create_open_clip(const char* merges_path, const char* vocab_path, sdxl_tokenizer** out) {
  if (!merges_path || !vocab_path || !out) return tfail(-1, "sdxl_tokenizer_create_open_clip: null argument");
  *out = nullptr;
  std::unique_ptr<sdxl_tokenizer> t(new sdxl_tokenizer());
  t->open_clip = true;
  std::vector<uint32_t> order;
  build_bytes_to_unicode(t.get(), order);
  std::vector<std::pair<std::string, std::string>> merges;
  int r = load_merges(merges_path, merges);
  if (r) return r;
  // load_vocab (open_clip.rs:63-68): one entry per line, verbatim
  std::ifstream f(vocab_path);
  if (!f) return tfail(7003, std::string("cannot open vocab file '") + vocab_path + "'");
  std::vector<std::string> vocab;
  std::string line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();  // BufRead::lines strips "\r\n" too
    vocab.push_back(line);
  }
  finish(t.get(), vocab, merges);
  t->pad = 0;  // open_clip.rs:218-220
  *out = t.release();
  return 0;
}

extern "C" void sdxl_tokenizer_destroy(sdxl_tokenizer* t) { delete t; }

extern "C" int sdxl_tokenizer_encode(const sdxl_tokenizer* t, const char* text_utf8, int add_sot, int add_eot, uint32_t* ids_out,
                                     int capacity, int* n_out) {
  if (!t || !text_utf8 || !n_out) return tfail(-1, "sdxl_tokenizer_encode: null argument");
  std::vector<uint32_t> ids;
  int r = encode_impl(t, text_utf8, add_sot, add_eot, ids);
  if (r) return r;
  *n_out = (int)ids.size();
  if (ids_out) {
    if ((int)ids.size() > capacity) return tfail(7011, "ids_out too small: need " + std::to_string(ids.size()));
    memcpy(ids_out, ids.data(), ids.size() * sizeof(uint32_t));
  }
  return 0;
}

// tokenize_text (stablediffusion/mod.rs:778-793): encode(text, true, true), then resize to seq_len with the padding token
// (a longer sequence is cut, end-of-text included).
extern "C" int sdxl_tokenize_text(const sdxl_tokenizer* t, const char* text_utf8, int seq_len, int32_t* tokens_out) {
  if (!t || !text_utf8 || !tokens_out || seq_len < 0) return tfail(-1, "sdxl_tokenize_text: bad argument");
  std::vector<uint32_t> ids;
  int r = encode_impl(t, text_utf8, 1, 1, ids);
  if (r) return r;
  ids.resize((size_t)seq_len, t->pad);
  for (int i = 0; i < seq_len; ++i) tokens_out[i] = (int32_t)ids[i];
  return 0;
}

// Tokenizer::decode (clip.rs:207-213)
extern "C" int sdxl_tokenizer_decode(const sdxl_tokenizer* t, const uint32_t* ids, int n, char* out, int capacity, int* n_out) {
  if (!t || (!ids && n) || !n_out) return tfail(-1, "sdxl_tokenizer_decode: null argument");
  std::string text;
  for (int i = 0; i < n; ++i) {
    if (ids[i] >= t->decoder.size()) return tfail(7020, "token id " + std::to_string(ids[i]) + " out of range");
    text += t->decoder[ids[i]];
  }
  const std::vector<uint32_t> cps = decode_utf8_lossy((const unsigned char*)text.data(), text.size());
  std::string bytes;
  for (uint32_t c : cps) {
    auto it = t->byte_decoder.find(c);
    if (it == t->byte_decoder.end()) return tfail(7021, "character U+" + std::to_string(c) + " has no byte (the reference panics here)");
    bytes.push_back((char)it->second);
  }
  std::string lossy;
  for (uint32_t c : decode_utf8_lossy((const unsigned char*)bytes.data(), bytes.size())) put_utf8(lossy, c);
  std::string res;
  for (size_t i = 0; i < lossy.size();) {
    if (lossy.compare(i, 4, "</w>") == 0) { res.push_back(' '); i += 4; }
    else res.push_back(lossy[i++]);
  }
  *n_out = (int)res.size();
  if (out) {
    if ((int)res.size() + 1 > capacity) return tfail(7022, "decode buffer too small: need " + std::to_string(res.size() + 1));
    memcpy(out, res.c_str(), res.size() + 1);
  }
  return 0;
}

extern "C" int sdxl_tokenizer_special(const sdxl_tokenizer* t, uint32_t* sot, uint32_t* eot, uint32_t* pad) {
  if (!t) return tfail(-1, "null tokenizer");
  if (sot) *sot = t->sot;
  if (eot) *eot = t->eot;
  if (pad) *pad = t->pad;
  return 0;
}
