// Host-side helper for the reference's shipped weight format: burn 0.13 `NamedMpkFileRecorder<HalfPrecisionSettings>`
// (MessagePack; src/bin/convert/main.rs:65-70, src/bin/sample/main.rs:28-51 in the reference). Every tensor is a
// `{"value": [u16 ...], "shape": [...]}` map whose `value` array holds the f16 BIT PATTERNS as variable-length MessagePack
// unsigned integers (half::f16 serialises as a newtype over u16) — 2.6 G of them for the base UNet. The tree walk lives in
// sdxl_b200/burn_record.py; this is the one hot loop: decode `count` consecutive MessagePack unsigned integers into u16.
#include <stddef.h>
#include <stdint.h>

#include "../../include/sdxl_b200.h"

extern "C" int sdxl_mpk_decode_u16(const uint8_t* buf, size_t len, size_t count, uint16_t* out, size_t* consumed) {
  if (!buf || !out || !consumed) return -1;
  size_t p = 0;
  for (size_t i = 0; i < count; ++i) {
    if (p >= len) return 7001;                       // truncated
    const uint8_t t = buf[p];
    if (t < 0x80) { out[i] = t; p += 1; }            // positive fixint
    else if (t == 0xcc) { if (p + 2 > len) return 7001; out[i] = buf[p + 1]; p += 2; }
    else if (t == 0xcd) { if (p + 3 > len) return 7001; out[i] = (uint16_t)((buf[p + 1] << 8) | buf[p + 2]); p += 3; }
    else if (t == 0xce) {                            // u32 (a writer is free to widen): must still fit 16 bits
      if (p + 5 > len) return 7001;
      if (buf[p + 1] | buf[p + 2]) return 7002;
      out[i] = (uint16_t)((buf[p + 3] << 8) | buf[p + 4]);
      p += 5;
    } else return 7003;                              // not an unsigned integer
  }
  *consumed = p;
  return 0;
}

// Inverse (fixture writer / `convert`): returns the number of bytes written (out must hold 3 * count).
extern "C" size_t sdxl_mpk_encode_u16(const uint16_t* in, size_t count, uint8_t* out) {
  size_t p = 0;
  for (size_t i = 0; i < count; ++i) {
    const uint16_t v = in[i];
    if (v < 0x80) out[p++] = (uint8_t)v;
    else if (v < 0x100) { out[p++] = 0xcc; out[p++] = (uint8_t)v; }
    else { out[p++] = 0xcd; out[p++] = (uint8_t)(v >> 8); out[p++] = (uint8_t)v; }
  }
  return p;
}
