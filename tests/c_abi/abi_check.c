/* Plain-C consumer of include/sdxl_b200.h: proves the header is C (not C++), that every declared entry point links against
 * libsdxl_b200.so, and exercises the CPU-only part of the ABI (the tokenizers) the way a cgo / Rust `extern "C"` binding would.
 * Built and run by tests/test_library_cpu.py; no GPU call is made. */
#include <stdio.h>
#include <string.h>

#include "sdxl_b200.h"

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: abi_check merges.txt vocab.txt\n"); return 2; }
  /* take the address of the entry points a binding uses: unresolved symbols fail at link time (function-pointer typed, so
   * the prototypes of the header are checked too) */
  typedef void (*fn_t)(void);
  fn_t fns[] = {(fn_t)sdxl_ctx_create, (fn_t)sdxl_ctx_destroy, (fn_t)sdxl_last_error, (fn_t)sdxl_unet_load, (fn_t)sdxl_unet_forward,
                (fn_t)sdxl_sample_latent, (fn_t)sdxl_qkv_attention, (fn_t)sdxl_vae_load, (fn_t)sdxl_vae_decode_latent,
                (fn_t)sdxl_vae_latent_to_image, (fn_t)sdxl_vae_encode_image, (fn_t)sdxl_vae_image_to_latent, (fn_t)sdxl_clip_load,
                (fn_t)sdxl_clip_forward_hidden, (fn_t)sdxl_clip_forward_hidden_pooled, (fn_t)sdxl_tokenizer_create_clip,
                (fn_t)sdxl_tokenize_text};
  size_t i;
  for (i = 0; i < sizeof fns / sizeof fns[0]; ++i)
    if (!fns[i]) return 3;
  /* struct layouts a binding relies on */
  if (sizeof(((sdxl_unet_cfg*)0)->channel_mults) != SDXL_MAX_LEVELS * sizeof(int32_t)) return 4;
  sdxl_tokenizer* tok = NULL;
  if (sdxl_tokenizer_create_open_clip(argv[1], argv[2], &tok) != 0) {
    fprintf(stderr, "create failed: %s\n", sdxl_tokenizer_last_error());
    return 5;
  }
  int32_t ids[77];
  if (sdxl_tokenize_text(tok, "a photo of a cat", 77, ids) != 0) return 6;
  if (ids[0] != 49406) return 7;                       /* start-of-text */
  int n_eot = 0, j;
  for (j = 0; j < 77; ++j) n_eot += ids[j] == 49407;
  if (n_eot != 1 || ids[76] != 0) return 8;            /* one end-of-text, zero padding (OpenCLIP) */
  sdxl_tokenizer* bad = NULL;
  if (sdxl_tokenizer_create_clip("/nonexistent", &bad) == 0 || strlen(sdxl_tokenizer_last_error()) == 0) return 9;
  sdxl_tokenizer_destroy(tok);
  printf("abi_check ok: %d %d %d ...\n", ids[0], ids[1], ids[2]);
  return 0;
}
