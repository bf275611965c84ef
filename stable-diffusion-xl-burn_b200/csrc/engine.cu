// Host side of libsdxl_b200.so: context, weight-pack loader (re-layout on device), the UNet launch
// plan (a flat list of kernel launches with pre-built TMA descriptors, replayed as a CUDA graph), the
// DDIM/CFG sampler loop, and the C ABI of include/sdxl_b200.h. No torch, no cuBLAS/cuDNN: every device
// op is one of this library's own sm_100a kernels.
//
// Structure mirrored from the reference (file:line relative to the reference root):
//   UNet::forward               src/model/unet/mod.rs:449-493
//   UNetConfig::init (blocks)   src/model/unet/mod.rs:72-430
//   ResBlock / SpatialTransformer / TransformerBlock / MHA / GEGLU
//                               src/model/unet/mod.rs:1082-1106, 820-845, 885-891, 1005-1023, 942-956
//   Diffuser::{sample_latent, sample_latent_with_inpainting, refine_latent, diffuse_latent*,
//              forward_diffuser, get_alpha}
//                               src/model/stablediffusion/mod.rs:317-541
#include "../../include/sdxl_b200.h"
#include "kernels.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

using namespace sdxl;

// ================================================================================================
// context
// ================================================================================================
struct sdxl_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  std::string err;
  uint64_t launches = 0;
};

static int fail(sdxl_ctx* c, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code ? code : -1;
}
#define CU(ctx, expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(ctx, (int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define KL(ctx, expr)                                                                         \
  do {                                                                                        \
    int _e = (expr);                                                                          \
    if (_e) return fail(ctx, _e, "%s failed with %d%s%s (%s:%d)", #expr, _e, _e < 1000 ? ": " : "", \
                        _e < 1000 ? cudaGetErrorString((cudaError_t)_e) : "", __FILE__, __LINE__); \
    (ctx)->launches++;                                                                        \
  } while (0)

extern "C" int sdxl_ctx_create(int device, void* cuda_stream, sdxl_ctx** out) {
  if (!out) return -1;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device >= n) {
    fprintf(stderr, "sdxl_b200: no CUDA device %d (this library has no CPU fallback)\n", device);
    return -2;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -3;
  if (prop.major != 10) {
    fprintf(stderr, "sdxl_b200: device %d is sm_%d%d; this library contains sm_100a code only\n", device,
            prop.major, prop.minor);
    return -4;
  }
  if (cudaSetDevice(device) != cudaSuccess) return -5;
  sdxl_ctx* c = new sdxl_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  if (cuda_stream) {
    c->stream = (cudaStream_t)cuda_stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return -6;
    }
    c->own_stream = true;
  }
  *out = c;
  return 0;
}
extern "C" void sdxl_ctx_destroy(sdxl_ctx* c) {
  if (!c) return;
  cudaStreamSynchronize(c->stream);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}
extern "C" const char* sdxl_last_error(const sdxl_ctx* c) { return c ? c->err.c_str() : "null ctx"; }
extern "C" int sdxl_ctx_synchronize(sdxl_ctx* c) {
  CU(c, cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" uint64_t sdxl_ctx_launch_count(const sdxl_ctx* c) { return c ? c->launches : 0; }

// ================================================================================================
// device arena (bump allocator over one cudaMalloc)
// ================================================================================================
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0;
  bool measure = false;  // dry run: only count
  int init(size_t bytes) {
    release();
    if (cudaMalloc((void**)&base, bytes) != cudaSuccess) return 1;
    cap = bytes;
    off = 0;
    return 0;
  }
  void release() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = off = 0;
  }
  void* alloc(size_t bytes) {
    const size_t a = (off + 1023) & ~size_t(1023);
    if (!measure && a + bytes > cap) return nullptr;
    off = a + bytes;
    return measure ? (void*)(uintptr_t)(0x1000 + a) : (void*)(base + a);
  }
  template <typename T>
  T* get(size_t n) { return (T*)alloc(n * sizeof(T)); }
};

// ================================================================================================
// weight pack
// ================================================================================================
#pragma pack(push, 1)
struct PackHeader {
  char magic[8];  // "SDXLPK01"
  uint32_t n_tensors;
  uint32_t reserved;
  uint64_t data_offset;
};
struct PackEntry {
  char name[120];
  uint32_t dtype;  // 0 = f16, 1 = f32
  uint32_t ndim;
  uint64_t shape[4];
  uint64_t offset;  // from pack start
  uint64_t nbytes;
};
#pragma pack(pop)

struct PackView {
  const uint8_t* dev = nullptr;  // pack bytes in device memory
  std::map<std::string, PackEntry> t;
  const PackEntry* find(const std::string& n) const {
    auto it = t.find(n);
    return it == t.end() ? nullptr : &it->second;
  }
};

// ================================================================================================
// model
// ================================================================================================
struct Lin { __half* w = nullptr; float* b = nullptr; int K = 0, Kpad = 0, N = 0, geglu_bn = 0; };
struct Conv { __half* w = nullptr; float* b = nullptr; int I = 0, O = 0, ks = 0, Ipad = 0, I2 = 0, I2pad = 0, Ktot = 0; };
struct Norm { float* g = nullptr; float* b = nullptr; int C = 0; float eps = 1e-5f; };
struct TBlock {
  Norm n1, n2, n3;
  Lin qkv, out1;      // self-attention (fused [3C, C])
  Lin q2, kv2, out2;  // cross-attention (kv fused [2C, ctx])
  Lin ff1, ff2;
};
struct STrans { Norm norm; Lin proj_in, proj_out; std::vector<TBlock> blocks; int C = 0, n_head = 0; };
struct Res {
  Norm n_in, n_out;
  Conv conv_in, conv_out;  // conv_out carries the fused skip 1x1 segment when Cin != Cout
  int Cin = 0, Cout = 0, temb_off = 0;
  bool has_skip = false;
};
enum BlockType { BT_CONV, BT_RES, BT_DOWN, BT_REST, BT_RESTU, BT_RESU };
struct Block {
  BlockType type = BT_RES;
  Res res;
  STrans st;
  Conv conv;  // BT_CONV (unused: first conv has its own path), BT_DOWN, upsample conv
  int Cout = 0;
};

struct Plan;
struct Sampler;

struct sdxl_unet {
  sdxl_ctx* ctx = nullptr;
  sdxl_unet_cfg cfg{};
  Arena warena;  // re-laid-out weights
  // embeddings
  Lin t1, t2, l1, l2;      // time / label MLPs
  Lin temb_all;            // concatenated lin_embed of every ResBlock [sumC, 4mc], bias folded with conv_in bias
  float* conv0_w = nullptr;  // first conv [mc][3][3][4] f32
  float* conv0_b = nullptr;
  std::vector<Block> in_blocks, out_blocks;
  Res mid_res1, mid_res2;
  STrans mid_st;
  Norm norm_out;
  Conv conv_out;
  std::vector<double> alphas;  // host copy (f16-stored values widened)
  int n_tblocks = 0;
  // conditioning state
  Arena carena;
  int condB = 0, n_ctx = 0, ctx_pitch = 0;
  __half* ctx16 = nullptr;     // [B*n_ctx, ctx_pitch]
  float* y32 = nullptr;        // [B, adm]
  float* lab1 = nullptr;       // [B, 4mc]
  float* label_emb = nullptr;  // [B, 4mc]
  std::vector<__half*> kv;     // per transformer block [B*n_ctx, 2C]
  std::vector<int> kvC;
  uint64_t cond_version = 0;
  // plan
  std::unique_ptr<Plan> plan;
  std::unique_ptr<Sampler> sampler;
  int* t_dev = nullptr;
  int* t_pinned = nullptr;
};

// ------------------------------------------------------------------------------------------------
// loader helpers
// ------------------------------------------------------------------------------------------------
struct Loader {
  sdxl_unet* u;
  sdxl_ctx* c;
  const PackView* pv;
  Arena* A;
  cudaStream_t st;
  int err = 0;

  const PackEntry* need(const std::string& name, int ndim) {
    const PackEntry* e = pv->find(name);
    if (!e) { err = fail(c, 4001, "weight pack: missing tensor '%s'", name.c_str()); return nullptr; }
    if (e->dtype != 0) { err = fail(c, 4002, "weight pack: tensor '%s' must be f16", name.c_str()); return nullptr; }
    if ((int)e->ndim != ndim) { err = fail(c, 4003, "weight pack: tensor '%s' has ndim %u, expected %d", name.c_str(), e->ndim, ndim); return nullptr; }
    return e;
  }
  const __half* ptr(const PackEntry* e) { return (const __half*)(pv->dev + e->offset); }
  bool has(const std::string& name) { return pv->find(name) != nullptr; }

  float* vec_f32(const std::string& name, int expectN, int geglu_bn = 0) {
    const PackEntry* e = need(name, 1);
    if (!e) return nullptr;
    if ((int)e->shape[0] != expectN) { err = fail(c, 4004, "weight pack: '%s' has %llu elements, expected %d", name.c_str(), (unsigned long long)e->shape[0], expectN); return nullptr; }
    float* d = A->get<float>(expectN);
    if (!d) { err = fail(c, 4005, "weight arena exhausted"); return nullptr; }
    if (!A->measure) { int r = bias_to_f32_launch(st, ptr(e), expectN, d, geglu_bn, 0); if (r) err = fail(c, r, "bias_to_f32 failed"); }
    return d;
  }
  // Linear stored [in,out]; produce K-major [N,Kpad]. Rows may be a slice of a fused matrix.
  int lin_into(const std::string& path, __half* dst, int Kpad, int row0, int expectK, int expectN, int geglu_bn) {
    const PackEntry* e = need(path + "/weight", 2);
    if (!e) return err;
    if ((int)e->shape[0] != expectK || (int)e->shape[1] != expectN)
      return err = fail(c, 4006, "weight pack: '%s/weight' is [%llu,%llu], expected [%d,%d]", path.c_str(),
                        (unsigned long long)e->shape[0], (unsigned long long)e->shape[1], expectK, expectN);
    if (!A->measure) { int r = transpose_linear_launch(st, ptr(e), expectK, expectN, dst, Kpad, row0, geglu_bn); if (r) return err = fail(c, r, "transpose_linear failed"); }
    return 0;
  }
  static int pad64(int k) { return (k + 63) / 64 * 64; }
  Lin linear(const std::string& path, int K, int N, bool bias, int geglu_bn = 0) {
    Lin L;
    L.K = K; L.N = N; L.Kpad = pad64(K); L.geglu_bn = geglu_bn;
    L.w = A->get<__half>((size_t)N * L.Kpad);
    if (!L.w) { err = fail(c, 4005, "weight arena exhausted"); return L; }
    if (lin_into(path, L.w, L.Kpad, 0, K, N, geglu_bn)) return L;
    if (bias) L.b = vec_f32(path + "/bias", N, geglu_bn);
    return L;
  }
  Norm norm(const std::string& path, int C) {
    Norm n;
    n.C = C;
    n.g = vec_f32(path + "/weight", C);
    n.b = vec_f32(path + "/bias", C);
    return n;
  }
  // conv OIHW -> [O, ks*ks*Ipad (+ I2pad)]
  // Opad > O: the matrix (and bias) get zero rows up to Opad so the GEMM's N is a multiple of 4 (cv.O = Opad).
  Conv conv(const std::string& path, int I, int O, int ks, const std::string& skip_path = "", int I2 = 0, int Opad = 0) {
    Conv cv;
    cv.I = I; cv.O = O; cv.ks = ks; cv.Ipad = pad64(I); cv.I2 = I2; cv.I2pad = I2 ? pad64(I2) : 0;
    cv.Ktot = ks * ks * cv.Ipad + cv.I2pad;
    const int rows = Opad > O ? Opad : O;
    cv.w = A->get<__half>((size_t)rows * cv.Ktot);
    if (!cv.w) { err = fail(c, 4005, "weight arena exhausted"); return cv; }
    if (rows > O && !A->measure && cudaMemsetAsync(cv.w, 0, (size_t)rows * cv.Ktot * sizeof(__half), st) != cudaSuccess) {
      err = fail(c, 4011, "memset failed");
      return cv;
    }
    const PackEntry* e = need(path + "/weight", 4);
    if (!e) return cv;
    if ((int)e->shape[0] != O || (int)e->shape[1] != I || (int)e->shape[2] != ks || (int)e->shape[3] != ks) {
      err = fail(c, 4007, "weight pack: '%s/weight' has shape [%llu,%llu,%llu,%llu], expected [%d,%d,%d,%d]", path.c_str(),
                 (unsigned long long)e->shape[0], (unsigned long long)e->shape[1], (unsigned long long)e->shape[2],
                 (unsigned long long)e->shape[3], O, I, ks, ks);
      return cv;
    }
    if (!A->measure) { int r = repack_conv_launch(st, ptr(e), O, I, ks, ks, cv.w, cv.Ktot, 0, cv.Ipad); if (r) err = fail(c, r, "repack_conv failed"); }
    if (rows > O) {
      const PackEntry* be = need(path + "/bias", 1);
      cv.b = A->get<float>(rows);
      if (!be || !cv.b || (int)be->shape[0] != O) { if (!err) err = fail(c, 4012, "weight pack: '%s/bias' missing or mis-sized", path.c_str()); return cv; }
      if (!A->measure) {
        int r = (int)cudaMemsetAsync(cv.b, 0, rows * sizeof(float), st);
        if (!r) r = bias_to_f32_launch(st, ptr(be), O, cv.b, 0, 0);
        if (r) err = fail(c, r, "padded bias failed");
      }
      cv.O = rows;
    } else {
      cv.b = vec_f32(path + "/bias", O);
    }
    if (I2) {
      const PackEntry* s = need(skip_path + "/weight", 4);
      if (!s) return cv;
      if ((int)s->shape[0] != O || (int)s->shape[1] != I2 || s->shape[2] != 1 || s->shape[3] != 1) { err = fail(c, 4008, "weight pack: '%s/weight' bad shape", skip_path.c_str()); return cv; }
      const PackEntry* sb = need(skip_path + "/bias", 1);
      if (!sb) return cv;
      if (!A->measure) {
        int r = repack_conv_launch(st, ptr(s), O, I2, 1, 1, cv.w, cv.Ktot, ks * ks * cv.Ipad, cv.I2pad);
        if (!r) r = bias_to_f32_launch(st, ptr(sb), O, cv.b, 0, 1);
        if (r) err = fail(c, r, "skip repack failed");
      }
    }
    return cv;
  }
};

static int geglu_bn_for(int n_out /*4C*/) {
  for (int hb = 128; hb >= 32; hb >>= 1)
    if (n_out % hb == 0) return 2 * hb;
  return 0;
}

// temb bookkeeping while building: list of (lin_embed path, Cout, conv_in bias) in block order
struct TembItem { std::string path; int Cout; float* conv_bias; };

static Res load_res(Loader& L, const std::string& path, int Cin, int Cout, int temb_dim, std::vector<TembItem>& tembs, int& temb_total) {
  Res r;
  r.Cin = Cin; r.Cout = Cout; r.has_skip = (Cin != Cout);
  r.n_in = L.norm(path + "/norm_in", Cin);
  r.conv_in = L.conv(path + "/conv_in", Cin, Cout, 3);
  r.n_out = L.norm(path + "/norm_out", Cout);
  if (r.has_skip) r.conv_out = L.conv(path + "/conv_out", Cout, Cout, 3, path + "/skip_connection", Cin);
  else r.conv_out = L.conv(path + "/conv_out", Cout, Cout, 3);
  r.temb_off = temb_total;
  tembs.push_back({path + "/lin_embed", Cout, r.conv_in.b});
  temb_total += Cout;
  (void)temb_dim;
  return r;
}

static STrans load_st(Loader& L, const std::string& path, int C, int ctx_dim, int n_head, int depth) {
  STrans s;
  s.C = C; s.n_head = n_head;
  s.norm = L.norm(path + "/norm", C);
  s.proj_in = L.linear(path + "/proj_in", C, C, true);
  s.proj_out = L.linear(path + "/proj_out", C, C, true);
  const int ctx_pad = Loader::pad64(ctx_dim);
  const int Cpad = Loader::pad64(C);
  for (int j = 0; j < depth && !L.err; ++j) {
    const std::string bp = path + "/transformer_" + std::to_string(j);
    TBlock b;
    b.n1 = L.norm(bp + "/norm1", C);
    b.n2 = L.norm(bp + "/norm2", C);
    b.n3 = L.norm(bp + "/norm3", C);
    // fused QKV for self-attention (reference unet/mod.rs:1009-1011: three bias-free Linears on x)
    b.qkv.K = C; b.qkv.Kpad = Cpad; b.qkv.N = 3 * C;
    b.qkv.w = L.A->get<__half>((size_t)3 * C * Cpad);
    if (!b.qkv.w) { L.err = fail(L.c, 4005, "weight arena exhausted"); break; }
    L.lin_into(bp + "/attn1/query", b.qkv.w, Cpad, 0, C, C, 0);
    L.lin_into(bp + "/attn1/key", b.qkv.w, Cpad, C, C, C, 0);
    L.lin_into(bp + "/attn1/value", b.qkv.w, Cpad, 2 * C, C, C, 0);
    b.out1 = L.linear(bp + "/attn1/out", C, C, true);
    b.q2 = L.linear(bp + "/attn2/query", C, C, false);
    b.kv2.K = ctx_dim; b.kv2.Kpad = ctx_pad; b.kv2.N = 2 * C;
    b.kv2.w = L.A->get<__half>((size_t)2 * C * ctx_pad);
    if (!b.kv2.w) { L.err = fail(L.c, 4005, "weight arena exhausted"); break; }
    L.lin_into(bp + "/attn2/key", b.kv2.w, ctx_pad, 0, ctx_dim, C, 0);
    L.lin_into(bp + "/attn2/value", b.kv2.w, ctx_pad, C, ctx_dim, C, 0);
    b.out2 = L.linear(bp + "/attn2/out", C, C, true);
    const int gbn = geglu_bn_for(4 * C);
    if (!gbn) { L.err = fail(L.c, 4009, "GEGLU width %d not tileable", 4 * C); break; }
    b.ff1 = L.linear(bp + "/mlp/geglu/proj", C, 8 * C, true, gbn);
    b.ff2 = L.linear(bp + "/mlp/lin", 4 * C, C, true);
    s.blocks.push_back(b);
  }
  return s;
}

// Builds every layer (in measure mode only sizes are accumulated).
static int build_model(sdxl_unet* u, const PackView& pv, Arena& A) {
  sdxl_ctx* c = u->ctx;
  const sdxl_unet_cfg& g = u->cfg;
  Loader L{u, c, &pv, &A, c->stream};
  const int mc = g.model_channels, ted = 4 * mc;
  u->in_blocks.clear();
  u->out_blocks.clear();
  std::vector<TembItem> tembs;
  int temb_total = 0;
  auto n_head = [&](int ch) { return ch / g.n_head_channels; };

  u->t1 = L.linear("lin1_time_embed", mc, ted, true);
  u->t2 = L.linear("lin2_time_embed", ted, ted, true);
  u->l1 = L.linear("lin1_label_embed", g.adm_in_channels, ted, true);
  u->l2 = L.linear("lin2_label_embed", ted, ted, true);
  if (L.err) return L.err;

  // first conv: OIHW f16 -> [O][kh][kw][I] f32 (CUDA-core kernel)
  {
    const PackEntry* e = L.need("input_blocks/0/weight", 4);
    if (!e) return L.err;
    if ((int)e->shape[0] != mc || (int)e->shape[1] != g.in_channels || e->shape[2] != 3 || e->shape[3] != 3)
      return fail(c, 4010, "input_blocks/0/weight bad shape");
    const size_t n = (size_t)mc * 9 * g.in_channels;
    __half* tmp = A.get<__half>(n);
    u->conv0_w = A.get<float>(n);
    if (!tmp || !u->conv0_w) return fail(c, 4005, "weight arena exhausted");
    if (!A.measure) {
      int r = repack_conv_launch(c->stream, L.ptr(e), mc, g.in_channels, 3, 3, tmp, 9 * g.in_channels, 0, g.in_channels);
      if (!r) r = cast_f16_to_f32_launch(c->stream, tmp, n, u->conv0_w);
      if (r) return fail(c, r, "conv0 repack failed");
    }
    u->conv0_b = L.vec_f32("input_blocks/0/bias", mc);
    if (L.err) return L.err;
  }
  {
    Block b0; b0.type = BT_CONV; b0.Cout = mc;
    u->in_blocks.push_back(b0);
  }
  // input blocks (reference unet/mod.rs:121-173)
  int idx = 1;
  for (int level = 0; level < g.n_levels && !L.err; ++level) {
    const int cin = g.channel_mults[level > 0 ? level - 1 : 0] * mc;
    const int cout = g.channel_mults[level] * mc;
    const bool tr = (level == 1 || level == 2);
    for (int k = 0; k < 2; ++k) {
      Block b;
      const std::string bp = "input_blocks/" + std::to_string(idx++);
      b.Cout = cout;
      if (!tr) {
        b.type = BT_RES;
        b.res = load_res(L, bp, k == 0 ? cin : cout, cout, ted, tembs, temb_total);
      } else {
        b.type = BT_REST;
        b.res = load_res(L, bp + "/res", k == 0 ? cin : cout, cout, ted, tembs, temb_total);
        b.st = load_st(L, bp + "/transformer", cout, g.context_dim, n_head(cout), g.transformer_depths[level]);
      }
      u->in_blocks.push_back(std::move(b));
    }
    if (level != g.n_levels - 1) {
      Block b;
      b.type = BT_DOWN;
      b.Cout = cout;
      b.conv = L.conv("input_blocks/" + std::to_string(idx++), cout, cout, 3);
      u->in_blocks.push_back(std::move(b));
    }
  }
  if (L.err) return L.err;
  // middle (reference unet/mod.rs:238-248)
  {
    const int cm = g.channel_mults[g.n_levels - 1] * mc;
    u->mid_res1 = load_res(L, "middle_block/res1", cm, cm, ted, tembs, temb_total);
    u->mid_st = load_st(L, "middle_block/transformer", cm, g.context_dim, n_head(cm), g.transformer_depths[g.n_levels - 1]);
    u->mid_res2 = load_res(L, "middle_block/res2", cm, cm, ted, tembs, temb_total);
  }
  if (L.err) return L.err;
  // output blocks (reference unet/mod.rs:250-328)
  idx = 0;
  for (int level = g.n_levels - 1; level >= 0 && !L.err; --level) {
    const int next_level = (level != g.n_levels - 1) ? level + 1 : level;
    const int cout = g.channel_mults[level] * mc;
    const int cin1 = g.channel_mults[next_level] * mc + cout;
    const int cin2 = 2 * cout;
    const int cin3 = cout + g.channel_mults[level > 0 ? level - 1 : 0] * mc;
    const bool tr = (level == 1 || level == 2);
    const int cins[3] = {cin1, cin2, cin3};
    for (int k = 0; k < 3; ++k) {
      Block b;
      const std::string bp = "output_blocks/" + std::to_string(idx++);
      b.Cout = cout;
      const bool up = (k == 2) && (tr || level != 0);
      if (!tr) {
        b.type = up ? BT_RESU : BT_RES;
        b.res = load_res(L, up ? bp + "/res" : bp, cins[k], cout, ted, tembs, temb_total);
      } else {
        b.type = up ? BT_RESTU : BT_REST;
        b.res = load_res(L, bp + "/res", cins[k], cout, ted, tembs, temb_total);
        b.st = load_st(L, bp + "/transformer", cout, g.context_dim, n_head(cout), g.transformer_depths[level]);
      }
      if (up) b.conv = L.conv(bp + "/upsample/conv", cout, cout, 3);
      u->out_blocks.push_back(std::move(b));
    }
  }
  if (L.err) return L.err;
  u->norm_out = L.norm("norm_out", mc);
  u->conv_out = L.conv("conv_out", mc, g.out_channels, 3);
  if (L.err) return L.err;

  // concatenated lin_embed matrix (one GEMV per forward for all ResBlocks); bias += conv_in bias
  {
    Lin& T = u->temb_all;
    T.K = ted; T.Kpad = Loader::pad64(ted); T.N = temb_total;
    T.w = A.get<__half>((size_t)temb_total * T.Kpad);
    T.b = A.get<float>(temb_total);
    if (!T.w || !T.b) return fail(c, 4005, "weight arena exhausted");
    int off = 0;
    for (auto& it : tembs) {
      if (L.lin_into(it.path, T.w, T.Kpad, off, ted, it.Cout, 0)) return L.err;
      const PackEntry* e = L.need(it.path + "/bias", 1);
      if (!e) return L.err;
      if (!A.measure) {
        int r = bias_to_f32_launch(c->stream, L.ptr(e), it.Cout, T.b + off, 0, 0);
        // fold the conv_in bias: h = conv_in(..) + b_conv + lin_embed(..)   (unet/mod.rs:1086-1092)
        if (!r) r = vec_add_f32_launch(c->stream, T.b + off, it.conv_bias, it.Cout);
        if (r) return fail(c, r, "temb bias failed");
      }
      off += it.Cout;
    }
  }
  // count transformer blocks (for the hoisted K/V buffers)
  int nt = 0;
  for (auto& b : u->in_blocks) nt += (int)b.st.blocks.size();
  nt += (int)u->mid_st.blocks.size();
  for (auto& b : u->out_blocks) nt += (int)b.st.blocks.size();
  u->n_tblocks = nt;
  return 0;
}

// ================================================================================================
// load
// ================================================================================================
static int parse_pack(sdxl_ctx* c, const void* pack, size_t bytes, int on_device, PackView& pv,
                      std::vector<uint8_t>& host_table) {
  if (bytes < sizeof(PackHeader)) return fail(c, 4100, "weight pack too small");
  PackHeader h;
  if (on_device) CU(c, cudaMemcpy(&h, pack, sizeof h, cudaMemcpyDeviceToHost));
  else memcpy(&h, pack, sizeof h);
  if (memcmp(h.magic, "SDXLPK01", 8) != 0) return fail(c, 4101, "weight pack: bad magic");
  const size_t tbytes = (size_t)h.n_tensors * sizeof(PackEntry);
  if (sizeof h + tbytes > bytes) return fail(c, 4102, "weight pack: truncated table");
  host_table.resize(tbytes);
  if (on_device) CU(c, cudaMemcpy(host_table.data(), (const uint8_t*)pack + sizeof h, tbytes, cudaMemcpyDeviceToHost));
  else memcpy(host_table.data(), (const uint8_t*)pack + sizeof h, tbytes);
  const PackEntry* e = (const PackEntry*)host_table.data();
  for (uint32_t i = 0; i < h.n_tensors; ++i) {
    if (e[i].offset + e[i].nbytes > bytes) return fail(c, 4103, "weight pack: tensor '%.*s' out of range", 119, e[i].name);
    if (e[i].offset % 16) return fail(c, 4104, "weight pack: tensor '%.*s' not 16B aligned", 119, e[i].name);
    std::string name(e[i].name, strnlen(e[i].name, sizeof e[i].name));
    pv.t[name] = e[i];
  }
  return 0;
}

extern "C" void sdxl_unet_destroy(sdxl_unet* u);

extern "C" int sdxl_unet_load(sdxl_ctx* c, const sdxl_unet_cfg* cfg, const void* pack, size_t bytes, int pack_on_device,
                              sdxl_unet** out) {
  if (!c || !cfg || !pack || !out) return fail(c, -1, "sdxl_unet_load: null argument");
  *out = nullptr;
  if (cfg->n_head_channels != 64) return fail(c, 4200, "n_head_channels must be 64 (got %d)", cfg->n_head_channels);
  if (cfg->n_levels < 1 || cfg->n_levels > SDXL_MAX_LEVELS) return fail(c, 4201, "bad n_levels");
  if (cfg->in_channels > 8 || cfg->model_channels % 32) return fail(c, 4202, "unsupported channel config");
  CU(c, cudaSetDevice(c->device));
  std::unique_ptr<sdxl_unet> u(new sdxl_unet());
  u->ctx = c;
  u->cfg = *cfg;
  PackView pv;
  std::vector<uint8_t> table;
  int r = parse_pack(c, pack, bytes, pack_on_device, pv, table);
  if (r) return r;
  void* dev_pack = nullptr;
  if (pack_on_device) {
    pv.dev = (const uint8_t*)pack;
  } else {
    CU(c, cudaMalloc(&dev_pack, bytes));
    cudaError_t e = cudaMemcpyAsync(dev_pack, pack, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { cudaFree(dev_pack); return fail(c, (int)e, "pack upload failed"); }
    pv.dev = (const uint8_t*)dev_pack;
  }
  // pass 1: measure, pass 2: build
  Arena meas;
  meas.measure = true;
  r = build_model(u.get(), pv, meas);
  if (!r) {
    if (u->warena.init(meas.off + (1 << 20))) r = fail(c, 4203, "cannot allocate %zu bytes for weights", meas.off);
  }
  if (!r) r = build_model(u.get(), pv, u->warena);
  // alphas_cumprod: f16-stored in the reference's record (HalfPrecisionSettings), read as f64 (mod.rs:485-492)
  if (!r) {
    const PackEntry* e = pv.find("alphas_cumprod");
    if (!e || e->ndim != 1 || e->dtype != 0) r = fail(c, 4204, "weight pack: missing f16 'alphas_cumprod'");
    else {
      std::vector<uint16_t> raw(e->shape[0]);
      cudaError_t ce = cudaMemcpyAsync(raw.data(), pv.dev + e->offset, raw.size() * 2, cudaMemcpyDeviceToHost, c->stream);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
      if (ce != cudaSuccess) r = fail(c, (int)ce, "alphas download failed");
      else {
        u->alphas.resize(raw.size());
        for (size_t i = 0; i < raw.size(); ++i) {
          __half_raw hr;
          hr.x = raw[i];
          u->alphas[i] = (double)__half2float(__half(hr));
        }
      }
    }
  }
  cudaError_t se = cudaStreamSynchronize(c->stream);
  if (dev_pack) cudaFree(dev_pack);
  if (!r && se != cudaSuccess) r = fail(c, (int)se, "weight re-layout failed: %s", cudaGetErrorString(se));
  if (r) return r;
  CU(c, cudaMalloc((void**)&u->t_dev, 64));
  CU(c, cudaMallocHost((void**)&u->t_pinned, 4096 * sizeof(int)));
  *out = u.release();
  return 0;
}

// ================================================================================================
// launch plan
// ================================================================================================
enum OpKind { OP_IGEMM, OP_ATTN, OP_GN, OP_LN, OP_GEMV, OP_TEMB, OP_CONV_IN, OP_UPS, OP_PHASE, OP_CAST16,
              OP_SOFTMAX, OP_TRANSPOSE, OP_PQ, OP_EMBED, OP_ATTN_SMALL, OP_ACT, OP_LN_GATHER };
static const char* const kOpNames[] = {"igemm", "attention", "group_norm", "layer_norm", "gemv", "temb", "conv_in", "upsample",
                                       "phase_split", "cast16", "softmax_rows", "transpose16", "post_quant", "embed_tokens",
                                       "attention_small", "mlp_act", "ln_gather"};
struct Op {
  OpKind kind;
  double flops = 0;  // algorithmic FLOPs of this launch (igemm / attention), 0 for HBM-bound ops
  IgemmParams ig;
  AttnParams at;
  GnParams gn;
  struct { const float* x; const float* g; const float* b; float eps; int rows, C; __half* y; } ln;
  struct { const float* in; int in_bstride, Bv, K; const __half* W; int ldw; const float* bias; const float* add; int add_bstride, N, in_silu, out_silu; float* out; int out_bstride; } gv;
  struct { const int* t; int n, dim; float* out; } te;
  struct { const float* x; int Bx, B, Cin, H, W; const float* w; const float* bias; int Cout; float* y; } ci;
  struct { const float* x; int B, H, W, C; __half* y; } rs;  // upsample / phase split
  struct { const float* x; size_t n; __half* y; } cs;
  struct { const float* S; size_t lds; int rows, cols; float scale; __half* P; size_t ldp; } sm;
  struct { const __half* x; size_t ldx; int rows, cols; __half* y; size_t ldy; } tr;
  struct { const float* x; int B, C, HW; const float* w; const float* bias; float inv_scale; float* y; } pq;
  struct { const int* tokens; int rows, T, C, n_vocab; const __half* tok; const __half* pos; float* x; int* err; } em;
  struct { const __half* q; int q_pitch, q_col0; const __half* k; const __half* v; int kv_pitch, k_col0, v_col0, B, T, S, n_head;
           const __half* mask; int causal; __half* out; int ldo; } as;
  struct { const float* x; size_t n; int quick; __half* y; } ac;
  struct { const float* x; const int* idx; int B, T, C; const float* g; const float* b; float eps; float* y; } lg;
};

struct Plan {
  int Bf = 0, Bx = 0, h = 0, w = 0;
  uint64_t cond_version = 0;
  Arena arena;
  std::vector<Op> ops;
  float* x_in = nullptr;  // [Bx, Cin, h, w] f32 NCHW
  float* eps = nullptr;   // [Bf, h*w, eps_ld] f32 NHWC
  int eps_ld = 4;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int runs = 0;
  double flops = 0;  // algorithmic FLOPs of one run (2*MAC over Linear/conv/attention)
  ~Plan() {
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    arena.release();
  }
};

struct ActView { const __half* p; int Bn, H, W, C; };
struct F32View { float* p; int C; };  // [Bf, HW, C]

struct PlanBuilder {
  sdxl_unet* u;
  sdxl_ctx* c;
  Plan* P;
  Arena* A;
  int Bf;
  int err = 0;
  // shared scratch
  float* gn_partial = nullptr;
  int kv_index = 0;

  template <typename T>
  T* buf(size_t n) {
    T* p = A->get<T>(n);
    if (!p && !err) err = fail(c, 5001, "plan arena exhausted");
    return p;
  }
  void add_flops(double f) {  // attribute to the op just pushed
    if (err || P->ops.empty()) return;
    P->ops.back().flops += f;
    P->flops += f;
  }
  // generic igemm op; segs reference view a0 (map 0) / a1 (map 1)
  void igemm(const ActView& a0, const ActView* a1, const std::vector<IgemmSeg>& segs, const __half* W, int N, int Ktot,
             int outH, int outW, int outB, int mode, int geglu_bn, void* out, int out_f32, int ldo, const float* bias,
             int bias_bstride, const float* res, int ldr) {
    if (err) return;
    Op op{};
    op.kind = OP_IGEMM;
    IgemmParams& p = op.ig;
    p.nseg = (int)segs.size();
    if (p.nseg > IGEMM_MAX_SEG) { err = fail(c, 5002, "too many igemm segments"); return; }
    for (int i = 0; i < p.nseg; ++i) p.seg[i] = segs[i];
    p.out = out; p.out_f32 = out_f32; p.ldo = ldo;
    p.bias = bias; p.bias_bstride = bias_bstride;
    p.res = res; p.ldr = ldr;
    if (!A->measure) {
      IgemmOperands o{a0.p, a0.Bn, a0.H, a0.W, a0.C, a0.C, a1 ? a1->p : nullptr, a1 ? a1->Bn : 0, a1 ? a1->H : 0,
                      a1 ? a1->W : 0, a1 ? a1->C : 0, a1 ? a1->C : 0, W, N, Ktot};
      int r = igemm_configure(p, o, outW, outH, outB, mode, geglu_bn);
      if (r) { err = fail(c, r, "igemm configuration failed (N=%d K=%d)", N, Ktot); return; }
    }
    P->ops.push_back(op);
  }
  void linear(const __half* x, int M, const Lin& L, int mode, void* out, int out_f32, int ldo, const float* res, int ldr) {
    ActView a{x, 1, 1, M, L.K};
    std::vector<IgemmSeg> segs{{0, 0, 0, 0, L.Kpad / 64}};
    igemm(a, nullptr, segs, L.w, L.N, L.Kpad, 1, M, 1, mode, L.geglu_bn, out, out_f32, ldo, L.b, 0, res, ldr);
    add_flops(2.0 * M * (double)L.K * L.N);
  }
  // 3x3 stride-1 conv (+ optional fused 1x1 skip segment on a1)
  void conv3(const ActView& a, const ActView* skip, const Conv& cv, float* out, const float* bias, int bias_bstride,
             const float* res) {
    std::vector<IgemmSeg> segs;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) segs.push_back({0, (int16_t)(kw - 1), (int16_t)(kh - 1), 0, cv.Ipad / 64});
    if (skip) segs.push_back({1, 0, 0, 0, cv.I2pad / 64});
    igemm(a, skip, segs, cv.w, cv.O, cv.Ktot, a.H, a.W, a.Bn, IGEMM_LINEAR, 0, out, 1, cv.O, bias, bias_bstride, res, cv.O);
    add_flops(2.0 * a.Bn * a.H * a.W * (double)cv.O * (9.0 * cv.I + cv.I2));
  }
  void gn(const float* x1, int C1, const float* x2, int C2, int HW, const Norm& n, int silu, __half* y, __half* raw) {
    if (err) return;
    Op op{};
    op.kind = OP_GN;
    op.gn = GnParams{x1, C1, x2, C2, Bf, HW, 32, n.g, n.b, n.eps, silu, y, raw, gn_partial, 0};
    P->ops.push_back(op);
  }
  void ln(const float* x, const Norm& n, int rows, __half* y) {
    if (err) return;
    Op op{};
    op.kind = OP_LN;
    op.ln = {x, n.g, n.b, n.eps, rows, n.C, y};
    P->ops.push_back(op);
  }
  void gemv(const float* in, int in_bstride, int Bv, const Lin& L, const float* add, int add_bstride, int in_silu,
            int out_silu, float* out, int out_bstride) {
    if (err) return;
    for (int b0 = 0; b0 < Bv; b0 += 8) {
      Op op{};
      op.kind = OP_GEMV;
      const int nb = Bv - b0 < 8 ? Bv - b0 : 8;
      op.gv = {in + (size_t)b0 * in_bstride, in_bstride, nb, L.K, L.w, L.Kpad, L.b, add ? add + (size_t)b0 * add_bstride : nullptr,
               add_bstride, L.N, in_silu, out_silu, out + (size_t)b0 * out_bstride, out_bstride};
      P->ops.push_back(op);
    }
    P->flops += 2.0 * Bv * (double)L.K * L.N;
  }

  // ---- ResBlock (reference unet/mod.rs:1082-1106) ----
  float* resblock(const Res& r, const float* xa, int Ca, const float* xb, int Cb, int H, int W, const float* temb_all,
                  int temb_total, __half* s_gn1, __half* s_raw, float* s_h, __half* s_gn2) {
    const int HW = H * W;
    float* out = buf<float>((size_t)Bf * HW * r.Cout);
    gn(xa, Ca, xb, Cb, HW, r.n_in, 1, s_gn1, r.has_skip ? s_raw : nullptr);
    ActView a1{s_gn1, Bf, H, W, r.Cin};
    // h = conv_in(silu(gn(x))) + b + lin_embed(silu(emb))[:, :, None, None]   (bias folded into temb_all)
    conv3(a1, nullptr, r.conv_in, s_h, temb_all + r.temb_off, temb_total, nullptr);
    gn(s_h, r.Cout, nullptr, 0, HW, r.n_out, 1, s_gn2, nullptr);
    ActView a2{s_gn2, Bf, H, W, r.Cout};
    if (r.has_skip) {
      ActView sk{s_raw, Bf, H, W, r.Cin};
      conv3(a2, &sk, r.conv_out, out, r.conv_out.b, 0, nullptr);  // skip 1x1 conv fused as a K segment
    } else {
      conv3(a2, nullptr, r.conv_out, out, r.conv_out.b, 0, xa);   // identity residual in the epilogue
    }
    return out;
  }

  // ---- SpatialTransformer (reference unet/mod.rs:820-845, 885-891, 1005-1023) ----
  float* strans(const STrans& s, const float* x, int H, int W, __half* s_a16, float* s_tok, __half* s_qkv, __half* s_ao,
                __half* s_q, __half* s_ff) {
    const int T = H * W, M = Bf * T, C = s.C;
    float* out = buf<float>((size_t)M * C);
    gn(x, C, nullptr, 0, T, s.norm, 0, s_a16, nullptr);
    linear(s_a16, M, s.proj_in, IGEMM_LINEAR, s_tok, 1, C, nullptr, 0);
    const float sl2e = (float)(1.4426950408889634 / sqrt(64.0));
    for (const TBlock& b : s.blocks) {
      // x = x + attn1(norm1(x))
      ln(s_tok, b.n1, M, s_a16);
      linear(s_a16, M, b.qkv, IGEMM_LINEAR, s_qkv, 0, 3 * C, nullptr, 0);
      attn(s_qkv, 3 * C, 0, s_qkv, 3 * C, C, 2 * C, T, T, s.n_head, s_ao, C, sl2e);
      linear(s_ao, M, b.out1, IGEMM_LINEAR, s_tok, 1, C, s_tok, C);
      // x = x + attn2(norm2(x), context)   (K/V hoisted to set_conditioning)
      ln(s_tok, b.n2, M, s_a16);
      linear(s_a16, M, b.q2, IGEMM_LINEAR, s_q, 0, C, nullptr, 0);
      const __half* kvp = A->measure ? nullptr : u->kv[kv_index];
      attn(s_q, C, 0, kvp, 2 * C, 0, C, T, u->n_ctx, s.n_head, s_ao, C, sl2e);
      P->flops += 2.0 * Bf * u->n_ctx * (double)b.kv2.K * b.kv2.N;  // hoisted K/V projections (algorithmic work)
      kv_index++;
      linear(s_ao, M, b.out2, IGEMM_LINEAR, s_tok, 1, C, s_tok, C);
      // x = x + mlp(norm3(x))
      ln(s_tok, b.n3, M, s_a16);
      linear(s_a16, M, b.ff1, IGEMM_GEGLU, s_ff, 0, 4 * C, nullptr, 0);
      linear(s_ff, M, b.ff2, IGEMM_LINEAR, s_tok, 1, C, s_tok, C);
    }
    // proj_out(tokens) + x_in  (f32 stream -> f16 operand)
    if (!err) {
      Op op{};
      op.kind = OP_CAST16;
      op.cs = {s_tok, (size_t)M * C, s_a16};
      P->ops.push_back(op);
    }
    linear(s_a16, M, s.proj_out, IGEMM_LINEAR, out, 1, C, x, C);
    return out;
  }
  void attn(const __half* qm, int q_pitch, int q_col0, const __half* kvm, int kv_pitch, int k_col0, int v_col0, int T,
            int S, int n_head, __half* out, int ldo, float sl2e) {
    if (err) return;
    Op op{};
    op.kind = OP_ATTN;
    AttnParams& p = op.at;
    p.T = T; p.S = S; p.n_head = n_head; p.B = Bf;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.out = out; p.ldo = ldo; p.scale_log2e = sl2e;
    if (!A->measure) {
      int r = make_tmap_rows(&p.tmQ, qm, T, Bf, q_pitch, q_pitch);
      if (!r) r = make_tmap_rows(&p.tmK, kvm, S, Bf, kv_pitch, kv_pitch);
      if (!r) p.tmV = p.tmK;
      if (r) { err = fail(c, r, "tensor map creation failed (attention)"); return; }
    }
    P->ops.push_back(op);
    add_flops(4.0 * Bf * T * (double)S * (n_head * 64));
  }
};

static int exec_op(sdxl_ctx* c, Op& op) {
  cudaStream_t st = c->stream;
  switch (op.kind) {
    case OP_IGEMM: KL(c, igemm_launch(st, op.ig)); break;
    case OP_ATTN: KL(c, attention_launch(st, op.at)); break;
    case OP_GN: KL(c, gn_launch(st, op.gn)); c->launches++; break;
    case OP_LN: KL(c, layernorm_launch(st, op.ln.x, op.ln.g, op.ln.b, op.ln.eps, op.ln.rows, op.ln.C, op.ln.y)); break;
    case OP_GEMV:
      KL(c, gemv_launch(st, op.gv.in, op.gv.in_bstride, op.gv.Bv, op.gv.K, op.gv.W, op.gv.ldw, op.gv.bias, op.gv.add, op.gv.add_bstride,
                        op.gv.N, op.gv.in_silu, op.gv.out_silu, op.gv.out, op.gv.out_bstride));
      break;
    case OP_TEMB: KL(c, timestep_embedding_launch(st, op.te.t, op.te.n, op.te.dim, 10000.f, op.te.out)); break;
    case OP_CONV_IN:
      KL(c, conv_in_launch_t(st, op.ci.x, 1, op.ci.Bx, op.ci.B, op.ci.Cin, op.ci.H, op.ci.W, op.ci.w, op.ci.bias, op.ci.Cout, op.ci.y));
      break;
    case OP_UPS: KL(c, upsample2x_launch(st, op.rs.x, op.rs.B, op.rs.H, op.rs.W, op.rs.C, op.rs.y)); break;
    case OP_PHASE: KL(c, phase_split_launch(st, op.rs.x, op.rs.B, op.rs.H, op.rs.W, op.rs.C, op.rs.y)); break;
    case OP_CAST16: KL(c, cast_f32_to_f16_launch(st, op.cs.x, op.cs.n, op.cs.y)); break;
    case OP_SOFTMAX: KL(c, softmax_rows_launch(st, op.sm.S, op.sm.lds, op.sm.rows, op.sm.cols, op.sm.scale, op.sm.P, op.sm.ldp)); break;
    case OP_TRANSPOSE: KL(c, transpose_f16_launch(st, op.tr.x, op.tr.ldx, op.tr.rows, op.tr.cols, op.tr.y, op.tr.ldy)); break;
    case OP_EMBED: KL(c, embed_tokens_launch(st, op.em.tokens, op.em.rows, op.em.T, op.em.C, op.em.n_vocab, op.em.tok, op.em.pos, op.em.x, op.em.err)); break;
    case OP_ATTN_SMALL:
      KL(c, attention_small_launch(st, op.as.q, op.as.q_pitch, op.as.q_col0, op.as.k, op.as.v, op.as.kv_pitch, op.as.k_col0, op.as.v_col0,
                                   op.as.B, op.as.T, op.as.S, op.as.n_head, op.as.mask, op.as.causal, op.as.out, op.as.ldo));
      break;
    case OP_ACT: KL(c, mlp_act_launch(st, op.ac.x, op.ac.n, op.ac.quick, op.ac.y)); break;
    case OP_LN_GATHER: KL(c, ln_gather_f32_launch(st, op.lg.x, op.lg.idx, op.lg.B, op.lg.T, op.lg.C, op.lg.g, op.lg.b, op.lg.eps, op.lg.y)); break;
    case OP_PQ: KL(c, post_quant_launch(st, op.pq.x, op.pq.B, op.pq.C, op.pq.HW, op.pq.w, op.pq.bias, op.pq.inv_scale, op.pq.y)); break;
  }
  return 0;
}

// Builds the op list for UNet::forward (reference unet/mod.rs:449-493) at batch Bf, latent h x w.
static int build_plan_ops(sdxl_unet* u, Plan* P, Arena* A) {
  sdxl_ctx* c = u->ctx;
  const sdxl_unet_cfg& g = u->cfg;
  PlanBuilder B{u, c, P, A, P->Bf};
  P->ops.clear();
  P->flops = 0;
  const int Bf = P->Bf, mc = g.model_channels, ted = 4 * mc;
  const int temb_total = u->temb_all.N;
  const int levels = g.n_levels;
  if ((P->h % (1 << (levels - 1))) || (P->w % (1 << (levels - 1)))) return fail(c, 5003, "latent %dx%d not divisible by %d", P->h, P->w, 1 << (levels - 1));

  P->x_in = B.buf<float>((size_t)P->Bx * g.in_channels * P->h * P->w);
  P->eps_ld = g.out_channels;
  P->eps = B.buf<float>((size_t)Bf * P->h * P->w * P->eps_ld);
  B.gn_partial = B.buf<float>(gn_scratch_floats(Bf, 32));
  float* te = B.buf<float>(mc);
  float* t1 = B.buf<float>(ted);
  float* semb = B.buf<float>((size_t)Bf * ted);
  float* temb_all = B.buf<float>((size_t)Bf * temb_total);

  // maxima for the shared scratch buffers
  size_t max_pixC_cat = 0, max_pixC = 0, max_tokC = 0;
  {
    int H = P->h, W = P->w;
    auto upd = [&](const Res& r, int hh, int ww) {
      max_pixC_cat = std::max(max_pixC_cat, (size_t)hh * ww * r.Cin);
      max_pixC = std::max(max_pixC, (size_t)hh * ww * r.Cout);
    };
    for (auto& b : u->in_blocks) {
      if (b.type == BT_RES || b.type == BT_REST) upd(b.res, H, W);
      if (b.type == BT_REST) max_tokC = std::max(max_tokC, (size_t)H * W * b.st.C);
      if (b.type == BT_DOWN) { H /= 2; W /= 2; }
    }
    upd(u->mid_res1, H, W);
    upd(u->mid_res2, H, W);
    max_tokC = std::max(max_tokC, (size_t)H * W * u->mid_st.C);
    for (auto& b : u->out_blocks) {
      upd(b.res, H, W);
      if (b.type == BT_REST || b.type == BT_RESTU) max_tokC = std::max(max_tokC, (size_t)H * W * b.st.C);
      if (b.type == BT_RESTU || b.type == BT_RESU) { H *= 2; W *= 2; }
    }
  }
  __half* s_gn1 = B.buf<__half>(Bf * max_pixC_cat);
  __half* s_raw = B.buf<__half>(Bf * max_pixC_cat);
  float* s_h = B.buf<float>(Bf * max_pixC);
  __half* s_gn2 = B.buf<__half>(Bf * max_pixC);
  __half* s_a16 = B.buf<__half>(Bf * max_tokC);
  float* s_tok = B.buf<float>(Bf * max_tokC);
  __half* s_qkv = B.buf<__half>(Bf * max_tokC * 3);
  __half* s_ao = B.buf<__half>(Bf * max_tokC);
  __half* s_q = B.buf<__half>(Bf * max_tokC);
  __half* s_ff = B.buf<__half>(Bf * max_tokC * 4);
  if (B.err) return B.err;

  // --- embeddings (unet/mod.rs:458-468): emb = time_mlp(temb(t)) + label_emb; only SiLU(emb) is consumed
  {
    Op op{};
    op.kind = OP_TEMB;
    op.te = {u->t_dev, 1, mc, te};
    P->ops.push_back(op);
  }
  B.gemv(te, 0, 1, u->t1, nullptr, 0, 0, 1, t1, 0);
  B.gemv(t1, 0, Bf, u->t2, u->label_emb, ted, 0, 1, semb, ted);
  B.gemv(semb, ted, Bf, u->temb_all, nullptr, 0, 0, 0, temb_all, temb_total);

  // --- input blocks
  struct Saved { float* p; int C, H, W; };
  std::vector<Saved> saved;
  int H = P->h, W = P->w;
  float* x = B.buf<float>((size_t)Bf * H * W * mc);
  int Cx = mc;
  {
    Op op{};
    op.kind = OP_CONV_IN;
    op.ci = {P->x_in, P->Bx, Bf, g.in_channels, H, W, u->conv0_w, u->conv0_b, mc, x};
    P->ops.push_back(op);
    P->flops += 2.0 * Bf * H * W * 9.0 * g.in_channels * mc;
  }
  saved.push_back({x, Cx, H, W});
  for (size_t i = 1; i < u->in_blocks.size() && !B.err; ++i) {
    const Block& b = u->in_blocks[i];
    if (b.type == BT_RES || b.type == BT_REST) {
      x = B.resblock(b.res, x, Cx, nullptr, 0, H, W, temb_all, temb_total, s_gn1, s_raw, s_h, s_gn2);
      Cx = b.res.Cout;
      if (b.type == BT_REST) x = B.strans(b.st, x, H, W, s_a16, s_tok, s_qkv, s_ao, s_q, s_ff);
    } else if (b.type == BT_DOWN) {
      // 3x3 stride 2 pad 1 (unet/mod.rs:760-774) on phase-split input: tap kh -> (phase, offset)
      __half* ph = B.buf<__half>((size_t)Bf * H * W * Cx);
      Op op{};
      op.kind = OP_PHASE;
      op.rs = {x, Bf, H, W, Cx, ph};
      P->ops.push_back(op);
      const int H2 = H / 2, W2 = W / 2;
      ActView a{ph, 4 * Bf, H2, W2, Cx};
      std::vector<IgemmSeg> segs;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          const int phh = (kh == 1) ? 0 : 1, pw = (kw == 1) ? 0 : 1;
          const int dh = (kh == 0) ? -1 : 0, dw = (kw == 0) ? -1 : 0;
          segs.push_back({0, (int16_t)dw, (int16_t)dh, (int16_t)((phh * 2 + pw) * Bf), b.conv.Ipad / 64});
        }
      float* y = B.buf<float>((size_t)Bf * H2 * W2 * Cx);
      B.igemm(a, nullptr, segs, b.conv.w, b.conv.O, b.conv.Ktot, H2, W2, Bf, IGEMM_LINEAR, 0, y, 1, b.conv.O, b.conv.b, 0, nullptr, 0);
      B.add_flops(2.0 * Bf * H2 * W2 * 9.0 * Cx * b.conv.O);
      x = y; H = H2; W = W2;
    }
    saved.push_back({x, Cx, H, W});
  }
  // --- middle
  x = B.resblock(u->mid_res1, x, Cx, nullptr, 0, H, W, temb_all, temb_total, s_gn1, s_raw, s_h, s_gn2);
  x = B.strans(u->mid_st, x, H, W, s_a16, s_tok, s_qkv, s_ao, s_q, s_ff);
  x = B.resblock(u->mid_res2, x, Cx, nullptr, 0, H, W, temb_all, temb_total, s_gn1, s_raw, s_h, s_gn2);
  // --- output blocks: cat([x, saved.pop()], channel) is never materialised (GN + skip conv read both)
  for (size_t i = 0; i < u->out_blocks.size() && !B.err; ++i) {
    const Block& b = u->out_blocks[i];
    if (saved.empty()) return fail(c, 5004, "skip stack underflow");
    Saved sk = saved.back();
    saved.pop_back();
    if (sk.H != H || sk.W != W || Cx + sk.C != b.res.Cin) return fail(c, 5005, "skip shape mismatch at output block %zu", i);
    x = B.resblock(b.res, x, Cx, sk.p, sk.C, H, W, temb_all, temb_total, s_gn1, s_raw, s_h, s_gn2);
    Cx = b.res.Cout;
    if (b.type == BT_REST || b.type == BT_RESTU) x = B.strans(b.st, x, H, W, s_a16, s_tok, s_qkv, s_ao, s_q, s_ff);
    if (b.type == BT_RESTU || b.type == BT_RESU) {
      // nearest-2x then 3x3 conv (unet/mod.rs:742-751)
      __half* up = B.buf<__half>((size_t)Bf * 4 * H * W * Cx);
      Op op{};
      op.kind = OP_UPS;
      op.rs = {x, Bf, H, W, Cx, up};
      P->ops.push_back(op);
      H *= 2; W *= 2;
      ActView a{up, Bf, H, W, Cx};
      float* y = B.buf<float>((size_t)Bf * H * W * Cx);
      B.conv3(a, nullptr, b.conv, y, b.conv.b, 0, nullptr);
      x = y;
    }
  }
  if (B.err) return B.err;
  // --- head: GN -> SiLU -> conv 3x3 (unet/mod.rs:488-490)
  B.gn(x, Cx, nullptr, 0, H * W, u->norm_out, 1, s_gn1, nullptr);
  {
    ActView a{s_gn1, Bf, H, W, Cx};
    std::vector<IgemmSeg> segs;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) segs.push_back({0, (int16_t)(kw - 1), (int16_t)(kh - 1), 0, u->conv_out.Ipad / 64});
    B.igemm(a, nullptr, segs, u->conv_out.w, u->conv_out.O, u->conv_out.Ktot, H, W, Bf, IGEMM_LINEAR, 0, P->eps, 1, P->eps_ld,
            u->conv_out.b, 0, nullptr, 0);
    B.add_flops(2.0 * Bf * H * W * 9.0 * Cx * u->conv_out.O);
  }
  return B.err;
}

static int ensure_plan(sdxl_unet* u, int Bf, int Bx, int h, int w) {
  sdxl_ctx* c = u->ctx;
  if (u->condB != Bf) return fail(c, 5010, "conditioning is set for batch %d but forward batch is %d (call sdxl_unet_set_conditioning first)", u->condB, Bf);
  if (u->plan && u->plan->Bf == Bf && u->plan->Bx == Bx && u->plan->h == h && u->plan->w == w && u->plan->cond_version == u->cond_version)
    return 0;
  CU(c, cudaStreamSynchronize(c->stream));
  u->plan.reset(new Plan());
  Plan* P = u->plan.get();
  P->Bf = Bf; P->Bx = Bx; P->h = h; P->w = w; P->cond_version = u->cond_version;
  Arena meas;
  meas.measure = true;
  int r = build_plan_ops(u, P, &meas);
  if (r) { u->plan.reset(); return r; }
  if (P->arena.init(meas.off + (1 << 20))) { u->plan.reset(); return fail(c, 5011, "cannot allocate %zu bytes of workspace", meas.off); }
  r = build_plan_ops(u, P, &P->arena);
  if (r) { u->plan.reset(); return r; }
  return 0;
}

static int run_plan_ops(sdxl_ctx* c, Plan* P) {
  static const bool no_graph = getenv("SDXL_B200_NO_GRAPH") != nullptr;
  if (P->gexec) {
    CU(c, cudaGraphLaunch(P->gexec, c->stream));
    c->launches += P->ops.size() + [&] { size_t g = 0; for (auto& o : P->ops) g += o.kind == OP_GN; return g; }();
    return 0;
  }
  const bool capture = !no_graph && P->runs >= 1;  // first run eager (sets func attributes), then capture
  if (capture) CU(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  int r = 0;
  for (auto& op : P->ops) {
    r = exec_op(c, op);
    if (r) break;
  }
  if (capture) {
    cudaGraph_t gph = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &gph);
    if (r) { if (gph) cudaGraphDestroy(gph); return r; }
    if (e != cudaSuccess) return fail(c, (int)e, "graph capture failed: %s", cudaGetErrorString(e));
    P->graph = gph;
    e = cudaGraphInstantiate(&P->gexec, gph, 0);
    if (e != cudaSuccess) { P->gexec = nullptr; return fail(c, (int)e, "graph instantiate failed: %s", cudaGetErrorString(e)); }
    CU(c, cudaGraphLaunch(P->gexec, c->stream));
  }
  P->runs++;
  return r;
}
static int run_plan(sdxl_unet* u) { return run_plan_ops(u->ctx, u->plan.get()); }

static int set_t(sdxl_unet* u, int t) {
  sdxl_ctx* c = u->ctx;
  static int slot = 0;
  slot = (slot + 1) % 4096;
  u->t_pinned[slot] = t;
  CU(c, cudaMemcpyAsync(u->t_dev, &u->t_pinned[slot], sizeof(int), cudaMemcpyHostToDevice, c->stream));
  return 0;
}

// ================================================================================================
// conditioning (step-invariant work hoisted out of UNet::forward)
// ================================================================================================
static int set_conditioning_dev(sdxl_unet* u, int B, int n_ctx, const __half* context_dev, const __half* y_dev) {
  sdxl_ctx* c = u->ctx;
  const sdxl_unet_cfg& g = u->cfg;
  const int ted = 4 * g.model_channels;
  if (B < 1 || n_ctx < 1) return fail(c, 5100, "bad conditioning shape");
  if (u->condB != B || u->n_ctx != n_ctx) {
    CU(c, cudaStreamSynchronize(c->stream));
    u->plan.reset();
    // collect transformer blocks
    std::vector<const TBlock*> tbs;
    for (auto& b : u->in_blocks) for (auto& t : b.st.blocks) tbs.push_back(&t);
    for (auto& t : u->mid_st.blocks) tbs.push_back(&t);
    for (auto& b : u->out_blocks) for (auto& t : b.st.blocks) tbs.push_back(&t);
    u->ctx_pitch = (g.context_dim + 7) / 8 * 8;
    size_t need = 0;
    auto al = [&](size_t b) { need = ((need + 1023) & ~size_t(1023)) + b; };
    al((size_t)B * n_ctx * u->ctx_pitch * 2);
    al((size_t)B * g.adm_in_channels * 4);
    al((size_t)B * ted * 4);
    al((size_t)B * ted * 4);
    for (auto* t : tbs) al((size_t)B * n_ctx * t->kv2.N * 2);
    if (u->carena.init(need + (1 << 16))) return fail(c, 5101, "cannot allocate conditioning buffers");
    u->ctx16 = u->carena.get<__half>((size_t)B * n_ctx * u->ctx_pitch);
    u->y32 = u->carena.get<float>((size_t)B * g.adm_in_channels);
    u->lab1 = u->carena.get<float>((size_t)B * ted);
    u->label_emb = u->carena.get<float>((size_t)B * ted);
    u->kv.clear();
    u->kvC.clear();
    for (auto* t : tbs) {
      u->kv.push_back(u->carena.get<__half>((size_t)B * n_ctx * t->kv2.N));
      u->kvC.push_back(t->kv2.N / 2);
    }
    u->condB = B;
    u->n_ctx = n_ctx;
    CU(c, cudaMemsetAsync(u->ctx16, 0, (size_t)B * n_ctx * u->ctx_pitch * 2, c->stream));
  }
  u->cond_version++;
  if (u->plan) u->plan->cond_version = u->cond_version;  // buffers unchanged: plan stays valid
  CU(c, cudaMemcpy2DAsync(u->ctx16, (size_t)u->ctx_pitch * 2, context_dev, (size_t)g.context_dim * 2, (size_t)g.context_dim * 2,
                          (size_t)B * n_ctx, cudaMemcpyDeviceToDevice, c->stream));
  KL(c, cast_f16_to_f32_launch(c->stream, y_dev, (size_t)B * g.adm_in_channels, u->y32));
  // label_emb = lin2(SiLU(lin1(y)))   (unet/mod.rs:464-466)
  for (int b0 = 0; b0 < B; b0 += 8) {
    const int nb = B - b0 < 8 ? B - b0 : 8;
    KL(c, gemv_launch(c->stream, u->y32 + (size_t)b0 * g.adm_in_channels, g.adm_in_channels, nb, g.adm_in_channels, u->l1.w,
                      u->l1.Kpad, u->l1.b, nullptr, 0, ted, 0, 1, u->lab1 + (size_t)b0 * ted, ted));
    KL(c, gemv_launch(c->stream, u->lab1 + (size_t)b0 * ted, ted, nb, u->l2.K, u->l2.w, u->l2.Kpad, u->l2.b, nullptr, 0, ted, 0, 0,
                      u->label_emb + (size_t)b0 * ted, ted));
  }
  // K/V projections of the context for every cross-attention (unet/mod.rs:1010-1011)
  {
    std::vector<const TBlock*> tbs;
    for (auto& b : u->in_blocks) for (auto& t : b.st.blocks) tbs.push_back(&t);
    for (auto& t : u->mid_st.blocks) tbs.push_back(&t);
    for (auto& b : u->out_blocks) for (auto& t : b.st.blocks) tbs.push_back(&t);
    const int M = B * n_ctx;
    for (size_t i = 0; i < tbs.size(); ++i) {
      const Lin& L = tbs[i]->kv2;
      IgemmParams p{};
      p.nseg = 1;
      p.seg[0] = {0, 0, 0, 0, L.Kpad / 64};
      p.out = u->kv[i]; p.out_f32 = 0; p.ldo = L.N;
      IgemmOperands o{u->ctx16, 1, 1, M, g.context_dim, u->ctx_pitch, nullptr, 0, 0, 0, 0, 0, L.w, L.N, L.Kpad};
      int r = igemm_configure(p, o, M, 1, 1, IGEMM_LINEAR, 0);
      if (r) return fail(c, r, "igemm configuration failed (kv projection)");
      KL(c, igemm_launch(c->stream, p));
    }
  }
  return 0;
}

extern "C" int sdxl_unet_set_conditioning(sdxl_unet* u, int B, int n_ctx, const sdxl_half* context, const sdxl_half* y) {
  if (!u || !context || !y) return -1;
  CU(u->ctx, cudaSetDevice(u->ctx->device));
  return set_conditioning_dev(u, B, n_ctx, (const __half*)context, (const __half*)y);
}

// ================================================================================================
// UNet::forward
// ================================================================================================
extern "C" int sdxl_unet_forward(sdxl_unet* u, int B, int h, int w, const sdxl_half* x, int32_t t_host, sdxl_half* eps_out) {
  if (!u || !x || !eps_out) return -1;
  sdxl_ctx* c = u->ctx;
  CU(c, cudaSetDevice(c->device));
  int r = ensure_plan(u, B, B, h, w);
  if (r) return r;
  Plan* P = u->plan.get();
  KL(c, cast_f16_to_f32_launch(c->stream, (const __half*)x, (size_t)B * u->cfg.in_channels * h * w, P->x_in));
  if ((r = set_t(u, t_host))) return r;
  if ((r = run_plan(u))) return r;
  KL(c, nhwc_to_nchw_f16_launch(c->stream, P->eps, B, h * w, u->cfg.out_channels, P->eps_ld, (__half*)eps_out));
  return 0;
}
extern "C" int sdxl_unet_forward_f32(sdxl_unet* u, int B, int h, int w, const float* x, int32_t t_host, float* eps_out) {
  if (!u || !x || !eps_out) return -1;
  sdxl_ctx* c = u->ctx;
  CU(c, cudaSetDevice(c->device));
  int r = ensure_plan(u, B, B, h, w);
  if (r) return r;
  Plan* P = u->plan.get();
  CU(c, cudaMemcpyAsync(P->x_in, x, (size_t)B * u->cfg.in_channels * h * w * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  if ((r = set_t(u, t_host))) return r;
  if ((r = run_plan(u))) return r;
  KL(c, nhwc_to_nchw_f32_launch(c->stream, P->eps, B, h * w, u->cfg.out_channels, P->eps_ld, eps_out));
  return 0;
}
// Per-kernel-kind device time of one plan execution, measured with CUDA events on the ctx stream
// (eager launches, one event pair per op). kinds: see OpKind. Arrays must hold 16 entries.
static int profile_plan_impl(sdxl_ctx* c, Plan* P, double* ms_by_kind, double* flops_by_kind, int* launches_by_kind) {
  const size_t n = P->ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CU(c, cudaEventCreate(&e));
  int r = 0;
  CU(c, cudaEventRecord(ev[0], c->stream));
  for (size_t i = 0; i < n && !r; ++i) {
    r = exec_op(c, P->ops[i]);
    if (!r && cudaEventRecord(ev[i + 1], c->stream) != cudaSuccess) r = -2;
  }
  cudaError_t se = cudaStreamSynchronize(c->stream);
  for (int k = 0; k < 16; ++k) { ms_by_kind[k] = 0; flops_by_kind[k] = 0; launches_by_kind[k] = 0; }
  if (!r && se == cudaSuccess)
    for (size_t i = 0; i < n; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      const int k = (int)P->ops[i].kind;
      ms_by_kind[k] += ms;
      flops_by_kind[k] += P->ops[i].flops;
      launches_by_kind[k] += (P->ops[i].kind == OP_GN) ? 2 : 1;
    }
  for (auto& e : ev) cudaEventDestroy(e);
  if (se != cudaSuccess) return fail(c, (int)se, "profile run failed: %s", cudaGetErrorString(se));
  return r;
}
extern "C" int sdxl_unet_profile_plan(sdxl_unet* u, double* ms_by_kind, double* flops_by_kind, int* launches_by_kind) {
  if (!u || !u->plan) return -1;
  return profile_plan_impl(u->ctx, u->plan.get(), ms_by_kind, flops_by_kind, launches_by_kind);
}
// Per-op dump of one eager plan execution (CUDA-event time per launch) as CSV: analysis aid for profiles/.
static int profile_dump_impl(sdxl_ctx* c, Plan* P, const char* path) {
  const size_t n = P->ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CU(c, cudaEventCreate(&e));
  int r = 0;
  CU(c, cudaEventRecord(ev[0], c->stream));
  for (size_t i = 0; i < n && !r; ++i) {
    r = exec_op(c, P->ops[i]);
    if (!r && cudaEventRecord(ev[i + 1], c->stream) != cudaSuccess) r = -2;
  }
  cudaError_t se = cudaStreamSynchronize(c->stream);
  if (!r && se == cudaSuccess) {
    FILE* f = fopen(path, "w");
    if (!f) r = fail(c, -3, "cannot open %s", path);
    else {
      fprintf(f, "op,kind,us,gflop,tflops,M_tiles,N,BN,Kblocks,T,S,heads,cluster\n");
      for (size_t i = 0; i < n; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        const Op& o = P->ops[i];
        int mt = 0, N = 0, BN = 0, kb = 0, T = 0, S = 0, H = 0;
        if (o.kind == OP_IGEMM) {
          mt = o.ig.tilesW * o.ig.tilesH * o.ig.tilesB; N = o.ig.N; BN = o.ig.BN;
          for (int s2 = 0; s2 < o.ig.nseg; ++s2) kb += o.ig.seg[s2].nkb;
        } else if (o.kind == OP_ATTN) { T = o.at.T; S = o.at.S; H = o.at.n_head; }
        fprintf(f, "%zu,%s,%.2f,%.3f,%.1f,%d,%d,%d,%d,%d,%d,%d,%dx%d\n", i, kOpNames[o.kind], ms * 1e3, o.flops * 1e-9,
                ms > 0 ? o.flops / (ms * 1e-3) * 1e-12 : 0.0, mt, N, BN, kb, T, S, H,
                o.kind == OP_IGEMM ? (o.ig.pair ? 9 : o.ig.CM) : 0, o.kind == OP_IGEMM ? o.ig.CN : 0);
      }
      fclose(f);
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (se != cudaSuccess) return fail(c, (int)se, "profile run failed: %s", cudaGetErrorString(se));
  return r;
}
extern "C" int sdxl_unet_profile_dump(sdxl_unet* u, const char* path) {
  if (!u || !u->plan || !path) return -1;
  return profile_dump_impl(u->ctx, u->plan.get(), path);
}
extern "C" double sdxl_unet_alpha(const sdxl_unet* u, int i) {
  if (!u || i < 0 || i >= (int)u->alphas.size()) return NAN;
  return u->alphas[i];
}
// algorithmic FLOPs of the current plan (debug / bench helper, not in the public header)
extern "C" double sdxl_unet_plan_flops(const sdxl_unet* u) { return (u && u->plan) ? u->plan->flops : 0.0; }
extern "C" int sdxl_unet_plan_num_ops(const sdxl_unet* u) { return (u && u->plan) ? (int)u->plan->ops.size() : 0; }

// ================================================================================================
// sampler (Diffuser)
// ================================================================================================
struct Sampler {
  int Bimg = 0, nfwd = 1, h = 0, w = 0;
  float guidance = 1.f;
  float* noise = nullptr;  // scratch [Bimg,4,h,w]
  float* ref = nullptr;
  uint8_t* mask = nullptr;
  __half* cond_ctx = nullptr;  // staged [nfwd*Bimg, n_ctx, ctx]
  __half* cond_y = nullptr;
  float* host_stage = nullptr;  // pinned
  size_t latent_elems = 0;
  Arena arena;
  ~Sampler() {
    arena.release();
    if (host_stage) cudaFreeHost(host_stage);
  }
};

// Uploads/assembles the batched conditioning: rows [0,Bimg) conditional, rows [Bimg,2*Bimg) the
// unconditional context repeated (reference stablediffusion/mod.rs:506-537).
static int sampler_begin(sdxl_unet* u, const sdxl_conditioning* cond, double guidance) {
  sdxl_ctx* c = u->ctx;
  const sdxl_unet_cfg& g = u->cfg;
  if (!cond) return fail(c, 5200, "null conditioning");
  const int Bimg = cond->n_batch, n_ctx = cond->n_ctx;
  const int h = cond->resolution[0] / 8, w = cond->resolution[1] / 8;
  const int nfwd = g.is_refiner ? 1 : 2;
  const sdxl_half* ctx_c = g.is_refiner ? cond->context_open_clip : cond->context_full;
  const sdxl_half* ctx_u = g.is_refiner ? cond->unconditional_context_open_clip : cond->unconditional_context_full;
  const sdxl_half* y_c = g.is_refiner ? cond->channel_context_refiner : cond->channel_context;
  const sdxl_half* y_u = g.is_refiner ? cond->unconditional_channel_context_refiner : cond->unconditional_channel_context;
  if (!ctx_c || !y_c || (nfwd == 2 && (!ctx_u || !y_u))) return fail(c, 5201, "conditioning tensors for this model are null");
  if (Bimg < 1 || h < 1 || w < 1) return fail(c, 5202, "bad conditioning batch/resolution");
  Sampler* S = u->sampler.get();
  const size_t lat = (size_t)Bimg * g.in_channels * h * w;
  if (!S || S->Bimg != Bimg || S->h != h || S->w != w || S->nfwd != nfwd) {
    CU(c, cudaStreamSynchronize(c->stream));
    u->sampler.reset(new Sampler());
    S = u->sampler.get();
    S->Bimg = Bimg; S->nfwd = nfwd; S->h = h; S->w = w; S->latent_elems = lat;
    const size_t ctx_elems = (size_t)nfwd * Bimg * n_ctx * g.context_dim;
    const size_t y_elems = (size_t)nfwd * Bimg * g.adm_in_channels;
    if (S->arena.init(lat * 4 * 2 + lat + ctx_elems * 2 + y_elems * 2 + (1 << 16))) return fail(c, 5203, "cannot allocate sampler buffers");
    S->noise = S->arena.get<float>(lat);
    S->ref = S->arena.get<float>(lat);
    S->mask = S->arena.get<uint8_t>(lat);
    S->cond_ctx = S->arena.get<__half>(ctx_elems);
    S->cond_y = S->arena.get<__half>(y_elems);
    CU(c, cudaMallocHost((void**)&S->host_stage, lat * sizeof(float)));
  }
  S->guidance = (float)guidance;
  const cudaMemcpyKind kind = cond->on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const size_t ctx_row = (size_t)n_ctx * g.context_dim * 2, y_row = (size_t)g.adm_in_channels * 2;
  CU(c, cudaMemcpyAsync(S->cond_ctx, ctx_c, ctx_row * Bimg, kind, c->stream));
  CU(c, cudaMemcpyAsync(S->cond_y, y_c, y_row * Bimg, kind, c->stream));
  if (nfwd == 2)
    for (int b = 0; b < Bimg; ++b) {  // unsqueeze().repeat(0, n_batch)
      CU(c, cudaMemcpyAsync((uint8_t*)S->cond_ctx + ctx_row * (Bimg + b), ctx_u, ctx_row, kind, c->stream));
      CU(c, cudaMemcpyAsync((uint8_t*)S->cond_y + y_row * (Bimg + b), y_u, y_row, kind, c->stream));
    }
  int r = set_conditioning_dev(u, nfwd * Bimg, n_ctx, S->cond_ctx, S->cond_y);
  if (r) return r;
  return ensure_plan(u, nfwd * Bimg, Bimg, h, w);
}

// one loop-body iteration (reference stablediffusion/mod.rs:406-429)
static int sampler_step(sdxl_unet* u, int t, int t_prev) {
  sdxl_ctx* c = u->ctx;
  Sampler* S = u->sampler.get();
  Plan* P = u->plan.get();
  if (!S || !P) return fail(c, 5210, "sampler not initialised (call sdxl_sampler_begin)");
  if (t < 0 || t >= (int)u->alphas.size() || t_prev >= (int)u->alphas.size()) return fail(c, 5211, "timestep out of range");
  const double a = u->alphas[t];
  const double ap = t_prev >= 0 ? u->alphas[t_prev] : 1.0;
  int r = set_t(u, t);
  if (r) return r;
  if ((r = run_plan(u))) return r;
  KL(c, cfg_ddim_launch(c->stream, P->eps, P->eps_ld, S->Bimg, u->cfg.in_channels, S->h * S->w, S->nfwd == 2, S->guidance,
                        (float)sqrt(a), (float)sqrt(1.0 - a), (float)sqrt(ap), (float)sqrt(1.0 - ap), P->x_in, nullptr));
  return 0;
}

extern "C" int sdxl_sampler_begin(sdxl_unet* u, const sdxl_conditioning* cond, double guidance_scale) {
  if (!u) return -1;
  CU(u->ctx, cudaSetDevice(u->ctx->device));
  return sampler_begin(u, cond, guidance_scale);
}
extern "C" int sdxl_sampler_step(sdxl_unet* u, int t, int t_prev) {
  if (!u) return -1;
  return sampler_step(u, t, t_prev);
}
extern "C" int sdxl_sampler_set_latent(sdxl_unet* u, const float* latent, int on_host) {
  if (!u || !u->sampler || !u->plan) return -1;
  sdxl_ctx* c = u->ctx;
  CU(c, cudaMemcpyAsync(u->plan->x_in, latent, u->sampler->latent_elems * 4, on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
  if (on_host) CU(c, cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int sdxl_sampler_get_latent(sdxl_unet* u, float* latent, int on_host) {
  if (!u || !u->sampler || !u->plan) return -1;
  sdxl_ctx* c = u->ctx;
  CU(c, cudaMemcpyAsync(latent, u->plan->x_in, u->sampler->latent_elems * 4, on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
  if (on_host) CU(c, cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int sdxl_sampler_step_host(sdxl_unet* u, int t, int t_prev, float* latent_host) {
  if (!u || !u->sampler || !u->plan || !latent_host) return -1;
  sdxl_ctx* c = u->ctx;
  Sampler* S = u->sampler.get();
  const size_t bytes = S->latent_elems * 4;
  memcpy(S->host_stage, latent_host, bytes);  // caller memory may be pageable: stage through pinned
  CU(c, cudaMemcpyAsync(u->plan->x_in, S->host_stage, bytes, cudaMemcpyHostToDevice, c->stream));
  int r = sampler_step(u, t, t_prev);
  if (r) return r;
  CU(c, cudaMemcpyAsync(S->host_stage, u->plan->x_in, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  memcpy(latent_host, S->host_stage, bytes);
  return 0;
}
extern "C" int sdxl_randn(sdxl_ctx* c, float* out, size_t n, uint64_t seed, uint64_t subsequence) {
  if (!c || !out) return -1;
  KL(c, randn_launch(c->stream, out, n, seed, subsequence));
  return 0;
}

extern "C" int sdxl_sample_latent(sdxl_unet* u, const sdxl_conditioning* cond, double guidance_scale, int n_steps,
                                  int step_start, const float* init_latent, const float* noise, int n_noise, uint64_t seed,
                                  const float* inpaint_ref, const uint8_t* inpaint_mask, float* latent_out) {
  if (!u || !cond || !latent_out) return -1;
  sdxl_ctx* c = u->ctx;
  CU(c, cudaSetDevice(c->device));
  const int total = u->cfg.n_steps;
  if (n_steps < 1 || n_steps > total) return fail(c, 5220, "n_steps must be in [1,%d]", total);
  if (step_start < 0 || step_start >= total) return fail(c, 5221, "bad step_start");
  if ((inpaint_ref == nullptr) != (inpaint_mask == nullptr)) return fail(c, 5222, "inpaint_ref and inpaint_mask must be given together");
  if (step_start > 0 && !init_latent) return fail(c, 5223, "refine (step_start>0) needs init_latent");
  int r = sampler_begin(u, cond, guidance_scale);
  if (r) return r;
  Sampler* S = u->sampler.get();
  Plan* P = u->plan.get();
  const size_t lat = S->latent_elems, bytes = lat * 4;
  const cudaMemcpyKind in_kind = cond->on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const int step_size = total / n_steps;      // mod.rs:400
  const int t_begin = total - step_start;     // mod.rs:404
  int noise_used = 0;
  uint64_t subseq = 0;
  auto next_noise = [&](float* dst) -> int {  // injected noise first, then the seeded stream
    if (noise && noise_used < n_noise) {
      CU(c, cudaMemcpyAsync(dst, noise + (size_t)noise_used * lat, bytes, in_kind, c->stream));
      noise_used++;
      return 0;
    }
    KL(c, randn_launch(c->stream, dst, lat, seed, subseq++));
    return 0;
  };
  // initial latent
  if (init_latent) CU(c, cudaMemcpyAsync(P->x_in, init_latent, bytes, in_kind, c->stream));
  else if ((r = next_noise(P->x_in))) return r;
  if (step_start > 0) {
    // refine_latent entry (mod.rs:363-367): x = x*sqrt(a_t0) + noise*sqrt(1-a_t0), t0 = n_steps_total - step_start
    const double a0 = u->alphas[t_begin];
    if ((r = next_noise(S->noise))) return r;
    KL(c, axpby_launch(c->stream, P->x_in, S->noise, lat, (float)sqrt(a0), (float)sqrt(1.0 - a0)));
  }
  if (inpaint_ref) {
    CU(c, cudaMemcpyAsync(S->ref, inpaint_ref, bytes, in_kind, c->stream));
    CU(c, cudaMemcpyAsync(S->mask, inpaint_mask, lat, in_kind, c->stream));
  }
  // for t in (0..t_begin).rev().step_by(step_size)   (mod.rs:406, 452)
  for (int t = t_begin - 1; t >= 0; t -= step_size) {
    const int t_prev = (t >= step_size) ? t - step_size : -1;
    if (inpaint_ref) {
      const double a = u->alphas[t];
      if ((r = next_noise(S->noise))) return r;
      KL(c, inpaint_blend_launch(c->stream, P->x_in, S->ref, S->noise, S->mask, lat, S->nfwd, (float)sqrt(a), (float)sqrt(1.0 - a), nullptr));
    }
    if ((r = sampler_step(u, t, t_prev))) return r;
  }
  CU(c, cudaMemcpyAsync(latent_out, P->x_in, bytes, cond->on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
  if (cond->on_host) CU(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" void sdxl_unet_destroy(sdxl_unet* u) {
  if (!u) return;
  cudaStreamSynchronize(u->ctx->stream);
  u->plan.reset();
  u->sampler.reset();
  u->warena.release();
  u->carena.release();
  if (u->t_dev) cudaFree(u->t_dev);
  if (u->t_pinned) cudaFreeHost(u->t_pinned);
  delete u;
}

// ================================================================================================
// operator-level entry points
// ================================================================================================
struct TmpBufs {
  std::vector<void*> p;
  cudaStream_t st;
  explicit TmpBufs(cudaStream_t s) : st(s) {}
  void* get(size_t bytes) {
    void* d = nullptr;
    if (cudaMallocAsync(&d, bytes ? bytes : 16, st) != cudaSuccess) return nullptr;
    p.push_back(d);
    return d;
  }
  ~TmpBufs() { for (void* d : p) cudaFreeAsync(d, st); }
};

extern "C" int sdxl_qkv_attention(sdxl_ctx* c, const sdxl_half* q, const sdxl_half* k, const sdxl_half* v, const sdxl_half* mask,
                                  int B, int T, int S, int C, int n_head, sdxl_half* out) {
  if (!c || !q || !k || !v || !out) return -1;
  if (n_head < 1 || C != n_head * 64) return fail(c, 5301, "sdxl_qkv_attention: head dim must be 64 (C=%d, n_head=%d)", C, n_head);
  if (mask) {
    // additive [T,S] mask (the text encoders' causal mask, clip/mod.rs:88): short sequences, CUDA-core kernel
    KL(c, attention_small_launch(c->stream, (const __half*)q, C, 0, (const __half*)k, (const __half*)v, C, 0, 0, B, T, S, n_head,
                                 (const __half*)mask, 0, (__half*)out, C));
    return 0;
  }
  AttnParams p{};
  p.T = T; p.S = S; p.n_head = n_head; p.B = B;
  p.q_col0 = p.k_col0 = p.v_col0 = 0;
  p.out = (__half*)out; p.ldo = C;
  p.scale_log2e = (float)(1.4426950408889634 / sqrt(64.0));
  int r = make_tmap_rows(&p.tmQ, (const __half*)q, T, B, C, C);
  if (!r) r = make_tmap_rows(&p.tmK, (const __half*)k, S, B, C, C);
  if (!r) r = make_tmap_rows(&p.tmV, (const __half*)v, S, B, C, C);
  if (r) return fail(c, r, "tensor map creation failed");
  KL(c, attention_launch(c->stream, p));
  return 0;
}

extern "C" int sdxl_op_linear(sdxl_ctx* c, const sdxl_half* x, const sdxl_half* w, const sdxl_half* bias, const float* residual,
                              int M, int K, int N, int geglu, int out_f16, void* out) {
  if (!c || !x || !w || !out) return -1;
  if (K % 8) return fail(c, 5310, "sdxl_op_linear: K must be a multiple of 8");
  TmpBufs T(c->stream);
  const int Kpad = (K + 63) / 64 * 64;
  int gbn = 0;
  if (geglu) {
    gbn = geglu_bn_for(N / 2);
    if (!gbn || (N & 1)) return fail(c, 5311, "sdxl_op_linear: GEGLU width not tileable");
  }
  __half* wt = (__half*)T.get((size_t)N * Kpad * 2);
  float* b32 = bias ? (float*)T.get((size_t)N * 4) : nullptr;
  if (!wt || (bias && !b32)) return fail(c, 5312, "temporary allocation failed");
  KL(c, transpose_linear_launch(c->stream, (const __half*)w, K, N, wt, Kpad, 0, gbn));
  if (bias) KL(c, bias_to_f32_launch(c->stream, (const __half*)bias, N, b32, gbn, 0));
  IgemmParams p{};
  p.nseg = 1;
  p.seg[0] = {0, 0, 0, 0, Kpad / 64};
  p.out = out;
  p.out_f32 = geglu ? 0 : !out_f16;
  p.ldo = geglu ? N / 2 : N;
  p.bias = b32; p.bias_bstride = 0;
  p.res = geglu ? nullptr : residual; p.ldr = N;
  IgemmOperands o{(const __half*)x, 1, 1, M, K, K, nullptr, 0, 0, 0, 0, 0, wt, N, Kpad};
  int r = igemm_configure(p, o, M, 1, 1, geglu ? IGEMM_GEGLU : IGEMM_LINEAR, gbn);
  if (r) return fail(c, r, "igemm configuration failed");
  KL(c, igemm_launch(c->stream, p));
  return 0;
}

extern "C" int sdxl_op_conv2d(sdxl_ctx* c, const float* x, const sdxl_half* w, const sdxl_half* bias, int B, int H, int W, int Cin,
                              int Cout, int ksize, int stride, int upsample, float* out) {
  if (!c || !x || !w || !out) return -1;
  if ((ksize != 1 && ksize != 3) || (stride != 1 && stride != 2) || (stride == 2 && (upsample || ksize != 3)))
    return fail(c, 5320, "sdxl_op_conv2d: unsupported ksize/stride/upsample combination");
  if (Cin % 8) return fail(c, 5321, "sdxl_op_conv2d: Cin must be a multiple of 8");
  TmpBufs T(c->stream);
  const int Ipad = (Cin + 63) / 64 * 64, Ktot = ksize * ksize * Ipad;
  __half* wt = (__half*)T.get((size_t)Cout * Ktot * 2);
  float* b32 = bias ? (float*)T.get((size_t)Cout * 4) : nullptr;
  const int Hi = upsample ? 2 * H : H, Wi = upsample ? 2 * W : W;  // conv input extent
  __half* a16 = (__half*)T.get((size_t)B * Hi * Wi * Cin * 2);
  if (!wt || !a16 || (bias && !b32)) return fail(c, 5322, "temporary allocation failed");
  KL(c, repack_conv_launch(c->stream, (const __half*)w, Cout, Cin, ksize, ksize, wt, Ktot, 0, Ipad));
  if (bias) KL(c, bias_to_f32_launch(c->stream, (const __half*)bias, Cout, b32, 0, 0));
  IgemmParams p{};
  int Ho = Hi, Wo = Wi;
  ActView a{a16, B, Hi, Wi, Cin};
  p.nseg = 0;
  if (stride == 2) {
    if ((H & 1) || (W & 1)) return fail(c, 5323, "sdxl_op_conv2d: stride 2 needs even H, W");
    KL(c, phase_split_launch(c->stream, x, B, H, W, Cin, a16));
    Ho = H / 2; Wo = W / 2;
    a = ActView{a16, 4 * B, Ho, Wo, Cin};
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int ph = (kh == 1) ? 0 : 1, pw = (kw == 1) ? 0 : 1;
        p.seg[p.nseg++] = {0, (int16_t)((kw == 0) ? -1 : 0), (int16_t)((kh == 0) ? -1 : 0), (int16_t)((ph * 2 + pw) * B), Ipad / 64};
      }
  } else {
    if (upsample) KL(c, upsample2x_launch(c->stream, x, B, H, W, Cin, a16));
    else KL(c, cast_f32_to_f16_launch(c->stream, x, (size_t)B * H * W * Cin, a16));
    const int pad = ksize / 2;
    for (int kh = 0; kh < ksize; ++kh)
      for (int kw = 0; kw < ksize; ++kw) p.seg[p.nseg++] = {0, (int16_t)(kw - pad), (int16_t)(kh - pad), 0, Ipad / 64};
  }
  p.out = out; p.out_f32 = 1; p.ldo = Cout;
  p.bias = b32; p.bias_bstride = 0; p.res = nullptr; p.ldr = 0;
  IgemmOperands o{a.p, a.Bn, a.H, a.W, a.C, a.C, nullptr, 0, 0, 0, 0, 0, wt, Cout, Ktot};
  int r = igemm_configure(p, o, Wo, Ho, B, IGEMM_LINEAR, 0);
  if (r) return fail(c, r, "igemm configuration failed");
  KL(c, igemm_launch(c->stream, p));
  return 0;
}

extern "C" int sdxl_op_group_norm(sdxl_ctx* c, const float* x1, int C1, const float* x2, int C2, int B, int HW, int n_group,
                                  const float* gamma, const float* beta, float eps, int silu, sdxl_half* out) {
  if (!c || !x1 || !gamma || !beta || !out) return -1;
  TmpBufs T(c->stream);
  float* part = (float*)T.get(gn_scratch_floats(B, n_group) * 4);
  if (!part) return fail(c, 5330, "temporary allocation failed");
  GnParams p{x1, C1, x2, x2 ? C2 : 0, B, HW, n_group, gamma, beta, eps, silu, (__half*)out, nullptr, part, 0};
  KL(c, gn_launch(c->stream, p));
  c->launches++;
  return 0;
}
extern "C" int sdxl_op_layer_norm(sdxl_ctx* c, const float* x, const float* gamma, const float* beta, float eps, int rows, int C,
                                  sdxl_half* out) {
  if (!c || !x || !gamma || !beta || !out) return -1;
  KL(c, layernorm_launch(c->stream, x, gamma, beta, eps, rows, C, (__half*)out));
  return 0;
}
extern "C" int sdxl_op_timestep_embedding(sdxl_ctx* c, const int32_t* t_host, int n, int dim, int max_period, float* out) {
  if (!c || !t_host || !out || n < 1 || (dim & 1)) return -1;
  TmpBufs T(c->stream);
  int* td = (int*)T.get((size_t)n * 4);
  if (!td) return fail(c, 5340, "temporary allocation failed");
  CU(c, cudaMemcpyAsync(td, t_host, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  KL(c, timestep_embedding_launch(c->stream, td, n, dim, (float)max_period, out));
  CU(c, cudaStreamSynchronize(c->stream));  // t_host is pageable caller memory
  return 0;
}

// Diagnostics: in-kernel timeline (%globaltimer, ns) of CTA 0 of one igemm launch on a synthetic [M,K]x[K,N] problem.
// stamps_host[0..6] = prologue done, dependencies resolved, first operands landed, first accumulator complete,
// first epilogue done, producer done, all roles done; stamps_host[7] = CUDA-event duration of the launch in ns.
extern "C" int sdxl_dbg_igemm_timeline(sdxl_ctx* c, int M, int K, int N, int geglu, int with_residual, uint64_t* stamps_host) {
  if (!c || !stamps_host) return -1;
  TmpBufs T(c->stream);
  const int Kpad = (K + 63) / 64 * 64;
  int gbn = geglu ? geglu_bn_for(N / 2) : 0;
  __half* x = (__half*)T.get((size_t)M * K * 2);
  __half* w = (__half*)T.get((size_t)N * Kpad * 2);
  float* bias = (float*)T.get((size_t)N * 4);
  float* res = (float*)T.get((size_t)M * N * 4);
  void* out = T.get((size_t)M * N * 4);
  unsigned long long* dbg = (unsigned long long*)T.get(16 * 8);
  if (!x || !w || !bias || !res || !out || !dbg) return fail(c, 5400, "temporary allocation failed");
  CU(c, cudaMemsetAsync(x, 0, (size_t)M * K * 2, c->stream));
  CU(c, cudaMemsetAsync(w, 0, (size_t)N * Kpad * 2, c->stream));
  CU(c, cudaMemsetAsync(bias, 0, (size_t)N * 4, c->stream));
  CU(c, cudaMemsetAsync(res, 0, (size_t)M * N * 4, c->stream));
  CU(c, cudaMemsetAsync(dbg, 0, 128, c->stream));
  IgemmParams p{};
  p.nseg = 1;
  p.seg[0] = {0, 0, 0, 0, Kpad / 64};
  p.out = out; p.out_f32 = geglu ? 0 : 1; p.ldo = geglu ? N / 2 : N;
  p.bias = bias; p.bias_bstride = 0;
  p.res = (geglu || !with_residual) ? nullptr : res; p.ldr = N;
  IgemmOperands o{x, 1, 1, M, K, K, nullptr, 0, 0, 0, 0, 0, w, N, Kpad};
  int r = igemm_configure(p, o, M, 1, 1, geglu ? IGEMM_GEGLU : IGEMM_LINEAR, gbn);
  if (r) return fail(c, r, "igemm configuration failed");
  cudaEvent_t e0, e1;
  CU(c, cudaEventCreate(&e0));
  CU(c, cudaEventCreate(&e1));
  if (getenv("SDXL_B200_DBG_MODE")) p.dbg_mode = atoi(getenv("SDXL_B200_DBG_MODE"));
  if (getenv("SDXL_B200_DBG_NST")) p.nstages = atoi(getenv("SDXL_B200_DBG_NST"));
  for (int i = 0; i < 3; ++i) KL(c, igemm_launch(c->stream, p));
  p.dbg = dbg;
  CU(c, cudaEventRecord(e0, c->stream));
  KL(c, igemm_launch(c->stream, p));
  CU(c, cudaEventRecord(e1, c->stream));
  unsigned long long hostbuf[16];
  CU(c, cudaMemcpyAsync(hostbuf, dbg, 128, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  for (int i = 0; i < 7; ++i) stamps_host[i] = hostbuf[i];
  fprintf(stderr, "  epi block0: tmem_ld %llu ns, st.shared+sync %llu ns, phase2 %llu ns (since acc ready: %llu)\n",
          hostbuf[10] - hostbuf[9], hostbuf[11] - hostbuf[10], hostbuf[12] - hostbuf[11], hostbuf[9] - hostbuf[3]);
  stamps_host[7] = (uint64_t)(ms * 1e6);
  stamps_host[8] = (uint64_t)p.BN | ((uint64_t)p.pair << 16) | ((uint64_t)p.CM << 20) | ((uint64_t)p.CN << 24) | ((uint64_t)p.nstages << 28);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

// ================================================================================================
// Latent decoder: LatentDecoder::{decode_latent, latent_to_image}
//   Autoencoder::decode_latent      src/model/autoencoder/mod.rs:66-69
//   Decoder::forward                src/model/autoencoder/mod.rs:202-216
//   Mid / ResnetBlock / ConvSelfAttentionBlock / DecoderBlock
//                                   src/model/autoencoder/mod.rs:436-452, 507-524, 548-586, 298-324
//   LatentDecoder                   src/model/stablediffusion/mod.rs:199-237, 263-266
// Same machinery as the UNet: weights re-laid-out once on the device, a flat launch plan replayed as a CUDA graph.
// All convolutions and the attention contractions run on the tcgen05 implicit-GEMM kernel with f16 operands and f32
// accumulation; the residual stream, GroupNorm statistics, the score matrix and the softmax are f32 (the reference
// runs this module in f32 end to end; tests/test_vae_gpu.py states the resulting tolerance).
// The attention block is single-head with d = C (512): scores are materialised (f32 [T,T] per image, 1.07 GB at
// 1024^2), soft-maxed by rows into f16 probabilities and multiplied with V by a second GEMM.
// ================================================================================================
struct VRes {
  Norm n1, n2;
  Conv c1, c2;  // c2 carries the fused nin_shortcut 1x1 segment when Cin != Cout
  int Cin = 0, Cout = 0;
  bool has_skip = false;
};
struct VBlock {
  VRes r[3];
  bool up = false;
  Conv upc;
  int Cout = 0;
};
struct sdxl_vae {
  sdxl_ctx* ctx = nullptr;
  sdxl_vae_cfg cfg{};
  Arena warena;
  float* pq_w = nullptr;   // [Cl, Cl] f32
  float* pq_b = nullptr;
  float* cin_w = nullptr;  // [C0][3][3][Cl] f32
  float* cin_b = nullptr;
  int C0 = 0;
  VRes mid1, mid2;
  Norm attn_norm;
  Lin aq, ak, av, aproj;
  std::vector<VBlock> blocks;
  Norm norm_out;
  Conv conv_out;  // O padded to 4
  // encoder half (optional)
  bool has_enc = false;
  float* ecin_w = nullptr;  // [EC0][3][3][3] f32
  float* ecin_b = nullptr;
  int EC0 = 0;
  struct EBlock { VRes r[2]; bool down = false; Conv downc; int Cout = 0; };
  std::vector<EBlock> eblocks;
  VRes emid1, emid2;
  Norm eattn_norm;
  Lin eq, ek, ev, eproj;
  Norm enorm_out;
  Conv econv_out;           // Cm -> enc_z_channels
  float* qc_w = nullptr;    // quant_conv [Cz][Cz] f32
  float* qc_b = nullptr;
  std::unique_ptr<Plan> enc_plan;
  float* enc_z = nullptr;        // [B, hw, Cz] f32 NHWC (conv_out output)
  float* enc_lat = nullptr;      // [B, Cl, hw] f32 NCHW staging
  uint8_t* enc_u8 = nullptr;     // [B, HW, 3] staging for host u8 input
  std::unique_ptr<Plan> plan;
  float* img_nhwc = nullptr;     // [B, 64hw, 4] f32 (decoder output, first 3 channels valid)
  float* out_f32 = nullptr;      // [B, 3, 8h, 8w] staging for host reads
  uint8_t* out_u8 = nullptr;     // [B, 8h, 8w, 3]
};

static Lin lin_from_conv1x1(const Conv& cv) {
  Lin L;
  L.w = cv.w; L.b = cv.b; L.K = cv.I; L.Kpad = cv.Ipad; L.N = cv.O;
  return L;
}
static VRes load_vres(Loader& L, const std::string& path, int Cin, int Cout) {
  VRes r;
  r.Cin = Cin; r.Cout = Cout; r.has_skip = (Cin != Cout);
  r.n1 = L.norm(path + "/norm1", Cin);
  r.c1 = L.conv(path + "/conv1", Cin, Cout, 3);
  r.n2 = L.norm(path + "/norm2", Cout);
  if (r.has_skip) r.c2 = L.conv(path + "/conv2", Cout, Cout, 3, path + "/nin_shortcut", Cin);
  else r.c2 = L.conv(path + "/conv2", Cout, Cout, 3);
  return r;
}
static int build_vae(sdxl_vae* v, const PackView& pv, Arena& A) {
  sdxl_ctx* c = v->ctx;
  const sdxl_vae_cfg& g = v->cfg;
  Loader L{nullptr, c, &pv, &A, c->stream};
  const int Cl = g.latent_channels;
  v->blocks.clear();
  v->C0 = g.block_in[0];
  // post_quant_conv: OIHW [Cl,Cl,1,1] f16 -> f32 [Cl][Cl]
  {
    const PackEntry* e = L.need("post_quant_conv/weight", 4);
    if (!e) return L.err;
    if ((int)e->shape[0] != Cl || (int)e->shape[1] != Cl || e->shape[2] != 1 || e->shape[3] != 1) return fail(c, 4301, "post_quant_conv/weight bad shape");
    v->pq_w = A.get<float>((size_t)Cl * Cl);
    if (!v->pq_w) return fail(c, 4005, "weight arena exhausted");
    if (!A.measure) { int r = cast_f16_to_f32_launch(c->stream, L.ptr(e), (size_t)Cl * Cl, v->pq_w); if (r) return fail(c, r, "post_quant cast failed"); }
    v->pq_b = L.vec_f32("post_quant_conv/bias", Cl);
    if (L.err) return L.err;
  }
  // decoder/conv_in: OIHW f16 -> [O][kh][kw][I] f32 (CUDA-core kernel, exact f32 like the reference)
  {
    const PackEntry* e = L.need("decoder/conv_in/weight", 4);
    if (!e) return L.err;
    if ((int)e->shape[0] != v->C0 || (int)e->shape[1] != Cl || e->shape[2] != 3 || e->shape[3] != 3) return fail(c, 4302, "decoder/conv_in/weight bad shape");
    const size_t n = (size_t)v->C0 * 9 * Cl;
    __half* tmp = A.get<__half>(n);
    v->cin_w = A.get<float>(n);
    if (!tmp || !v->cin_w) return fail(c, 4005, "weight arena exhausted");
    if (!A.measure) {
      int r = repack_conv_launch(c->stream, L.ptr(e), v->C0, Cl, 3, 3, tmp, 9 * Cl, 0, Cl);
      if (!r) r = cast_f16_to_f32_launch(c->stream, tmp, n, v->cin_w);
      if (r) return fail(c, r, "decoder conv_in repack failed");
    }
    v->cin_b = L.vec_f32("decoder/conv_in/bias", v->C0);
    if (L.err) return L.err;
  }
  const int Cm = v->C0;
  v->mid1 = load_vres(L, "decoder/mid/block_1", Cm, Cm);
  v->attn_norm = L.norm("decoder/mid/attn/norm", Cm);
  v->aq = lin_from_conv1x1(L.conv("decoder/mid/attn/q", Cm, Cm, 1));
  v->ak = lin_from_conv1x1(L.conv("decoder/mid/attn/k", Cm, Cm, 1));
  v->av = lin_from_conv1x1(L.conv("decoder/mid/attn/v", Cm, Cm, 1));
  v->aproj = lin_from_conv1x1(L.conv("decoder/mid/attn/proj_out", Cm, Cm, 1));
  v->mid2 = load_vres(L, "decoder/mid/block_2", Cm, Cm);
  if (L.err) return L.err;
  for (int i = 0; i < g.n_blocks && !L.err; ++i) {
    VBlock b;
    const std::string bp = "decoder/blocks/" + std::to_string(i);
    const int ci = g.block_in[i], co = g.block_out[i];
    b.Cout = co;
    b.r[0] = load_vres(L, bp + "/res1", ci, co);
    b.r[1] = load_vres(L, bp + "/res2", co, co);
    b.r[2] = load_vres(L, bp + "/res3", co, co);
    b.up = (i != g.n_blocks - 1);
    if (b.up) b.upc = L.conv(bp + "/upsampler", co, co, 3);
    v->blocks.push_back(b);
  }
  if (L.err) return L.err;
  const int Cf = g.block_out[g.n_blocks - 1];
  v->norm_out = L.norm("decoder/norm_out", Cf);
  v->conv_out = L.conv("decoder/conv_out", Cf, 3, 3, "", 0, 4);
  if (L.err) return L.err;
  // ---- encoder half (autoencoder/load.rs:82-116)
  v->has_enc = g.n_enc_blocks > 0;
  v->eblocks.clear();
  if (v->has_enc) {
    v->EC0 = g.enc_in[0];
    const PackEntry* e = L.need("encoder/conv_in/weight", 4);
    if (!e) return L.err;
    if ((int)e->shape[0] != v->EC0 || e->shape[1] != 3 || e->shape[2] != 3 || e->shape[3] != 3) return fail(c, 4320, "encoder/conv_in/weight bad shape");
    const size_t n = (size_t)v->EC0 * 27;
    __half* tmp = A.get<__half>(n);
    v->ecin_w = A.get<float>(n);
    if (!tmp || !v->ecin_w) return fail(c, 4005, "weight arena exhausted");
    if (!A.measure) {
      int r = repack_conv_launch(c->stream, L.ptr(e), v->EC0, 3, 3, 3, tmp, 27, 0, 3);
      if (!r) r = cast_f16_to_f32_launch(c->stream, tmp, n, v->ecin_w);
      if (r) return fail(c, r, "encoder conv_in repack failed");
    }
    v->ecin_b = L.vec_f32("encoder/conv_in/bias", v->EC0);
    for (int i = 0; i < g.n_enc_blocks && !L.err; ++i) {
      sdxl_vae::EBlock b;
      const std::string bp = "encoder/blocks/" + std::to_string(i);
      const int ci = g.enc_in[i], co = g.enc_out[i];
      b.Cout = co;
      b.r[0] = load_vres(L, bp + "/res1", ci, co);
      b.r[1] = load_vres(L, bp + "/res2", co, co);
      b.down = (i != g.n_enc_blocks - 1);
      if (b.down) b.downc = L.conv(bp + "/downsampler/conv", co, co, 3);
      v->eblocks.push_back(b);
    }
    if (L.err) return L.err;
    const int Ce = g.enc_out[g.n_enc_blocks - 1], Cz = g.enc_z_channels;
    v->emid1 = load_vres(L, "encoder/mid/block_1", Ce, Ce);
    v->eattn_norm = L.norm("encoder/mid/attn/norm", Ce);
    v->eq = lin_from_conv1x1(L.conv("encoder/mid/attn/q", Ce, Ce, 1));
    v->ek = lin_from_conv1x1(L.conv("encoder/mid/attn/k", Ce, Ce, 1));
    v->ev = lin_from_conv1x1(L.conv("encoder/mid/attn/v", Ce, Ce, 1));
    v->eproj = lin_from_conv1x1(L.conv("encoder/mid/attn/proj_out", Ce, Ce, 1));
    v->emid2 = load_vres(L, "encoder/mid/block_2", Ce, Ce);
    v->enorm_out = L.norm("encoder/norm_out", Ce);
    v->econv_out = L.conv("encoder/conv_out", Ce, Cz, 3);
    if (L.err) return L.err;
    const PackEntry* q = L.need("quant_conv/weight", 4);
    if (!q) return L.err;
    if ((int)q->shape[0] != Cz || (int)q->shape[1] != Cz || q->shape[2] != 1 || q->shape[3] != 1) return fail(c, 4321, "quant_conv/weight bad shape");
    v->qc_w = A.get<float>((size_t)Cz * Cz);
    if (!v->qc_w) return fail(c, 4005, "weight arena exhausted");
    if (!A.measure) { int r = cast_f16_to_f32_launch(c->stream, L.ptr(q), (size_t)Cz * Cz, v->qc_w); if (r) return fail(c, r, "quant_conv cast failed"); }
    v->qc_b = L.vec_f32("quant_conv/bias", Cz);
  }
  return L.err;
}

extern "C" void sdxl_vae_destroy(sdxl_vae* v) {
  if (!v) return;
  cudaStreamSynchronize(v->ctx->stream);
  v->plan.reset();
  v->enc_plan.reset();
  v->warena.release();
  delete v;
}

extern "C" int sdxl_vae_load(sdxl_ctx* c, const sdxl_vae_cfg* cfg, const void* pack, size_t bytes, int pack_on_device,
                             sdxl_vae** out) {
  if (!c || !cfg || !pack || !out) return fail(c, -1, "sdxl_vae_load: null argument");
  *out = nullptr;
  if (cfg->n_blocks < 1 || cfg->n_blocks > SDXL_MAX_LEVELS) return fail(c, 4310, "bad n_blocks");
  if (cfg->latent_channels < 1 || cfg->latent_channels > 8) return fail(c, 4311, "latent_channels must be 1..8");
  if (cfg->n_group != 32) return fail(c, 4312, "n_group must be 32 (got %d)", cfg->n_group);
  if (!(cfg->scale_factor > 0)) return fail(c, 4313, "scale_factor must be positive");
  for (int i = 0; i < cfg->n_blocks; ++i) {
    if (cfg->block_in[i] % 64 || cfg->block_out[i] % 64) return fail(c, 4314, "decoder widths must be multiples of 64");
    if (i && cfg->block_in[i] != cfg->block_out[i - 1]) return fail(c, 4315, "block_in[%d] != block_out[%d]", i, i - 1);
  }
  if (cfg->n_enc_blocks < 0 || cfg->n_enc_blocks > SDXL_MAX_LEVELS) return fail(c, 4316, "bad n_enc_blocks");
  for (int i = 0; i < cfg->n_enc_blocks; ++i) {
    if (cfg->enc_in[i] % 64 || cfg->enc_out[i] % 64) return fail(c, 4317, "encoder widths must be multiples of 64");
    if (i && cfg->enc_in[i] != cfg->enc_out[i - 1]) return fail(c, 4318, "enc_in[%d] != enc_out[%d]", i, i - 1);
  }
  if (cfg->n_enc_blocks && (cfg->enc_z_channels < cfg->latent_channels || cfg->enc_z_channels > 16 || cfg->enc_z_channels % 4))
    return fail(c, 4319, "enc_z_channels must be a multiple of 4 in [latent_channels, 16]");
  CU(c, cudaSetDevice(c->device));
  std::unique_ptr<sdxl_vae> v(new sdxl_vae());
  v->ctx = c;
  v->cfg = *cfg;
  PackView pv;
  std::vector<uint8_t> table;
  int r = parse_pack(c, pack, bytes, pack_on_device, pv, table);
  if (r) return r;
  void* dev_pack = nullptr;
  if (pack_on_device) {
    pv.dev = (const uint8_t*)pack;
  } else {
    CU(c, cudaMalloc(&dev_pack, bytes));
    cudaError_t e = cudaMemcpyAsync(dev_pack, pack, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { cudaFree(dev_pack); return fail(c, (int)e, "pack upload failed"); }
    pv.dev = (const uint8_t*)dev_pack;
  }
  Arena meas;
  meas.measure = true;
  r = build_vae(v.get(), pv, meas);
  if (!r && v->warena.init(meas.off + (1 << 20))) r = fail(c, 4203, "cannot allocate %zu bytes for weights", meas.off);
  if (!r) r = build_vae(v.get(), pv, v->warena);
  cudaError_t se = cudaStreamSynchronize(c->stream);
  if (dev_pack) cudaFree(dev_pack);
  if (!r && se != cudaSuccess) r = fail(c, (int)se, "weight re-layout failed: %s", cudaGetErrorString(se));
  if (r) { v->warena.release(); return r; }
  *out = v.release();
  return 0;
}

// Shared pieces of the encoder / decoder plans: ping-pong f32 stream buffers + scratch, ResnetBlock and the mid attention.
struct VaeStage {
  PlanBuilder& B;
  Plan* P;
  int Bn;
  int H = 0, W = 0;
  float* xb[2] = {nullptr, nullptr};
  int cur = 0;
  __half* s_gn1 = nullptr; __half* s_raw = nullptr; float* s_h = nullptr; __half* s_gn2 = nullptr;
  // attention scratch
  __half* q16 = nullptr; __half* k16 = nullptr; __half* v16 = nullptr; __half* vT = nullptr; __half* ao = nullptr;
  float* S = nullptr; __half* Pm = nullptr;

  void alloc(size_t max_x, size_t max_in, size_t max_out, int T, int Cm) {
    xb[0] = B.buf<float>(Bn * max_x);
    xb[1] = B.buf<float>(Bn * max_x);
    s_gn1 = B.buf<__half>(Bn * max_in);
    s_raw = B.buf<__half>(Bn * max_in);
    s_h = B.buf<float>(Bn * max_out);
    s_gn2 = B.buf<__half>(Bn * max_out);
    q16 = B.buf<__half>((size_t)Bn * T * Cm);
    k16 = B.buf<__half>((size_t)Bn * T * Cm);
    v16 = B.buf<__half>((size_t)Bn * T * Cm);
    vT = B.buf<__half>((size_t)T * Cm);
    ao = B.buf<__half>((size_t)Bn * T * Cm);
    S = B.buf<float>((size_t)T * T);
    Pm = B.buf<__half>((size_t)T * T);
  }
  float* x() const { return xb[cur]; }
  float* other() const { return xb[cur ^ 1]; }
  void flip() { cur ^= 1; }

  // ResnetBlock::forward (autoencoder/mod.rs:507-524)
  void vres(const VRes& r) {
    const int HW = H * W;
    B.gn(x(), r.Cin, nullptr, 0, HW, r.n1, 1, s_gn1, r.has_skip ? s_raw : nullptr);
    ActView a1{s_gn1, Bn, H, W, r.Cin};
    B.conv3(a1, nullptr, r.c1, s_h, r.c1.b, 0, nullptr);
    B.gn(s_h, r.Cout, nullptr, 0, HW, r.n2, 1, s_gn2, nullptr);
    ActView a2{s_gn2, Bn, H, W, r.Cout};
    if (r.has_skip) {
      ActView sk{s_raw, Bn, H, W, r.Cin};
      B.conv3(a2, &sk, r.c2, other(), r.c2.b, 0, nullptr);  // nin_shortcut(x) + h as one GEMM (autoencoder/mod.rs:519-523)
    } else {
      B.conv3(a2, nullptr, r.c2, other(), r.c2.b, 0, x());
    }
    flip();
  }
  // ConvSelfAttentionBlock::forward (autoencoder/mod.rs:548-586): single head, d = C, scores materialised per image
  void attn(const Norm& norm, const Lin& aq, const Lin& ak, const Lin& av, const Lin& aproj) {
    const int T = H * W, Cm = aq.K, M = Bn * T;
    B.gn(x(), Cm, nullptr, 0, T, norm, 0, s_gn1, nullptr);
    B.linear(s_gn1, M, aq, IGEMM_LINEAR, q16, 0, Cm, nullptr, 0);
    B.linear(s_gn1, M, ak, IGEMM_LINEAR, k16, 0, Cm, nullptr, 0);
    B.linear(s_gn1, M, av, IGEMM_LINEAR, v16, 0, Cm, nullptr, 0);
    const int Kp = Loader::pad64(Cm);
    for (int b = 0; b < Bn && !B.err; ++b) {
      const size_t o = (size_t)b * T * Cm;
      {  // S = q k^T  (f32)
        ActView a{q16 + o, 1, 1, T, Cm};
        std::vector<IgemmSeg> segs{{0, 0, 0, 0, Kp / 64}};
        B.igemm(a, nullptr, segs, k16 + o, T, Kp, 1, T, 1, IGEMM_LINEAR, 0, S, 1, T, nullptr, 0, nullptr, 0);
        B.add_flops(2.0 * T * (double)T * Cm);
      }
      {
        Op op{};
        op.kind = OP_SOFTMAX;
        op.sm = {S, (size_t)T, T, T, (float)(1.0 / sqrt((double)Cm)), Pm, (size_t)T};
        P->ops.push_back(op);
      }
      {
        Op op{};
        op.kind = OP_TRANSPOSE;
        op.tr = {v16 + o, (size_t)Cm, T, Cm, vT, (size_t)T};
        P->ops.push_back(op);
      }
      {  // O = P v
        ActView a{Pm, 1, 1, T, T};
        std::vector<IgemmSeg> segs{{0, 0, 0, 0, T / 64}};
        B.igemm(a, nullptr, segs, vT, Cm, T, 1, T, 1, IGEMM_LINEAR, 0, ao + o, 0, Cm, nullptr, 0, nullptr, 0);
        B.add_flops(2.0 * T * (double)T * Cm);
      }
    }
    B.linear(ao, M, aproj, IGEMM_LINEAR, other(), 1, Cm, x(), Cm);  // x + proj_out(attn)
    flip();
  }
};

// Builds the op list of Decoder::forward at batch B, latent h x w.
static int build_vae_plan(sdxl_vae* v, Plan* P, Arena* A) {
  sdxl_ctx* c = v->ctx;
  const sdxl_vae_cfg& g = v->cfg;
  PlanBuilder B{nullptr, c, P, A, P->Bf};
  P->ops.clear();
  P->flops = 0;
  const int Bn = P->Bf, Cl = g.latent_channels;
  VaeStage st{B, P, Bn};
  st.H = P->h; st.W = P->w;
  if ((st.H * st.W) % 64) return fail(c, 5101, "latent %dx%d: h*w must be a multiple of 64", st.H, st.W);

  // buffer maxima over the stages
  size_t max_x = (size_t)st.H * st.W * v->C0, max_in = max_x, max_out = max_x, max_up = 0;
  {
    int hh = st.H, ww = st.W;
    for (const VBlock& b : v->blocks) {
      for (int k = 0; k < 3; ++k) {
        max_in = std::max(max_in, (size_t)hh * ww * b.r[k].Cin);
        max_out = std::max(max_out, (size_t)hh * ww * b.r[k].Cout);
      }
      if (b.up) { hh *= 2; ww *= 2; max_up = std::max(max_up, (size_t)hh * ww * b.Cout); }
      max_x = std::max(max_x, (size_t)hh * ww * b.Cout);
    }
    max_in = std::max(max_in, max_x);  // norm_out operand
  }
  P->x_in = B.buf<float>((size_t)Bn * Cl * st.H * st.W);
  float* pq_out = B.buf<float>((size_t)Bn * Cl * st.H * st.W);
  B.gn_partial = B.buf<float>(gn_scratch_floats(Bn, 32));
  st.alloc(max_x, max_in, max_out, st.H * st.W, v->C0);
  __half* s_up = max_up ? B.buf<__half>(Bn * max_up) : nullptr;
  if (B.err) return B.err;

  // post_quant_conv(latent / scale_factor), conv_in
  {
    Op op{};
    op.kind = OP_PQ;
    op.pq = {P->x_in, Bn, Cl, st.H * st.W, v->pq_w, v->pq_b, (float)(1.0 / g.scale_factor), pq_out};
    P->ops.push_back(op);
    P->flops += 2.0 * Bn * st.H * st.W * (double)Cl * Cl;
  }
  {
    Op op{};
    op.kind = OP_CONV_IN;
    op.ci = {pq_out, Bn, Bn, Cl, st.H, st.W, v->cin_w, v->cin_b, v->C0, st.x()};
    P->ops.push_back(op);
    P->flops += 2.0 * Bn * st.H * st.W * 9.0 * Cl * v->C0;
  }
  // mid: ResnetBlock, ConvSelfAttentionBlock, ResnetBlock
  st.vres(v->mid1);
  st.attn(v->attn_norm, v->aq, v->ak, v->av, v->aproj);
  st.vres(v->mid2);
  // up blocks
  for (const VBlock& b : v->blocks) {
    if (B.err) break;
    for (int k = 0; k < 3; ++k) st.vres(b.r[k]);
    if (b.up) {
      // nearest-2x then 3x3 conv (autoencoder/mod.rs:311-319)
      Op op{};
      op.kind = OP_UPS;
      op.rs = {st.x(), Bn, st.H, st.W, b.Cout, s_up};
      P->ops.push_back(op);
      st.H *= 2; st.W *= 2;
      ActView a{s_up, Bn, st.H, st.W, b.Cout};
      B.conv3(a, nullptr, b.upc, st.other(), b.upc.b, 0, nullptr);
      st.flip();
    }
  }
  if (B.err) return B.err;
  // head: GN -> SiLU -> conv 3x3 to RGB (autoencoder/mod.rs:213-214); N padded to 4
  const int Cf = g.block_out[g.n_blocks - 1];
  const int H = st.H, W = st.W;
  B.gn(st.x(), Cf, nullptr, 0, H * W, v->norm_out, 1, st.s_gn1, nullptr);
  v->img_nhwc = B.buf<float>((size_t)Bn * H * W * 4);
  {
    ActView a{st.s_gn1, Bn, H, W, Cf};
    std::vector<IgemmSeg> segs;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) segs.push_back({0, (int16_t)(kw - 1), (int16_t)(kh - 1), 0, v->conv_out.Ipad / 64});
    B.igemm(a, nullptr, segs, v->conv_out.w, 4, v->conv_out.Ktot, H, W, Bn, IGEMM_LINEAR, 0, v->img_nhwc, 1, 4, v->conv_out.b, 0,
            nullptr, 0);
    B.add_flops(2.0 * Bn * H * W * 9.0 * Cf * 3);
  }
  v->out_f32 = B.buf<float>((size_t)Bn * 3 * H * W);
  v->out_u8 = B.buf<uint8_t>((size_t)Bn * 3 * H * W);
  return B.err;
}

// Builds the op list of Encoder::forward (autoencoder/mod.rs:128-144) + quant_conv at batch B, image H x W.
static int build_vae_enc_plan(sdxl_vae* v, Plan* P, Arena* A) {
  sdxl_ctx* c = v->ctx;
  const sdxl_vae_cfg& g = v->cfg;
  PlanBuilder B{nullptr, c, P, A, P->Bf};
  P->ops.clear();
  P->flops = 0;
  const int Bn = P->Bf, nb = g.n_enc_blocks;
  VaeStage st{B, P, Bn};
  st.H = P->h; st.W = P->w;
  const int down = 1 << (nb - 1);
  if (st.H % down || st.W % down) return fail(c, 5102, "image %dx%d not divisible by %d", st.H, st.W, down);
  const int hl = st.H / down, wl = st.W / down;
  if ((hl * wl) % 64) return fail(c, 5103, "image %dx%d: (H/%d)*(W/%d) must be a multiple of 64", st.H, st.W, down, down);
  size_t max_x = (size_t)st.H * st.W * v->EC0, max_in = 0, max_out = 0, max_ph = 0;
  {
    int hh = st.H, ww = st.W;
    for (const auto& b : v->eblocks) {
      for (int k = 0; k < 2; ++k) {
        max_in = std::max(max_in, (size_t)hh * ww * b.r[k].Cin);
        max_out = std::max(max_out, (size_t)hh * ww * b.r[k].Cout);
        max_x = std::max(max_x, (size_t)hh * ww * b.r[k].Cout);
      }
      if (b.down) { max_ph = std::max(max_ph, (size_t)hh * ww * b.Cout); hh /= 2; ww /= 2; }
    }
    max_in = std::max(max_in, max_x);
  }
  const int Ce = g.enc_out[nb - 1], Cz = g.enc_z_channels, Cl = g.latent_channels;
  P->x_in = B.buf<float>((size_t)Bn * 3 * st.H * st.W);
  v->enc_u8 = B.buf<uint8_t>((size_t)Bn * 3 * st.H * st.W);
  B.gn_partial = B.buf<float>(gn_scratch_floats(Bn, 32));
  st.alloc(max_x, max_in, max_out, hl * wl, Ce);
  __half* s_ph = max_ph ? B.buf<__half>(Bn * max_ph) : nullptr;
  v->enc_z = B.buf<float>((size_t)Bn * hl * wl * Cz);
  v->enc_lat = B.buf<float>((size_t)Bn * Cl * hl * wl);
  if (B.err) return B.err;
  {
    Op op{};
    op.kind = OP_CONV_IN;
    op.ci = {P->x_in, Bn, Bn, 3, st.H, st.W, v->ecin_w, v->ecin_b, v->EC0, st.x()};
    P->ops.push_back(op);
    P->flops += 2.0 * Bn * st.H * st.W * 27.0 * v->EC0;
  }
  for (const auto& b : v->eblocks) {
    if (B.err) break;
    st.vres(b.r[0]);
    st.vres(b.r[1]);
    if (b.down) {
      // PaddedConv2d(3x3, stride 2, padding (left 0, right 1, top 0, bottom 1)), autoencoder/mod.rs:326-407: output (i, j) reads
      // input rows 2i..2i+2 / cols 2j..2j+2 with zeros past the bottom/right edge -> tap k: phase k&1, offset k>>1.
      Op op{};
      op.kind = OP_PHASE;
      op.rs = {st.x(), Bn, st.H, st.W, b.Cout, s_ph};
      P->ops.push_back(op);
      const int H2 = st.H / 2, W2 = st.W / 2;
      ActView a{s_ph, 4 * Bn, H2, W2, b.Cout};
      std::vector<IgemmSeg> segs;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
          segs.push_back({0, (int16_t)(kw >> 1), (int16_t)(kh >> 1), (int16_t)((((kh & 1) * 2) + (kw & 1)) * Bn), b.downc.Ipad / 64});
      B.igemm(a, nullptr, segs, b.downc.w, b.downc.O, b.downc.Ktot, H2, W2, Bn, IGEMM_LINEAR, 0, st.other(), 1, b.downc.O, b.downc.b, 0,
              nullptr, 0);
      B.add_flops(2.0 * Bn * H2 * W2 * 9.0 * b.Cout * b.downc.O);
      st.flip();
      st.H = H2; st.W = W2;
    }
  }
  if (B.err) return B.err;
  st.vres(v->emid1);
  st.attn(v->eattn_norm, v->eq, v->ek, v->ev, v->eproj);
  st.vres(v->emid2);
  B.gn(st.x(), Ce, nullptr, 0, st.H * st.W, v->enorm_out, 1, st.s_gn1, nullptr);
  {
    ActView a{st.s_gn1, Bn, st.H, st.W, Ce};
    B.conv3(a, nullptr, v->econv_out, v->enc_z, v->econv_out.b, 0, nullptr);
  }
  P->flops += 2.0 * Bn * st.H * st.W * (double)Cz * Cz;  // quant_conv (all Cz outputs in the reference)
  return B.err;
}

static int vae_encode_run(sdxl_vae* v, int Bn, int H, int W, const float* image, const uint8_t* rgb, int on_host, float* latent_out) {
  sdxl_ctx* c = v->ctx;
  if (!v->has_enc) return fail(c, 5104, "this sdxl_vae was loaded without the encoder half (n_enc_blocks = 0)");
  if (!latent_out || (!image && !rgb)) return fail(c, -1, "null argument");
  if (Bn < 1 || H < 1 || W < 1) return fail(c, 5100, "bad encode shape B=%d H=%d W=%d", Bn, H, W);
  CU(c, cudaSetDevice(c->device));
  if (!v->enc_plan || v->enc_plan->Bf != Bn || v->enc_plan->h != H || v->enc_plan->w != W) {
    CU(c, cudaStreamSynchronize(c->stream));
    v->enc_plan.reset(new Plan());
    Plan* P = v->enc_plan.get();
    P->Bf = Bn; P->Bx = Bn; P->h = H; P->w = W;
    Arena meas;
    meas.measure = true;
    int r = build_vae_enc_plan(v, P, &meas);
    if (!r && P->arena.init(meas.off + (1 << 20))) r = fail(c, 5011, "cannot allocate %zu bytes of workspace", meas.off);
    if (!r) r = build_vae_enc_plan(v, P, &P->arena);
    if (r) { v->enc_plan.reset(); return r; }
  }
  Plan* P = v->enc_plan.get();
  const size_t npix = (size_t)Bn * H * W;
  if (image) {
    CU(c, cudaMemcpyAsync(P->x_in, image, npix * 3 * sizeof(float), on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
  } else {
    const uint8_t* src = rgb;
    if (on_host) {
      CU(c, cudaMemcpyAsync(v->enc_u8, rgb, npix * 3, cudaMemcpyHostToDevice, c->stream));
      src = v->enc_u8;
    }
    KL(c, image_from_u8_launch(c->stream, src, Bn, (long)H * W, P->x_in));
  }
  int r = run_plan_ops(c, P);
  if (r) return r;
  const int down = 1 << (v->cfg.n_enc_blocks - 1);
  const long hw = (long)(H / down) * (W / down);
  float* dst = on_host ? v->enc_lat : latent_out;
  KL(c, quant_out_launch(c->stream, v->enc_z, Bn, v->cfg.enc_z_channels, v->cfg.latent_channels, hw, v->qc_w, v->qc_b,
                         (float)v->cfg.scale_factor, dst));
  if (on_host) {
    CU(c, cudaMemcpyAsync(latent_out, dst, (size_t)Bn * v->cfg.latent_channels * hw * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return 0;
}
extern "C" int sdxl_vae_encode_image(sdxl_vae* v, int Bn, int H, int W, const float* image, int on_host, float* latent_out) {
  if (!v || !image) return fail(v ? v->ctx : nullptr, -1, "sdxl_vae_encode_image: null argument");
  return vae_encode_run(v, Bn, H, W, image, nullptr, on_host, latent_out);
}
extern "C" int sdxl_vae_image_to_latent(sdxl_vae* v, int Bn, int H, int W, const uint8_t* rgb, int on_host, float* latent_out) {
  if (!v || !rgb) return fail(v ? v->ctx : nullptr, -1, "sdxl_vae_image_to_latent: null argument");
  return vae_encode_run(v, Bn, H, W, nullptr, rgb, on_host, latent_out);
}
extern "C" double sdxl_vae_encode_plan_flops(const sdxl_vae* v) { return (v && v->enc_plan) ? v->enc_plan->flops : 0.0; }

static int vae_ensure_plan(sdxl_vae* v, int Bn, int h, int w) {
  sdxl_ctx* c = v->ctx;
  if (Bn < 1 || h < 1 || w < 1) return fail(c, 5100, "bad decode shape B=%d h=%d w=%d", Bn, h, w);
  if (v->plan && v->plan->Bf == Bn && v->plan->h == h && v->plan->w == w) return 0;
  CU(c, cudaStreamSynchronize(c->stream));
  v->plan.reset(new Plan());
  Plan* P = v->plan.get();
  P->Bf = Bn; P->Bx = Bn; P->h = h; P->w = w;
  Arena meas;
  meas.measure = true;
  int r = build_vae_plan(v, P, &meas);
  if (r) { v->plan.reset(); return r; }
  if (P->arena.init(meas.off + (1 << 20))) { v->plan.reset(); return fail(c, 5011, "cannot allocate %zu bytes of workspace", meas.off); }
  r = build_vae_plan(v, P, &P->arena);
  if (r) { v->plan.reset(); return r; }
  return 0;
}

static int vae_run(sdxl_vae* v, int Bn, int h, int w, const float* latent, int on_host) {
  sdxl_ctx* c = v->ctx;
  if (!latent) return fail(c, -1, "null latent");
  CU(c, cudaSetDevice(c->device));
  int r = vae_ensure_plan(v, Bn, h, w);
  if (r) return r;
  Plan* P = v->plan.get();
  const size_t n = (size_t)Bn * v->cfg.latent_channels * h * w;
  CU(c, cudaMemcpyAsync(P->x_in, latent, n * sizeof(float), on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
  return run_plan_ops(c, P);
}

extern "C" int sdxl_vae_decode_latent(sdxl_vae* v, int Bn, int h, int w, const float* latent, int on_host, float* image_out) {
  if (!v || !image_out) return fail(v ? v->ctx : nullptr, -1, "sdxl_vae_decode_latent: null argument");
  sdxl_ctx* c = v->ctx;
  int r = vae_run(v, Bn, h, w, latent, on_host);
  if (r) return r;
  const int up = 1 << (v->cfg.n_blocks - 1);
  const int HW = h * up * w * up;
  float* dst = on_host ? v->out_f32 : image_out;
  KL(c, nhwc_to_nchw_f32_launch(c->stream, v->img_nhwc, Bn, HW, 3, 4, dst));
  if (on_host) {
    CU(c, cudaMemcpyAsync(image_out, dst, (size_t)Bn * 3 * HW * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return 0;
}
extern "C" int sdxl_vae_latent_to_image(sdxl_vae* v, int Bn, int h, int w, const float* latent, int on_host, uint8_t* rgb_out) {
  if (!v || !rgb_out) return fail(v ? v->ctx : nullptr, -1, "sdxl_vae_latent_to_image: null argument");
  sdxl_ctx* c = v->ctx;
  int r = vae_run(v, Bn, h, w, latent, on_host);
  if (r) return r;
  const int up = 1 << (v->cfg.n_blocks - 1);
  const long npix = (long)Bn * h * up * w * up;
  uint8_t* dst = on_host ? v->out_u8 : rgb_out;
  KL(c, image_u8_launch(c->stream, v->img_nhwc, npix, 4, dst));
  if (on_host) {
    CU(c, cudaMemcpyAsync(rgb_out, dst, (size_t)npix * 3, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return 0;
}
extern "C" double sdxl_vae_plan_flops(const sdxl_vae* v) { return (v && v->plan) ? v->plan->flops : 0.0; }
extern "C" int sdxl_vae_profile_plan(sdxl_vae* v, double* ms_by_kind, double* flops_by_kind, int* launches_by_kind) {
  if (!v || !v->plan) return -1;
  return profile_plan_impl(v->ctx, v->plan.get(), ms_by_kind, flops_by_kind, launches_by_kind);
}
extern "C" int sdxl_vae_profile_dump(sdxl_vae* v, const char* path) {
  if (!v || !v->plan || !path) return -1;
  return profile_dump_impl(v->ctx, v->plan.get(), path);
}


// ================================================================================================
// Text encoders of the Embedder: CLIP::{forward_hidden, forward_hidden_pooled}
//   CLIP / ResidualDecoderAttentionBlock / MultiHeadSelfAttention / MLP / QuickGELU
//                                   src/model/clip/mod.rs:82-147, 176-182, 228-245, 289-305, 315-319
//   weight names                    src/model/clip/load.rs:15-115
// 77-token sequences: Linear layers on the tcgen05 GEMM (one M tile), causal attention / activation / embedding on
// small CUDA-core kernels (clip_kernels.cu). Residual stream f32, GEMM operands f16 (the reference runs f32).
// ================================================================================================
struct CBlock {
  Norm attn_ln, mlp_ln;
  Lin qkv, out, fc1, fc2;
};
struct sdxl_clip {
  sdxl_ctx* ctx = nullptr;
  sdxl_clip_cfg cfg{};
  Arena warena;
  __half* tok_emb = nullptr;
  __half* pos_emb = nullptr;
  std::vector<CBlock> blocks;
  Norm ln_final;
  Lin proj;
  bool has_proj = false;
  // plan (keyed by batch, number of blocks run, captured hidden index, pooled)
  std::unique_ptr<Plan> plan;
  int pB = 0, p_nrun = 0, p_hidden = -1, p_pooled = 0;
  int* tokens_dev = nullptr;
  int* eot_dev = nullptr;
  int* err_dev = nullptr;
  float* hidden = nullptr;   // [B*T, C] result of forward_hidden / h_out
  float* pooled = nullptr;   // [B, embed_dim]
};

static int build_clip(sdxl_clip* m, const PackView& pv, Arena& A) {
  sdxl_ctx* c = m->ctx;
  const sdxl_clip_cfg& g = m->cfg;
  Loader L{nullptr, c, &pv, &A, c->stream};
  const int C = g.n_state;
  m->blocks.clear();
  auto table = [&](const std::string& name, int rows, __half*& dst) {
    const PackEntry* e = L.need(name, 2);
    if (!e) return;
    if ((int)e->shape[0] != rows || (int)e->shape[1] != C) { L.err = fail(c, 4401, "weight pack: '%s' is [%llu,%llu], expected [%d,%d]", name.c_str(), (unsigned long long)e->shape[0], (unsigned long long)e->shape[1], rows, C); return; }
    dst = A.get<__half>((size_t)rows * C);
    if (!dst) { L.err = fail(c, 4005, "weight arena exhausted"); return; }
    if (!A.measure && cudaMemcpyAsync(dst, L.ptr(e), (size_t)rows * C * sizeof(__half), cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess)
      L.err = fail(c, 4402, "embedding copy failed");
  };
  table("token_embedding/weight", g.n_vocab, m->tok_emb);
  table("position_embedding/weight", g.n_ctx, m->pos_emb);
  if (L.err) return L.err;
  const int Cpad = Loader::pad64(C);
  for (int i = 0; i < g.n_layer && !L.err; ++i) {
    const std::string bp = "blocks/" + std::to_string(i);
    CBlock b;
    b.attn_ln = L.norm(bp + "/attn_ln", C);
    b.mlp_ln = L.norm(bp + "/mlp_ln", C);
    // fused q/k/v projection (clip/mod.rs:229-231: three Linears with bias on the same input)
    b.qkv.K = C; b.qkv.Kpad = Cpad; b.qkv.N = 3 * C;
    b.qkv.w = A.get<__half>((size_t)3 * C * Cpad);
    b.qkv.b = A.get<float>((size_t)3 * C);
    if (!b.qkv.w || !b.qkv.b) { L.err = fail(c, 4005, "weight arena exhausted"); break; }
    const char* names[3] = {"query", "key", "value"};
    for (int j = 0; j < 3 && !L.err; ++j) {
      const std::string lp = bp + "/attn/" + names[j];
      L.lin_into(lp, b.qkv.w, Cpad, j * C, C, C, 0);
      const PackEntry* be = L.need(lp + "/bias", 1);
      if (!be) break;
      if ((int)be->shape[0] != C) { L.err = fail(c, 4403, "weight pack: '%s/bias' mis-sized", lp.c_str()); break; }
      if (!A.measure) { int r = bias_to_f32_launch(c->stream, L.ptr(be), C, b.qkv.b + j * C, 0, 0); if (r) L.err = fail(c, r, "bias_to_f32 failed"); }
    }
    b.out = L.linear(bp + "/attn/out", C, C, true);
    b.fc1 = L.linear(bp + "/mlp/fc1", C, 4 * C, true);
    b.fc2 = L.linear(bp + "/mlp/fc2", 4 * C, C, true);
    m->blocks.push_back(b);
  }
  if (L.err) return L.err;
  m->ln_final = L.norm("layer_norm", C);
  m->has_proj = pv.find("text_projection") != nullptr;
  if (m->has_proj) {
    const PackEntry* e = L.need("text_projection", 2);
    if (!e) return L.err;
    if ((int)e->shape[0] != C || (int)e->shape[1] != g.embed_dim) return fail(c, 4404, "text_projection is [%llu,%llu], expected [%d,%d]", (unsigned long long)e->shape[0], (unsigned long long)e->shape[1], C, g.embed_dim);
    Lin& P = m->proj;
    P.K = C; P.Kpad = Cpad; P.N = g.embed_dim;
    P.w = A.get<__half>((size_t)g.embed_dim * Cpad);
    if (!P.w) return fail(c, 4005, "weight arena exhausted");
    if (!A.measure) { int r = transpose_linear_launch(c->stream, L.ptr(e), C, g.embed_dim, P.w, Cpad, 0, 0); if (r) return fail(c, r, "text_projection re-layout failed"); }
  }
  return L.err;
}

extern "C" void sdxl_clip_destroy(sdxl_clip* m) {
  if (!m) return;
  cudaStreamSynchronize(m->ctx->stream);
  m->plan.reset();
  m->warena.release();
  if (m->tokens_dev) cudaFree(m->tokens_dev);
  delete m;
}

extern "C" int sdxl_clip_load(sdxl_ctx* c, const sdxl_clip_cfg* cfg, const void* pack, size_t bytes, int pack_on_device, sdxl_clip** out) {
  if (!c || !cfg || !pack || !out) return fail(c, -1, "sdxl_clip_load: null argument");
  *out = nullptr;
  if (cfg->n_head < 1 || cfg->n_state != cfg->n_head * 64) return fail(c, 4410, "text encoder head dim must be 64 (n_state=%d, n_head=%d)", cfg->n_state, cfg->n_head);
  if (cfg->n_ctx < 1 || cfg->n_ctx > 1024 || cfg->n_layer < 1 || cfg->n_vocab < 1 || cfg->embed_dim < 1) return fail(c, 4411, "bad text encoder config");
  CU(c, cudaSetDevice(c->device));
  std::unique_ptr<sdxl_clip> m(new sdxl_clip());
  m->ctx = c;
  m->cfg = *cfg;
  PackView pv;
  std::vector<uint8_t> table;
  int r = parse_pack(c, pack, bytes, pack_on_device, pv, table);
  if (r) return r;
  void* dev_pack = nullptr;
  if (pack_on_device) {
    pv.dev = (const uint8_t*)pack;
  } else {
    CU(c, cudaMalloc(&dev_pack, bytes));
    cudaError_t e = cudaMemcpyAsync(dev_pack, pack, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { cudaFree(dev_pack); return fail(c, (int)e, "pack upload failed"); }
    pv.dev = (const uint8_t*)dev_pack;
  }
  Arena meas;
  meas.measure = true;
  r = build_clip(m.get(), pv, meas);
  if (!r && m->warena.init(meas.off + (1 << 20))) r = fail(c, 4203, "cannot allocate %zu bytes for weights", meas.off);
  if (!r) r = build_clip(m.get(), pv, m->warena);
  cudaError_t se = cudaStreamSynchronize(c->stream);
  if (dev_pack) cudaFree(dev_pack);
  if (!r && se != cudaSuccess) r = fail(c, (int)se, "weight re-layout failed: %s", cudaGetErrorString(se));
  if (!r && cudaMalloc((void**)&m->tokens_dev, (size_t)(64 * cfg->n_ctx + 64 + 16) * sizeof(int)) != cudaSuccess) r = fail(c, 4412, "cudaMalloc failed");
  if (r) { m->warena.release(); return r; }
  m->eot_dev = m->tokens_dev + 64 * cfg->n_ctx;
  m->err_dev = m->eot_dev + 64;
  *out = m.release();
  return 0;
}

// n_run blocks are executed; when capture >= 0 the stream entering block `capture` is preserved as the hidden output.
static int build_clip_plan(sdxl_clip* m, Plan* P, Arena* A, int n_run, int capture, int pooled) {
  sdxl_ctx* c = m->ctx;
  const sdxl_clip_cfg& g = m->cfg;
  PlanBuilder B{nullptr, c, P, A, P->Bf};
  P->ops.clear();
  P->flops = 0;
  const int Bn = P->Bf, T = g.n_ctx, C = g.n_state, M = Bn * T;
  float* xa = B.buf<float>((size_t)M * C);
  float* xb = B.buf<float>((size_t)M * C);
  __half* a16 = B.buf<__half>((size_t)M * C);
  __half* qkv16 = B.buf<__half>((size_t)M * 3 * C);
  __half* ao16 = B.buf<__half>((size_t)M * C);
  float* h32 = B.buf<float>((size_t)M * 4 * C);
  __half* h16 = B.buf<__half>((size_t)M * 4 * C);
  float* pin = B.buf<float>((size_t)Bn * C);
  m->pooled = B.buf<float>((size_t)Bn * g.embed_dim);
  if (B.err) return B.err;
  {
    Op op{};
    op.kind = OP_EMBED;
    op.em = {m->tokens_dev, M, T, C, g.n_vocab, m->tok_emb, m->pos_emb, xa, m->err_dev};
    P->ops.push_back(op);
  }
  float* x = xa;
  m->hidden = nullptr;
  for (int i = 0; i < n_run && !B.err; ++i) {
    const CBlock& b = m->blocks[i];
    float* xn = x;
    if (i == capture) {  // keep the input of this block: write the updated stream into the other buffer
      m->hidden = x;
      xn = (x == xa) ? xb : xa;
    }
    // x = x + attn(attn_ln(x), causal mask)    (clip/mod.rs:177-179)
    B.ln(x, b.attn_ln, M, a16);
    B.linear(a16, M, b.qkv, IGEMM_LINEAR, qkv16, 0, 3 * C, nullptr, 0);
    {
      Op op{};
      op.kind = OP_ATTN_SMALL;
      op.as = {qkv16, 3 * C, 0, qkv16, qkv16, 3 * C, C, 2 * C, Bn, T, T, g.n_head, nullptr, 1, ao16, C};
      P->ops.push_back(op);
      B.add_flops(4.0 * Bn * T * (double)T * C);
    }
    B.linear(ao16, M, b.out, IGEMM_LINEAR, xn, 1, C, x, C);
    // x = x + mlp(mlp_ln(x))
    B.ln(xn, b.mlp_ln, M, a16);
    B.linear(a16, M, b.fc1, IGEMM_LINEAR, h32, 1, 4 * C, nullptr, 0);
    {
      Op op{};
      op.kind = OP_ACT;
      op.ac = {h32, (size_t)M * 4 * C, g.quick_gelu ? 1 : 0, h16};
      P->ops.push_back(op);
    }
    B.linear(h16, M, b.fc2, IGEMM_LINEAR, xn, 1, C, xn, C);
    x = xn;
  }
  if (capture < 0 || capture >= n_run) m->hidden = x;
  if (pooled && !B.err) {
    // features of the end-of-text position: layer_norm(x)[b, argmax(tokens[b])] (@ text_projection)   (clip/mod.rs:130-141)
    Op op{};
    op.kind = OP_LN_GATHER;
    op.lg = {x, m->eot_dev, Bn, T, C, m->ln_final.g, m->ln_final.b, m->ln_final.eps, m->has_proj ? pin : m->pooled};
    P->ops.push_back(op);
    if (m->has_proj) B.gemv(pin, C, Bn, m->proj, nullptr, 0, 0, 0, m->pooled, g.embed_dim);
  }
  return B.err;
}

static int clip_run(sdxl_clip* m, int Bn, const int32_t* tokens_host, int n_run, int capture, int pooled) {
  sdxl_ctx* c = m->ctx;
  const sdxl_clip_cfg& g = m->cfg;
  if (!tokens_host) return fail(c, -1, "null tokens");
  if (Bn < 1 || Bn > 64) return fail(c, 5201, "text encoder batch must be 1..64 (got %d)", Bn);
  if (n_run < 0 || n_run > g.n_layer) return fail(c, 5202, "hidden_idx %d out of range (n_layer %d)", n_run, g.n_layer);
  CU(c, cudaSetDevice(c->device));
  if (!m->plan || m->pB != Bn || m->p_nrun != n_run || m->p_hidden != capture || m->p_pooled != pooled) {
    CU(c, cudaStreamSynchronize(c->stream));
    m->plan.reset(new Plan());
    Plan* P = m->plan.get();
    P->Bf = Bn; P->Bx = Bn;
    Arena meas;
    meas.measure = true;
    int r = build_clip_plan(m, P, &meas, n_run, capture, pooled);
    if (!r && P->arena.init(meas.off + (1 << 20))) r = fail(c, 5011, "cannot allocate %zu bytes of workspace", meas.off);
    if (!r) r = build_clip_plan(m, P, &P->arena, n_run, capture, pooled);
    if (r) { m->plan.reset(); return r; }
    m->pB = Bn; m->p_nrun = n_run; m->p_hidden = capture; m->p_pooled = pooled;
  }
  // eot_indices = tokens.argmax(1): first position of the largest id (clip/mod.rs:130)
  std::vector<int> meta(64 + 1, 0);
  for (int b = 0; b < Bn; ++b) {
    int best = 0;
    for (int t = 1; t < g.n_ctx; ++t)
      if (tokens_host[b * g.n_ctx + t] > tokens_host[b * g.n_ctx + best]) best = t;
    meta[b] = best;
  }
  CU(c, cudaMemcpyAsync(m->tokens_dev, tokens_host, (size_t)Bn * g.n_ctx * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(m->eot_dev, meta.data(), 65 * sizeof(int), cudaMemcpyHostToDevice, c->stream));  // also clears err_dev
  CU(c, cudaStreamSynchronize(c->stream));  // meta / tokens_host are pageable host memory
  int r = run_plan_ops(c, m->plan.get());
  if (r) return r;
  int err = 0;
  CU(c, cudaMemcpyAsync(&err, m->err_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (err) return fail(c, 5203, "token id outside [0, %d) (the reference's embedding lookup panics)", g.n_vocab);
  return 0;
}

static int clip_copy_out(sdxl_clip* m, const float* src, size_t n, float* dst, int on_host) {
  sdxl_ctx* c = m->ctx;
  if (!dst) return 0;
  CU(c, cudaMemcpyAsync(dst, src, n * sizeof(float), on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
  if (on_host) CU(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int sdxl_clip_forward_hidden(sdxl_clip* m, int Bn, const int32_t* tokens_host, int hidden_idx, float* hidden_out, int out_on_host) {
  if (!m || !hidden_out) return fail(m ? m->ctx : nullptr, -1, "sdxl_clip_forward_hidden: null argument");
  int r = clip_run(m, Bn, tokens_host, hidden_idx, -1, 0);
  if (r) return r;
  return clip_copy_out(m, m->hidden, (size_t)Bn * m->cfg.n_ctx * m->cfg.n_state, hidden_out, out_on_host);
}
extern "C" int sdxl_clip_forward_hidden_pooled(sdxl_clip* m, int Bn, const int32_t* tokens_host, int hidden_idx, float* hidden_out,
                                               float* pooled_out, int out_on_host) {
  if (!m || !hidden_out || !pooled_out) return fail(m ? m->ctx : nullptr, -1, "sdxl_clip_forward_hidden_pooled: null argument");
  if (hidden_idx < 0 || hidden_idx >= m->cfg.n_layer) return fail(m->ctx, 5204, "hidden_idx %d out of range: the reference returns an uninitialised tensor there (clip/mod.rs:120-126)", hidden_idx);
  int r = clip_run(m, Bn, tokens_host, m->cfg.n_layer, hidden_idx, 1);
  if (r) return r;
  r = clip_copy_out(m, m->hidden, (size_t)Bn * m->cfg.n_ctx * m->cfg.n_state, hidden_out, out_on_host);
  if (r) return r;
  return clip_copy_out(m, m->pooled, (size_t)Bn * (m->has_proj ? m->cfg.embed_dim : m->cfg.n_state), pooled_out, out_on_host);
}
extern "C" double sdxl_clip_plan_flops(const sdxl_clip* m) { return (m && m->plan) ? m->plan->flops : 0.0; }
