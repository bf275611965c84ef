// Host-side launch API of the sm_100a kernels (internal to libsdxl_b200.so; the public boundary is
// include/sdxl_b200.h). All pointers are device pointers. All launchers return cudaError_t-style
// int (0 = ok) and never synchronise.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdxl {

// Launch helper: cudaLaunchKernelEx with optional programmatic dependent launch (PDL). Kernels launched
// with pdl=true MUST execute griddep_wait() (common.cuh) before touching global memory written by the
// preceding kernel. SDXL_B200_NO_PDL=1 disables the attribute globally.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline int launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                         Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline int launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 int cluster_size, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_size > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_size;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl && pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return (int)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Opt-in to more than 48 KB of dynamic shared memory for `kernel`. The attribute is per DEVICE (one process may drive several
// devices through several contexts), so call sites keep one flag per device: `static bool done[64]`.
template <typename K>
inline int smem_optin(K kernel, int bytes, bool (&done)[64]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64) return 2001;
  if (!done[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    done[dev] = true;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Implicit GEMM on tcgen05 (igemm.cu): out[pixel, n] = epilogue( sum_seg sum_c A_seg[pixel+tap, c] *
// Wt[n, k(seg,c)] ). Linear layers are the 1-segment / 1x1 case of the same kernel.
// ------------------------------------------------------------------------------------------------
constexpr int IGEMM_MAX_SEG = 20;
struct IgemmSeg {
  int16_t map;   // 0 -> tmA0, 1 -> tmA1
  int16_t dw, dh;  // tap offset added to the tile's (w0,h0)
  int16_t db;    // offset added to the tile's batch coordinate (stride-2 phase images)
  int32_t nkb;   // number of 64-channel K blocks in this segment
};
enum IgemmMode : int { IGEMM_LINEAR = 0, IGEMM_GEGLU = 1 };
struct alignas(64) IgemmParams {
  CUtensorMap tmA0, tmA1, tmB;
  IgemmSeg seg[IGEMM_MAX_SEG];
  int nseg;
  int Wt, Ht, Bt;             // A box in pixels, Wt*Ht*Bt == 128
  int W, H, Bn;               // output extents
  int tilesW, tilesH, tilesB, tilesN;
  int pair;                   // 1: 2-CTA MMA kernel (cta_group::2), CM = 2, CN = 1
  int CM, CN;                 // cluster shape (M x N CTAs sharing operand tiles via TMA multicast)
  int a_split_dim, a_split_ext;  // how the A tile is sliced across the CN peers (0=W,1=H,2=B; extent per slice)
  int N;                      // valid output columns (GEGLU: columns of the fused [value|gate] GEMM)
  int BN;                     // N tile (multiple of 16, <= 256)
  int nstages;
  int mode;                   // IgemmMode
  void* out;                  // f16 or f32 [pixels, ldo]
  int out_f32;
  int ldo;
  const float* bias;          // nullable; index b*bias_bstride + n
  int bias_bstride;
  const float* res;           // nullable f32 residual [pixels, ldr]
  int ldr;
  // output pixel of tile pixel (b, h, w): (b*H + h) * opix_row + w * opix_w + opix_off (defaults W, 1, 0). The nearest-2x
  // upsample + 3x3 conv is run as four 2x2 convolutions on the original image, each writing one (row, column) parity of the
  // upsampled output: opix_row = 4W, opix_w = 2, opix_off = a*2W + b.
  int opix_row, opix_w, opix_off;
  // TMA epilogue (igemm.cu "epilogue through TMA"): set by igemm_configure when the tile's 128 rows are contiguous rows of a plain
  // [pixels, ldo] output (N and BN multiples of 32, LINEAR mode); igemm_launch drops it if the caller re-mapped the output pixels.
  CUtensorMap tmOut, tmRes;   // 2-D [pixels, ldo] views, box 32 rows x 32 columns (f32: SWIZZLE_128B, f16: SWIZZLE_64B)
  int epi_tma;
  int epi_box_bytes;          // bytes of one epilogue box in shared memory: 4096 (32 rows x 128 B), or 2048 for f16 outputs on the TMA epilogue
  // host-computed reciprocals (floor(2^32/d)+1; q = umulhi(n, m), exact while n*d < 2^32; 0 = use '/') for the tile-index
  // divisions of the producer warp: on the critical path between griddepcontrol.wait and the first TMA issue
  unsigned fd_pm, fd_w, fd_h, fd_wh;   // divisors: pair M tiles (or M tiles), tilesW, tilesH, tilesW*tilesH
  int dbg_mode;               // diagnostics only: 1 = skip TMA loads, 2 = skip MMAs (results are garbage)
  unsigned long long* dbg;    // nullable: per-role %globaltimer stamps of CTA 0 (tools/igemm_timeline.py)
  unsigned long long* dbg_all;  // nullable: [gridDim.x][4] stamps of EVERY CTA (entry, prologue done, dependencies resolved, exit): launch ramp / drain
};
// A operand view: NHWC f16 tensor [Bn, H, W, C] with channel pitch `pitch` (elements, multiple of 8).
int make_tmap_act(CUtensorMap* tm, const __half* base, int Bn, int H, int W, int C, int pitch, int Wt,
                  int Ht, int Bt);
// B operand view: K-major weights [N, K] f16 with row pitch K (multiple of 8); box (64, BN).
int make_tmap_wgt(CUtensorMap* tm, const __half* base, int N, int K, int BN);
// Picks Wt/Ht/Bt (product 128) for an output image of W x H.
void igemm_pick_box(int W, int H, int* Wt, int* Ht, int* Bt);
// Operand views for igemm_configure: NHWC f16 activations (channel pitch in elements) and K-major weights.
struct IgemmOperands {
  const __half* a0; int a0Bn, a0H, a0W, a0C, a0pitch;
  const __half* a1; int a1Bn, a1H, a1W, a1C, a1pitch;   // nullable second source (segments with map == 1)
  const __half* w; int N, Ktot;
};
// Chooses the pixel box, N tile, cluster shape and pipeline depth, and builds the TMA descriptors. The caller
// fills seg[]/nseg and the epilogue fields (out, out_f32, ldo, bias, bias_bstride, res, ldr).
int igemm_configure(IgemmParams& p, const IgemmOperands& o, int outW, int outH, int outB, int mode, int geglu_bn);
int igemm_launch(cudaStream_t st, IgemmParams& p);
// Chooses an N tile for (M pixels, N columns) minimising wave-quantisation loss on `num_sms` SMs.
int igemm_pick_bn(int m_tiles, int N, int num_sms, bool geglu);

// ------------------------------------------------------------------------------------------------
// Attention (attention.cu): out[b, t, h*64:(h+1)*64] = softmax(q k^T / 8) v, head dim 64.
// q/k/v are column windows of row-major f16 matrices (fused QKV / KV GEMM outputs).
// ------------------------------------------------------------------------------------------------
struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;  // 3D maps (64 cols, rows, batch)
  int T, S, n_head, B;
  int q_col0, k_col0, v_col0;  // column of head 0 inside each matrix
  __half* out;                 // [B*T, ldo]
  int ldo;
  float scale_log2e;           // (1/sqrt(d)) * log2(e)
  long long* dbg;              // nullable: clock64 stamps of CTA 0 (sdxl_dbg_attention_timeline)
};
int make_tmap_rows(CUtensorMap* tm, const __half* base, int rows_per_batch, int nbatch, int cols, int pitch);
int attention_launch(cudaStream_t st, const AttnParams& p);

// ------------------------------------------------------------------------------------------------
// Norms (norm.cu)
// ------------------------------------------------------------------------------------------------
// GroupNorm over NHWC f32 input (optionally the channel-concatenation of two tensors), 32 groups,
// biased variance, eps inside the sqrt (reference groupnorm/mod.rs:52-82). Writes f16.
struct GnParams {
  const float* x1; int C1;    // [B, HW, C1]
  const float* x2; int C2;    // nullable, [B, HW, C2]  (cat([x1, x2], channel))
  int B, HW, n_group;
  const float* gamma; const float* beta;  // [C1+C2]
  float eps;
  int silu;                   // apply x*sigmoid(x) after the affine
  __half* y;                  // [B, HW, C1+C2] normalised (+SiLU) output
  __half* raw;                // nullable: un-normalised f16 copy of cat([x1,x2]) (skip-conv operand)
  float* partial;             // scratch of gn_scratch_floats(B, n_group) floats, initialised once with gn_scratch_init
  int nchunk;                 // filled by gn_launch
  __half* y_lo;               // nullable: f16(t - float(y)), the rounding residue of y (hi/lo split operand of the UNet's last conv)
};
size_t gn_scratch_floats(int B, int n_group);
// zeroes the arrival counters of a freshly allocated scratch (once; the kernels leave them at zero)
int gn_scratch_init(cudaStream_t st, float* scratch, int B, int n_group);
int gn_launch(cudaStream_t st, GnParams& p);
// LayerNorm over the last dim of f32 [rows, C] -> f16 [rows, C] (reference layernorm/mod.rs:34-49).
int layernorm_launch(cudaStream_t st, const float* x, const float* gamma, const float* beta, float eps,
                     int rows, int C, __half* y);

// ------------------------------------------------------------------------------------------------
// Small / elementwise kernels (elementwise.cu)
// ------------------------------------------------------------------------------------------------
// out[b,n] = act( dot(in[b,:], W[n,:]) + bias[n] + add[b*add_bstride + n] ); W f16 [N,ldw] K-major (row pitch ldw >= K).
// in_silu: apply SiLU to the input on load. out_silu: SiLU on the output.
int gemv_launch(cudaStream_t st, const float* in, int in_bstride, int Bv, int K, const __half* W, int ldw,
                const float* bias, const float* add, int add_bstride, int N, int in_silu, int out_silu,
                float* out, int out_bstride);
// timestep_embedding (reference unet/mod.rs:21-39): out[b, :] = [cos(t*f_i), sin(t*f_i)], dim even.
int timestep_embedding_launch(cudaStream_t st, const int* t_dev, int nt, int dim, float max_period,
                              float* out);
// First conv: x NCHW f16 [B,Cin,H,W] (Cin<=8) -> NHWC f32 [B,H,W,Cout], 3x3 pad 1. w: [Cout][3][3][Cin] f32.
int conv_in_launch(cudaStream_t st, const __half* x, int B, int Cin, int H, int W, const float* w,
                   const float* bias, int Cout, float* y);
// general form: x f16 or f32 NCHW with Bx images; output batch b reads image b % Bx.
int conv_in_launch_t(cudaStream_t st, const void* x, int x_f32, int Bx, int B, int Cin, int H, int W,
                     const float* w, const float* bias, int Cout, float* y);
// Nearest-2x upsample of NHWC: f32 [B,H,W,C] -> f16 [B,2H,2W,C] (reference unet/mod.rs:742-751).
int upsample2x_launch(cudaStream_t st, const float* x, int B, int H, int W, int C, __half* y);
// Stride-2 phase split: f32 [B,H,W,C] -> f16 [4(phase=ph*2+pw), B, H/2, W/2, C]; phase image
// P[ph][pw][i][j] = x[2i+ph][2j+pw].
int phase_split_launch(cudaStream_t st, const float* x, int B, int H, int W, int C, __half* y);
int cast_f32_to_f16_launch(cudaStream_t st, const float* x, size_t n, __half* y);
int cast_f16_to_f32_launch(cudaStream_t st, const __half* x, size_t n, float* y);
// eps NHWC f32 [B,HW,ldx] (first C valid) -> NCHW [B,C,HW], f16 or f32 output
int nhwc_to_nchw_f16_launch(cudaStream_t st, const float* x, int B, int HW, int C, int ldx, __half* y);
int nhwc_to_nchw_f32_launch(cudaStream_t st, const float* x, int B, int HW, int C, int ldx, float* y);
// Sampler elementwise (reference stablediffusion/mod.rs:407-428, 463-465, 539-540).
// eps layout: NHWC f32 [nfwd*Bimg, HW, ld]; cond rows first, then uncond (if cfg).
// x: NCHW f32 master latent [Bimg,C,HW], updated in place; x16: f16 NCHW copy duplicated for the next
// forward ([nfwd*Bimg, C, HW]).
int cfg_ddim_launch(cudaStream_t st, const float* eps, int ld, int Bimg, int C, int HW, int use_cfg,
                    float guidance, float sqrt_a, float sqrt_1ma, float sqrt_ap, float sqrt_1map,
                    float* x, __half* x16);
// x = mask ? x : ref*sqrt_a + noise*sqrt_1ma ; also refreshes x16 (both forwards).
int inpaint_blend_launch(cudaStream_t st, float* x, const float* ref, const float* noise,
                         const uint8_t* mask, size_t n_per_img_batch, int nfwd, float sqrt_a,
                         float sqrt_1ma, __half* x16);
// x = x*sa + noise*sb  (refine_latent entry, reference stablediffusion/mod.rs:363-367)
int axpby_launch(cudaStream_t st, float* x, const float* noise, size_t n, float sa, float sb);
// Standard normal noise, Philox4x32-10 + Box-Muller, element i of stream (seed, subseq).
int randn_launch(cudaStream_t st, float* out, size_t n, uint64_t seed, uint64_t subseq);
int dup_latent_f16_launch(cudaStream_t st, const float* x, size_t n, int nfwd, __half* x16);

// Latent-decoder kernels (vae_kernels.cu)
// P[r,:] = softmax(scale * S[r,:]); S f32 [rows, lds] -> P f16 [rows, ldp]; cols % 4 == 0.
int softmax_rows_launch(cudaStream_t st, const float* S, size_t lds, int rows, int cols, float scale, __half* P,
                        size_t ldp);
// y[c, r] = x[r, c]; x f16 [rows, ldx], y f16 [cols, ldy].
int transpose_f16_launch(cudaStream_t st, const __half* x, size_t ldx, int rows, int cols, __half* y, size_t ldy);
// 1x1 conv on the rescaled latent: y = W (x * inv_scale) + b; NCHW f32 [B, C<=8, HW]; W f32 [C, C].
int post_quant_launch(cudaStream_t st, const float* x, int B, int C, int HW, const float* w, const float* bias,
                      float inv_scale, float* y);
// u8[b,p,c] = trunc(clamp(((x+1)/2)*255, 0, 255)), c < 3, from NHWC f32 [npix, ldx].
int image_u8_launch(cudaStream_t st, const float* x, long npix, int ldx, uint8_t* out);

// u8 [B,HW,3] -> f32 NCHW [B,3,HW], ((v/255)*2)-1.
int image_from_u8_launch(cudaStream_t st, const uint8_t* in, int B, long HW, float* out);
// quant_conv on the first Cout output channels, times scale: x NHWC f32 [B,HW,Cz] -> y NCHW f32 [B,Cout,HW]; w f32 [Cz,Cz].
int quant_out_launch(cudaStream_t st, const float* x, int B, int Cz, int Cout, long HW, const float* w, const float* bias,
                     float scale, float* y);

// Text-encoder kernels (clip_kernels.cu)
// x[b*T+t,:] = tok_emb[tokens[b,t],:] + pos_emb[t,:] (f16 tables -> f32); *err is set to 1 on an id outside [0, n_vocab).
int embed_tokens_launch(cudaStream_t st, const int* tokens, int rows, int T, int C, int n_vocab, const __half* tok_emb,
                        const __half* pos_emb, float* x, int* err);
// Masked attention for short sequences, head dim 64: additive f16 mask [T,S] (nullable) and/or causal (key <= query).
int attention_small_launch(cudaStream_t st, const __half* q, int q_pitch, int q_col0, const __half* k, const __half* v,
                           int kv_pitch, int k_col0, int v_col0, int B, int T, int S, int n_head, const __half* mask,
                           int causal, __half* out, int ldo);
// y = gelu_erf(x) (quick = 0) or x * sigmoid(1.702 x) (quick = 1); f32 -> f16, n % 4 == 0.
int mlp_act_launch(cudaStream_t st, const float* x, size_t n, int quick, __half* y);
// y[b,:] = LayerNorm(x[b*T + idx[b], :]) in f32.
int ln_gather_f32_launch(cudaStream_t st, const float* x, const int* idx, int B, int T, int C, const float* gamma,
                         const float* beta, float eps, float* y);

// Weight re-layout at load time (elementwise.cu)
// Linear [K(in), N(out)] row-major f16 -> K-major [N, Kpad] f16 (zero padded), dst row pitch Kpad;
// rows written at dst_row0 + perm(n) where perm handles the GEGLU value/gate interleave (geglu_bn>0).
int transpose_linear_launch(cudaStream_t st, const __half* src, int K, int N, __half* dst, int Kpad,
                            int dst_row0, int geglu_bn);
// Conv OIHW f16 -> [O, (kh,kw,I padded to Ipad)] f16 at column offset col0 of a [O, Ktot] matrix.
int repack_conv_launch(cudaStream_t st, const __half* src, int O, int I, int KH, int KW, __half* dst,
                       int Ktot, int col0, int Ipad);
// 3x3 conv that follows a nearest-2x upsample -> four 2x2 phase kernels on the original image (sums of the 3x3 taps that
// read the same source pixel, added in f32, rounded once): dst [4 (a*2+b)][O][4 (th*2+tw) * Ipad].
int repack_upconv_launch(cudaStream_t st, const __half* src, int O, int I, __half* dst, int Ipad);
int vec_add_f32_launch(cudaStream_t st, float* dst, const float* src, int n);  // dst += src
// bias f16 [N] -> f32, optional GEGLU permutation, optional accumulate (dst += src).
int bias_to_f32_launch(cudaStream_t st, const __half* src, int N, float* dst, int geglu_bn, int accumulate);

}  // namespace sdxl
