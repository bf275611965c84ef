// Small HBM-bound kernels of the UNet step: embedding GEMVs, first conv, resampling copies, sampler
// elementwise math (CFG + DDIM, inpainting blend), Philox noise, and load-time weight re-layout.
// All are plain coalesced / 128-bit vectorised CUDA; none of them is GEMM-shaped.
#include "common.cuh"
#include "kernels.h"

#include <stdlib.h>

namespace sdxl {

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

bool pdl_enabled() {
  static const bool on = getenv("SDXL_B200_NO_PDL") == nullptr;
  return on;
}

// ------------------------------------------------------------------------------------------------
// GEMV: one warp per output column, up to 8 batch rows accumulated together.
// ------------------------------------------------------------------------------------------------
template <int MAXB>
__global__ void gemv_kernel(const float* __restrict__ in, int in_bstride, int Bv, int K,
                            const __half* __restrict__ W, int ldw, const float* __restrict__ bias,
                            const float* __restrict__ add, int add_bstride, int N, int in_silu, int out_silu,
                            float* __restrict__ out, int out_bstride) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float acc[MAXB];
#pragma unroll
  for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
  const __half* w = W + (size_t)n * ldw;
  if ((K & 7) == 0 && (ldw & 7) == 0 && (in_bstride & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    // four 16-byte weight loads in flight per lane before the first FMA (a K = 1280 row is 5 loads per lane: one DRAM round
    // trip each if issued one per iteration), activations as float4 (L1 hits)
    for (int k0 = lane * 8; k0 < K; k0 += 1024) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        raw[u] = (k0 + u * 256 < K) ? __ldg(reinterpret_cast<const uint4*>(w + k0 + u * 256)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = k0 + u * 256;
        if (kk < K) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&raw[u]);
          float wf[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h2[i]);
            wf[2 * i] = f.x;
            wf[2 * i + 1] = f.y;
          }
#pragma unroll
          for (int b = 0; b < MAXB; ++b) {
            if (b < Bv) {
              const float4* x4 = reinterpret_cast<const float4*>(in + (size_t)b * in_bstride + kk);
              const float4 xa = __ldg(x4), xb = __ldg(x4 + 1);
              float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float v = xv[i];
                if (in_silu) v = silu_f(v);
                acc[b] = fmaf(v, wf[i], acc[b]);
              }
            }
          }
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float wv = __half2float(w[k]);
#pragma unroll
      for (int b = 0; b < MAXB; ++b) {
        if (b < Bv) {
          float v = in[(size_t)b * in_bstride + k];
          if (in_silu) v = silu_f(v);
          acc[b] = fmaf(v, wv, acc[b]);
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < MAXB; ++b) {
    if (b < Bv) {
      float s = warp_sum(acc[b]);
      if (lane == 0) {
        if (bias) s += bias[n];
        if (add) s += add[(size_t)b * add_bstride + n];
        if (out_silu) s = silu_f(s);
        out[(size_t)b * out_bstride + n] = s;
      }
    }
  }
}
int gemv_launch(cudaStream_t st, const float* in, int in_bstride, int Bv, int K, const __half* W, int ldw, const float* bias,
                const float* add, int add_bstride, int N, int in_silu, int out_silu, float* out, int out_bstride) {
  if (Bv > 8) return 2001;
  const int warps = 8;
  if (Bv <= 2)
    gemv_kernel<2><<<cdiv(N, warps), warps * 32, 0, st>>>(in, in_bstride, Bv, K, W, ldw, bias, add, add_bstride, N, in_silu,
                                                           out_silu, out, out_bstride);
  else
    gemv_kernel<8><<<cdiv(N, warps), warps * 32, 0, st>>>(in, in_bstride, Bv, K, W, ldw, bias, add, add_bstride, N, in_silu,
                                                           out_silu, out, out_bstride);
  return (int)cudaGetLastError();
}

// timestep_embedding (reference unet/mod.rs:21-39): cos half first, then sin half.
__global__ void timestep_embedding_kernel(const int* __restrict__ t, int nt, int dim, float max_period,
                                          float* __restrict__ out) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt * half) return;
  const int b = i / half, j = i % half;
  const float freq = expf((float)j * (-logf(max_period) / (float)half));
  const float arg = (float)t[b] * freq;
  out[(size_t)b * dim + j] = cosf(arg);
  out[(size_t)b * dim + half + j] = sinf(arg);
}
int timestep_embedding_launch(cudaStream_t st, const int* t_dev, int nt, int dim, float max_period, float* out) {
  const int n = nt * (dim / 2);
  timestep_embedding_kernel<<<cdiv(n, 128), 128, 0, st>>>(t_dev, nt, dim, max_period, out);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// First conv (4 -> model_channels, 3x3 pad 1; reference unet/mod.rs:116-120). K = 36: CUDA cores.
// x: NCHW (f16 or f32) [Bx, Cin, H, W]; output batch b reads image (b % Bx). y: NHWC f32.
// ------------------------------------------------------------------------------------------------
// One thread = 4 output channels x 8 consecutive pixels of a row: every weight float4 read from shared memory feeds 32 FMAs
// (at one pixel per thread the kernel was bound by the LDS.128 per 4 FMAs: 158 us for the UNet's 2x128x128x320 output).
constexpr int kConvInPix = 8;
template <typename TIn>
__global__ void __launch_bounds__(256) conv_in_kernel(const TIn* __restrict__ x, int Bx, int B, int Cin, int H, int W,
                                                      const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                                                      float* __restrict__ y) {
  extern __shared__ float sw[];  // [9*Cin][Cout]  (k-major: lanes = consecutive output channels, conflict-free)
  const int kk = 9 * Cin;
  for (int co = threadIdx.x; co < Cout; co += blockDim.x)          // lanes = consecutive rows of w: conflict-free smem writes,
    for (int k = 0; k < kk; ++k) sw[k * Cout + co] = w[co * kk + k];   // the rows' lines stay in L1 across the k loop
  __syncthreads();
  const int cvec = Cout / 4;
  const int nseg = (W + kConvInPix - 1) / kConvInPix;
  const long total = (long)B * H * nseg * cvec;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvec);
    const long seg = idx / cvec;
    const int w0 = (int)(seg % nseg) * kConvInPix;
    const int hh = (int)((seg / nseg) % H);
    const int b = (int)(seg / ((long)nseg * H));
    const TIn* xb = x + (size_t)(b % Bx) * Cin * H * W;
    const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + cv * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc[kConvInPix];
#pragma unroll
    for (int i = 0; i < kConvInPix; ++i) acc[i] = b4;
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = hh + kh - 1;
      if (ih < 0 || ih >= H) continue;
      for (int c = 0; c < Cin; ++c) {
        const TIn* xr = xb + ((size_t)c * H + ih) * W;
        float xv[kConvInPix + 2];                                  // the row segment with its halo; same address in all lanes
#pragma unroll
        for (int i = 0; i < kConvInPix + 2; ++i) {
          const int iw = w0 + i - 1;
          xv[i] = (iw >= 0 && iw < W) ? (float)xr[iw] : 0.f;
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 wv = *reinterpret_cast<const float4*>(sw + ((kh * 3 + kw) * Cin + c) * Cout + cv * 4);
#pragma unroll
          for (int i = 0; i < kConvInPix; ++i) {
            const float v = xv[i + kw];
            acc[i].x = fmaf(v, wv.x, acc[i].x); acc[i].y = fmaf(v, wv.y, acc[i].y);
            acc[i].z = fmaf(v, wv.z, acc[i].z); acc[i].w = fmaf(v, wv.w, acc[i].w);
          }
        }
      }
    }
    float* yo = y + (((size_t)b * H + hh) * W + w0) * Cout + cv * 4;
#pragma unroll
    for (int i = 0; i < kConvInPix; ++i)
      if (w0 + i < W) *reinterpret_cast<float4*>(yo + (size_t)i * Cout) = acc[i];
  }
}
int conv_in_launch_t(cudaStream_t st, const void* x, int x_f32, int Bx, int B, int Cin, int H, int W, const float* w,
                     const float* bias, int Cout, float* y) {
  if (Cin > 8 || (Cout & 3)) return 2002;
  const size_t smem = (size_t)Cout * 9 * Cin * sizeof(float);
  static bool done_f[64], done_h[64];
  if (smem > 200 * 1024) return 2002;
  if (int r = smem_optin(conv_in_kernel<float>, 200 * 1024, done_f)) return r;
  if (int r = smem_optin(conv_in_kernel<__half>, 200 * 1024, done_h)) return r;
  const long total = (long)B * H * ((W + kConvInPix - 1) / kConvInPix) * (Cout / 4);
  int grid = cdiv(total, 256);
  if (grid > 148 * 4) grid = 148 * 4;
  if (x_f32)
    conv_in_kernel<float><<<grid, 256, smem, st>>>((const float*)x, Bx, B, Cin, H, W, w, bias, Cout, y);
  else
    conv_in_kernel<__half><<<grid, 256, smem, st>>>((const __half*)x, Bx, B, Cin, H, W, w, bias, Cout, y);
  return (int)cudaGetLastError();
}
int conv_in_launch(cudaStream_t st, const __half* x, int B, int Cin, int H, int W, const float* w, const float* bias,
                   int Cout, float* y) {
  return conv_in_launch_t(st, x, 0, B, B, Cin, H, W, w, bias, Cout, y);
}

// ------------------------------------------------------------------------------------------------
// resampling copies (NHWC, 4 channels per thread)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 pack4h(float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  return r;
}
__global__ void upsample2x_kernel(const float* __restrict__ x, int B, int H, int W, int C, __half* __restrict__ y) {
  const int cv = C / 4;
  const long total = (long)B * (2 * H) * (2 * W) * cv;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv);
    const long pix = idx / cv;
    const int ow = (int)(pix % (2 * W));
    const int oh = (int)((pix / (2 * W)) % (2 * H));
    const int b = (int)(pix / ((long)4 * W * H));
    const float4 v = *reinterpret_cast<const float4*>(x + (((size_t)b * H + (oh >> 1)) * W + (ow >> 1)) * C + c * 4);
    *reinterpret_cast<uint2*>(y + pix * C + c * 4) = pack4h(v);
  }
}
int upsample2x_launch(cudaStream_t st, const float* x, int B, int H, int W, int C, __half* y) {
  if (C & 3) return 2003;
  const long total = (long)B * 4 * H * W * (C / 4);
  int grid = cdiv(total, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  upsample2x_kernel<<<grid, 256, 0, st>>>(x, B, H, W, C, y);
  return (int)cudaGetLastError();
}
__global__ void phase_split_kernel(const float* __restrict__ x, int B, int H, int W, int C, __half* __restrict__ y) {
  const int cv = C / 4;
  const int H2 = H / 2, W2 = W / 2;
  const long total = (long)B * H * W * cv;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv);
    const long pix = idx / cv;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int b = (int)(pix / ((long)W * H));
    const float4 v = *reinterpret_cast<const float4*>(x + pix * C + c * 4);
    const int ph = (h & 1) * 2 + (w & 1);
    const size_t o = ((((size_t)ph * B + b) * H2 + (h >> 1)) * W2 + (w >> 1)) * C + c * 4;
    *reinterpret_cast<uint2*>(y + o) = pack4h(v);
  }
}
int phase_split_launch(cudaStream_t st, const float* x, int B, int H, int W, int C, __half* y) {
  if ((C & 3) || (H & 1) || (W & 1)) return 2004;
  const long total = (long)B * H * W * (C / 4);
  int grid = cdiv(total, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  phase_split_kernel<<<grid, 256, 0, st>>>(x, B, H, W, C, y);
  return (int)cudaGetLastError();
}

__global__ void cast_f32_f16_kernel(const float* __restrict__ x, size_t n, __half* __restrict__ y) {
  griddep_wait();
  griddep_launch_dependents();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(x[i]);
}
__global__ void cast_f16_f32_kernel(const __half* __restrict__ x, size_t n, float* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __half2float(x[i]);
}
int cast_f32_to_f16_launch(cudaStream_t st, const float* x, size_t n, __half* y) {
  int grid = cdiv((long)n, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  return launch_kernel(cast_f32_f16_kernel, dim3(grid), dim3(256), (size_t)0, st, true, x, n, y);
}
int cast_f16_to_f32_launch(cudaStream_t st, const __half* x, size_t n, float* y) {
  int grid = cdiv((long)n, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  cast_f16_f32_kernel<<<grid, 256, 0, st>>>(x, n, y);
  return (int)cudaGetLastError();
}
__global__ void nhwc_to_nchw_f16_kernel(const float* __restrict__ x, int B, int HW, int C, int ldx,
                                        __half* __restrict__ y) {
  const long total = (long)B * C * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int b = (int)(i / ((long)HW * C));
    y[i] = __float2half_rn(x[((size_t)b * HW + p) * ldx + c]);
  }
}
__global__ void nhwc_to_nchw_f32_kernel(const float* __restrict__ x, int B, int HW, int C, int ldx,
                                        float* __restrict__ y) {
  const long total = (long)B * C * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int b = (int)(i / ((long)HW * C));
    y[i] = x[((size_t)b * HW + p) * ldx + c];
  }
}
int nhwc_to_nchw_f32_launch(cudaStream_t st, const float* x, int B, int HW, int C, int ldx, float* y) {
  const long total = (long)B * C * HW;
  nhwc_to_nchw_f32_kernel<<<cdiv(total, 256), 256, 0, st>>>(x, B, HW, C, ldx, y);
  return (int)cudaGetLastError();
}
int nhwc_to_nchw_f16_launch(cudaStream_t st, const float* x, int B, int HW, int C, int ldx, __half* y) {
  const long total = (long)B * C * HW;
  nhwc_to_nchw_f16_kernel<<<cdiv(total, 256), 256, 0, st>>>(x, B, HW, C, ldx, y);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// sampler elementwise math. All latents are f32 NCHW [Bimg, C, HW].
// ------------------------------------------------------------------------------------------------
__global__ void cfg_ddim_kernel(const float* __restrict__ eps, int ld, int Bimg, int C, int HW, int use_cfg,
                                float g, float sqrt_a, float sqrt_1ma, float sqrt_ap, float sqrt_1map,
                                float* __restrict__ x) {
  const long total = (long)Bimg * C * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int b = (int)(i / ((long)HW * C));
    const float ec = eps[((size_t)b * HW + p) * ld + c];
    float e = ec;
    if (use_cfg) {
      // u + (c - u) * s   (reference stablediffusion/mod.rs:539-540)
      const float eu = eps[((size_t)(Bimg + b) * HW + p) * ld + c];
      e = eu + (ec - eu) * g;
    }
    // DDIM eta=0 (reference stablediffusion/mod.rs:423-428)
    const float xv = x[i];
    const float predx0 = (xv - e * sqrt_1ma) / sqrt_a;
    x[i] = predx0 * sqrt_ap + e * sqrt_1map;
  }
}
int cfg_ddim_launch(cudaStream_t st, const float* eps, int ld, int Bimg, int C, int HW, int use_cfg, float guidance,
                    float sqrt_a, float sqrt_1ma, float sqrt_ap, float sqrt_1map, float* x, __half* /*x16*/) {
  const long total = (long)Bimg * C * HW;
  cfg_ddim_kernel<<<cdiv(total, 256), 256, 0, st>>>(eps, ld, Bimg, C, HW, use_cfg, guidance, sqrt_a, sqrt_1ma,
                                                     sqrt_ap, sqrt_1map, x);
  return (int)cudaGetLastError();
}
// x = mask ? x : (ref*sqrt_a + noise*sqrt_1ma)   (reference stablediffusion/mod.rs:463-465)
__global__ void inpaint_blend_kernel(float* __restrict__ x, const float* __restrict__ ref,
                                     const float* __restrict__ noise, const uint8_t* __restrict__ mask, size_t n,
                                     float sqrt_a, float sqrt_1ma) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float nr = ref[i] * sqrt_a + noise[i] * sqrt_1ma;
    x[i] = mask[i] ? x[i] : nr;
  }
}
int inpaint_blend_launch(cudaStream_t st, float* x, const float* ref, const float* noise, const uint8_t* mask,
                         size_t n, int /*nfwd*/, float sqrt_a, float sqrt_1ma, __half* /*x16*/) {
  inpaint_blend_kernel<<<cdiv((long)n, 256), 256, 0, st>>>(x, ref, noise, mask, n, sqrt_a, sqrt_1ma);
  return (int)cudaGetLastError();
}
__global__ void axpby_kernel(float* __restrict__ x, const float* __restrict__ noise, size_t n, float sa, float sb) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = x[i] * sa + noise[i] * sb;
}
int axpby_launch(cudaStream_t st, float* x, const float* noise, size_t n, float sa, float sb) {
  axpby_kernel<<<cdiv((long)n, 256), 256, 0, st>>>(x, noise, n, sa, sb);
  return (int)cudaGetLastError();
}
int dup_latent_f16_launch(cudaStream_t st, const float* x, size_t n, int nfwd, __half* x16) {
  for (int f = 0; f < nfwd; ++f) {
    int e = cast_f32_to_f16_launch(st, x, n, x16 + (size_t)f * n);
    if (e) return e;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG + Box-Muller. Element i of stream (seed, subseq): counter =
// (i/4 lo32, i/4 hi32, subseq lo32, subseq hi32), key = (seed lo32, seed hi32); lane i%4 of the
// 4 normals produced from the 4 output words. oracle/philox.py is the bit-identical restatement.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__global__ void randn_kernel(float* __restrict__ out, size_t n, uint64_t seed, uint64_t subseq) {
  const size_t nblk = (n + 3) / 4;
  for (size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x; blk < nblk;
       blk += (size_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)subseq, (uint32_t)(subseq >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float z[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float u1 = ((float)(c[2 * j] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float u2 = ((float)(c[2 * j + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float rad = sqrtf(-2.0f * logf(u1));
      const float ang = 6.283185307179586f * u2;
      z[2 * j] = rad * cosf(ang);
      z[2 * j + 1] = rad * sinf(ang);
    }
    for (int j = 0; j < 4; ++j)
      if (blk * 4 + j < n) out[blk * 4 + j] = z[j];
  }
}
int randn_launch(cudaStream_t st, float* out, size_t n, uint64_t seed, uint64_t subseq) {
  randn_kernel<<<cdiv((long)((n + 3) / 4), 256), 256, 0, st>>>(out, n, seed, subseq);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// load-time weight re-layout
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int geglu_perm(int n, int N, int bn) {
  if (bn <= 0) return n;
  const int half_n = N / 2, hb = bn / 2;
  const int g = n >= half_n;
  const int m = g ? n - half_n : n;
  return (m / hb) * bn + g * hb + (m % hb);
}
// src [K][N] row-major -> dst [N][Kpad] (tiled transpose through smem)
__global__ void transpose_linear_kernel(const __half* __restrict__ src, int K, int N, __half* __restrict__ dst,
                                        int Kpad, int dst_row0, int geglu_bn) {
  __shared__ __half tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? src[(size_t)k * N + n] : __float2half(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < Kpad) dst[(size_t)(dst_row0 + geglu_perm(n, N, geglu_bn)) * Kpad + k] = tile[threadIdx.x][i];
  }
}
int transpose_linear_launch(cudaStream_t st, const __half* src, int K, int N, __half* dst, int Kpad, int dst_row0,
                            int geglu_bn) {
  dim3 grid(cdiv(N, 32), cdiv(Kpad, 32));
  transpose_linear_kernel<<<grid, dim3(32, 8), 0, st>>>(src, K, N, dst, Kpad, dst_row0, geglu_bn);
  return (int)cudaGetLastError();
}
// OIHW -> dst[o][col0 + (kh*KW+kw)*Ipad + i]
__global__ void repack_conv_kernel(const __half* __restrict__ src, int O, int I, int KH, int KW,
                                   __half* __restrict__ dst, int Ktot, int col0, int Ipad) {
  const long total = (long)O * KH * KW * Ipad;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % Ipad);
    const int tap = (int)((idx / Ipad) % (KH * KW));
    const int o = (int)(idx / ((long)Ipad * KH * KW));
    const __half v = i < I ? src[(((size_t)o * I + i) * KH + tap / KW) * KW + tap % KW] : __float2half(0.f);
    dst[(size_t)o * Ktot + col0 + (size_t)tap * Ipad + i] = v;
  }
}
int repack_conv_launch(cudaStream_t st, const __half* src, int O, int I, int KH, int KW, __half* dst, int Ktot,
                       int col0, int Ipad) {
  const long total = (long)O * KH * KW * Ipad;
  int grid = cdiv(total, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  repack_conv_kernel<<<grid, 256, 0, st>>>(src, O, I, KH, KW, dst, Ktot, col0, Ipad);
  return (int)cudaGetLastError();
}
// Nearest-2x upsample followed by a 3x3 pad-1 conv (reference unet/mod.rs:742-751, autoencoder/mod.rs:311-319): output pixel
// (2i+a, 2j+b) reads upsampled rows 2i+a-1 .. 2i+a+1 = source rows {i-1, i, i} for a = 0 and {i, i, i+1} for a = 1 (same for
// columns), so each output parity (a, b) is a 2x2 convolution of the source image with summed taps:
//   a = 0: tap th=0 (dh=-1) = kh{0},   th=1 (dh=0)  = kh{1,2};     a = 1: th=0 (dh=0) = kh{0,1},   th=1 (dh=+1) = kh{2}
__global__ void repack_upconv_kernel(const __half* __restrict__ src, int O, int I, __half* __restrict__ dst, int Ipad) {
  const long total = (long)4 * O * 4 * Ipad;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % Ipad);
    const int tap = (int)((idx / Ipad) % 4);
    const int o = (int)((idx / ((long)Ipad * 4)) % O);
    const int ph = (int)(idx / ((long)Ipad * 4 * O));
    const int a = ph >> 1, b = ph & 1, th = tap >> 1, tw = tap & 1;
    float acc = 0.f;
    if (i < I) {
      const int kh0 = a == 0 ? (th == 0 ? 0 : 1) : (th == 0 ? 0 : 2), kh1 = a == 0 ? (th == 0 ? 0 : 2) : (th == 0 ? 1 : 2);
      const int kw0 = b == 0 ? (tw == 0 ? 0 : 1) : (tw == 0 ? 0 : 2), kw1 = b == 0 ? (tw == 0 ? 0 : 2) : (tw == 0 ? 1 : 2);
      for (int kh = kh0; kh <= kh1; ++kh)
        for (int kw = kw0; kw <= kw1; ++kw) acc += __half2float(src[(((size_t)o * I + i) * 3 + kh) * 3 + kw]);
    }
    dst[idx] = __float2half_rn(acc);
  }
}
int repack_upconv_launch(cudaStream_t st, const __half* src, int O, int I, __half* dst, int Ipad) {
  const long total = (long)4 * O * 4 * Ipad;
  int grid = cdiv(total, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  repack_upconv_kernel<<<grid, 256, 0, st>>>(src, O, I, dst, Ipad);
  return (int)cudaGetLastError();
}
__global__ void bias_to_f32_kernel(const __half* __restrict__ src, int N, float* __restrict__ dst, int geglu_bn,
                                   int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int d = geglu_perm(n, N, geglu_bn);
  const float v = __half2float(src[n]);
  dst[d] = accumulate ? dst[d] + v : v;
}
__global__ void vec_add_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
int vec_add_f32_launch(cudaStream_t st, float* dst, const float* src, int n) {
  vec_add_f32_kernel<<<cdiv(n, 256), 256, 0, st>>>(dst, src, n);
  return (int)cudaGetLastError();
}
int bias_to_f32_launch(cudaStream_t st, const __half* src, int N, float* dst, int geglu_bn, int accumulate) {
  bias_to_f32_kernel<<<cdiv(N, 256), 256, 0, st>>>(src, N, dst, geglu_bn, accumulate);
  return (int)cudaGetLastError();
}

}  // namespace sdxl
