// Implicit-GEMM on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA
// into 128B-swizzled shared memory). One kernel serves every dense contraction of the UNet step:
//   * Linear layers (reference unet/mod.rs:830,839,917,944,1009-1011,1021) = 1 segment, 1x1 tap;
//   * 3x3 / 1x1 convolutions (unet/mod.rs:1086,1096,1099,750,767-770,490): one K-segment per filter
//     tap; the tap shift is a TMA box offset on the NHWC activation, zero padding comes from TMA
//     out-of-bounds fill; the ResBlock skip 1x1 conv is just one more K-segment on a second tensor.
// Warp roles (192 threads): warp0 = TMA producer, warp1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias / residual / GEGLU -> global).
#include "common.cuh"
#include "kernels.h"

#include <stdio.h>

namespace sdxl {

static constexpr int kTileM = 128;
static constexpr int kBlockK = 64;                  // 64 halves = 128 B = one swizzle row
static constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KB

// Persistent, warp-specialised: grid = min(#tiles, #SMs); every role walks the same static tile sequence
// (tile = blockIdx.x + i*gridDim.x, M fastest so concurrently running CTAs share the weight tile in L2).
// The smem operand ring runs across tile boundaries, and the accumulator is double-buffered in TMEM
// (2 x BN columns), so the epilogue of tile i overlaps the MMA main loop of tile i+1.
//   warp 0      : TMA producer (one lane)
//   warp 1      : TMEM owner + MMA issuer (one lane)
//   warps 2..9  : epilogue, 2 warps per TMEM lane quarter (each takes half of the tile's columns)
static constexpr int kEpiWarps = 8;
static constexpr int kThreads = 64 + kEpiWarps * 32;

__global__ void __launch_bounds__(kThreads, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A 16KB | B BN*128)] [full][empty][tmem_full x2][tmem_empty x2][tmem ptr]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = p.BN;
  const int nst = p.nstages;
  const uint32_t stage_bytes = kABytes + BN * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)nst * stage_bytes);
  uint64_t* empty_bar = full_bar + nst;
  uint64_t* tmem_full = empty_bar + nst;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = p.tilesW * p.tilesH * p.tilesB;
  const int num_tiles = m_tiles * p.tilesN;

  int total_kb = 0;
  for (int s = 0; s < p.nseg; ++s) total_kb += p.seg[s].nkb;

  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * BN)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmA1);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < nst; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], kEpiWarps);
    mbar_init(&tmem_empty[1], kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // PDL: everything above overlapped the previous kernel's tail; from here on we touch its outputs.
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % m_tiles, nt = tile / m_tiles;
        const int tw = mt % p.tilesW;
        const int th = (mt / p.tilesW) % p.tilesH;
        const int tb = mt / (p.tilesW * p.tilesH);
        const int w0 = tw * p.Wt, h0 = th * p.Ht, b0 = tb * p.Bt;
        const int n0 = nt * BN;
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const IgemmSeg sg = p.seg[s];
          const void* mapA = sg.map ? (const void*)&p.tmA1 : (const void*)&p.tmA0;
          for (int j = 0; j < sg.nkb; ++j, ++it, ++kb) {
            const int stage = it % nst;
            const uint32_t par = (it / nst) & 1;
            mbar_wait(&empty_bar[stage], par ^ 1);
            uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
            uint8_t* b_dst = a_dst + kABytes;
            mbar_expect_tx(&full_bar[stage], stage_bytes);
            tma_load_4d(a_dst, mapA, &full_bar[stage], j * kBlockK, w0 + sg.dw, h0 + sg.dh, b0 + sg.db);
            tma_load_2d(b_dst, &p.tmB, &full_bar[stage], kb * kBlockK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16((uint32_t)BN, false);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1) ^ 1);  // epilogue drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int stage = it % nst;
          const uint32_t par = (it / nst) & 1;
          mbar_wait(&full_bar[stage], par);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t ad = make_sw128_desc(a_addr + k * 32);
            const uint64_t bd = make_sw128_desc(b_addr + k * 32);
            tc_mma_f16(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        }
        tc_commit(&tmem_full[buf]);  // accumulator of this tile complete
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;     // which half of the tile's column chunks
    const int r = q * 32 + lane;          // tile row == TMEM lane
    const int wt = r % p.Wt;
    const int ht = (r / p.Wt) % p.Ht;
    const int bt = r / (p.Wt * p.Ht);
    const int nchunks = (p.mode == IGEMM_LINEAR ? BN : (BN >> 1)) >> 4;
    const int c_begin = half == 0 ? 0 : ((nchunks + 1) >> 1);
    const int c_end = half == 0 ? ((nchunks + 1) >> 1) : nchunks;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int mt = tile % m_tiles, nt = tile / m_tiles;
      const int tw = mt % p.tilesW;
      const int th = (mt / p.tilesW) % p.tilesH;
      const int tb = mt / (p.tilesW * p.tilesH);
      const int bb = tb * p.Bt + bt, hh = th * p.Ht + ht, ww = tw * p.Wt + wt;
      const int n0 = nt * BN;
      const bool row_ok = (bb < p.Bn) && (hh < p.H) && (ww < p.W);
      const size_t pix = ((size_t)bb * p.H + hh) * p.W + ww;
      const int buf = lt & 1;
      mbar_wait(&tmem_full[buf], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);

      if (p.mode == IGEMM_LINEAR) {
        const float* bias = p.bias ? p.bias + (size_t)bb * p.bias_bstride : nullptr;
        const float* res = p.res ? p.res + pix * p.ldr : nullptr;
        for (int ch = c_begin; ch < c_end; ++ch) {
          const int c = ch << 4;
          uint32_t v[16];
          tmem_ld16(trow + c, v);
          tmem_ld_wait();
          const int n = n0 + c;
          if (row_ok && n < p.N) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            if (n + 16 <= p.N) {
              if (bias) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n + i));
                  f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
                }
              }
              if (res) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 r4 = *reinterpret_cast<const float4*>(res + n + i);
                  f[i] += r4.x; f[i + 1] += r4.y; f[i + 2] += r4.z; f[i + 3] += r4.w;
                }
              }
              if (p.out_f32) {
                float* o = reinterpret_cast<float*>(p.out) + pix * p.ldo + n;
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                  *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
              } else {
                __half* o = reinterpret_cast<__half*>(p.out) + pix * p.ldo + n;
                uint32_t h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                  h[i] = *reinterpret_cast<uint32_t*>(&t);
                }
                *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(o + 8) = make_uint4(h[4], h[5], h[6], h[7]);
              }
            } else {
              // ragged N tail (e.g. the 320->4 output conv): scalar, guarded
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (n + i < p.N) {
                  float x = f[i];
                  if (bias) x += bias[n + i];
                  if (res) x += res[n + i];
                  if (p.out_f32) reinterpret_cast<float*>(p.out)[pix * p.ldo + n + i] = x;
                  else reinterpret_cast<__half*>(p.out)[pix * p.ldo + n + i] = __float2half_rn(x);
                }
              }
            }
          }
        }
      } else {
        // GEGLU (reference unet/mod.rs:942-956): tile columns [0,BN/2) = value, [BN/2,BN) = matching gate
        const int hb = BN >> 1;
        __half* o = reinterpret_cast<__half*>(p.out) + pix * p.ldo + (size_t)nt * hb;
        for (int ch = c_begin; ch < c_end; ++ch) {
          const int c = ch << 4;
          uint32_t v[16], g[16];
          tmem_ld16(trow + c, v);
          tmem_ld16(trow + hb + c, g);
          tmem_ld_wait();
          if (row_ok && n0 + c < p.N) {
            uint32_t h[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float x0 = __uint_as_float(v[i]), x1 = __uint_as_float(v[i + 1]);
              float g0 = __uint_as_float(g[i]), g1 = __uint_as_float(g[i + 1]);
              if (p.bias) {
                x0 += __ldg(p.bias + n0 + c + i);
                x1 += __ldg(p.bias + n0 + c + i + 1);
                g0 += __ldg(p.bias + n0 + hb + c + i);
                g1 += __ldg(p.bias + n0 + hb + c + i + 1);
              }
              __half2 t = __floats2half2_rn(x0 * gelu_erf_f(g0), x1 * gelu_erf_f(g1));
              h[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(o + c) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(o + c + 8) = make_uint4(h[4], h[5], h[6], h[7]);
          }
        }
      }
      // all TMEM reads of this buffer are complete (tcgen05.wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(f);
  }
  return fn;
}

static int encode(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  PFN_encodeTiled fn = get_encode();
  if (!fn) return 1001;
  cuuint64_t gd[5];
  cuuint64_t gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "sdxl_b200: cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]\n",
            (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
            (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
            rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return 1002;
  }
  return 0;
}

int make_tmap_act(CUtensorMap* tm, const __half* base, int Bn, int H, int W, int C, int pitch, int Wt, int Ht,
                  int Bt) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)Bn};
  uint64_t str[3] = {(uint64_t)pitch * 2, (uint64_t)W * pitch * 2, (uint64_t)H * W * pitch * 2};
  uint32_t box[4] = {64, (uint32_t)Wt, (uint32_t)Ht, (uint32_t)Bt};
  return encode(tm, base, 4, dims, str, box);
}
int make_tmap_wgt(CUtensorMap* tm, const __half* base, int N, int K, int BN) {
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
  uint64_t str[1] = {(uint64_t)K * 2};
  uint32_t box[2] = {64, (uint32_t)BN};
  return encode(tm, base, 2, dims, str, box);
}
int make_tmap_rows(CUtensorMap* tm, const __half* base, int rows_per_batch, int nbatch, int cols, int pitch) {
  uint64_t dims[3] = {(uint64_t)cols, (uint64_t)rows_per_batch, (uint64_t)nbatch};
  uint64_t str[2] = {(uint64_t)pitch * 2, (uint64_t)rows_per_batch * pitch * 2};
  uint32_t box[3] = {64, 128, 1};
  return encode(tm, base, 3, dims, str, box);
}

void igemm_pick_box(int W, int H, int* Wt, int* Ht, int* Bt) {
  int wt = 1;
  while (wt * 2 <= W && wt * 2 <= 128) wt *= 2;
  int ht = 1;
  while (ht * 2 <= H && wt * ht * 2 <= 128) ht *= 2;
  *Wt = wt;
  *Ht = ht;
  *Bt = 128 / (wt * ht);
}

int igemm_pick_bn(int m_tiles, int N, int num_sms, bool geglu) {
  // candidate tiles: multiples of 16 (32 for GEGLU so both halves are x16-aligned) that divide N exactly
  // (or cover N when N is tiny). Cost model: waves * BN (tensor time ~ BN per tile) with a mild penalty
  // for narrow tiles (operand re-reads from smem/L2).
  if (N <= 16) return 16;
  double best = 1e30;
  int best_bn = 0;
  const int step = geglu ? 32 : 16;
  for (int bn = 256; bn >= 32; bn -= step) {
    if (N % bn) continue;
    const long tiles = (long)m_tiles * (N / bn);
    const long waves = (tiles + num_sms - 1) / num_sms;
    double cost = (double)waves * (bn + 48.0);  // +48: fixed A-operand/epilogue cost per tile
    if (cost < best) {
      best = cost;
      best_bn = bn;
    }
  }
  if (!best_bn) {  // N has no suitable divisor: cover with padding
    best_bn = N >= 256 ? 256 : ((N + 15) / 16) * 16;
  }
  return best_bn;
}

int igemm_launch(cudaStream_t st, IgemmParams& p) {
  p.tilesW = (p.W + p.Wt - 1) / p.Wt;
  p.tilesH = (p.H + p.Ht - 1) / p.Ht;
  p.tilesB = (p.Bn + p.Bt - 1) / p.Bt;
  const int stage_bytes = kABytes + p.BN * 128;
  int nst = (224 * 1024) / stage_bytes;
  if (nst > 8) nst = 8;
  if (nst < 2) nst = 2;
  p.nstages = nst;
  const size_t smem = (size_t)nst * stage_bytes + 1024 /*align slack*/ + (2 * nst + 4) * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  p.tilesN = (p.mode == IGEMM_GEGLU) ? (p.N / p.BN) : ((p.N + p.BN - 1) / p.BN);
  const int tiles = p.tilesW * p.tilesH * p.tilesB * p.tilesN;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_kernel(igemm_kernel, dim3(grid), dim3(kThreads), smem, st, true, p);
}

}  // namespace sdxl
