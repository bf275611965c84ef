// Fused multi-head attention for the UNet's SpatialTransformer blocks (head dim 64, no mask):
//   out = softmax(q k^T / sqrt(64)) v            (reference src/backend.rs:4-19,32-79,88-128;
//                                                 called from unet/mod.rs:1013-1019)
// Flash-style: one CTA owns 128 queries of one head and streams 128-key blocks. Both contractions run
// on tcgen05 tensor cores with accumulators in TMEM:
//   S = Q K^T   : A = Q tile (K-major, TMA SW128), B = K tile (K-major)           -> TMEM cols [0,128)
//   O_j = P V   : A = P (f16, written by the softmax warps into SW128 smem),
//                 B = V tile (MN-major: keys are the contraction dim)              -> TMEM cols [128,192)
// Online softmax (running max / sum, f32) lives in registers of 128 threads (one query row each);
// the running output is rescaled in registers, so TMEM never needs a read-modify-write.
// Warp roles (192 threads): warp0 TMA producer, warp1 MMA issuer + TMEM owner, warps 2..5 softmax.
#include "common.cuh"
#include "kernels.h"

namespace sdxl {

static constexpr int kQBytes = 128 * 128;       // 128 rows x 64 halves
static constexpr int kKVBytes = 128 * 128;
static constexpr int kPBytes = 2 * 128 * 128;   // 128 rows x 128 keys, two 64-key swizzle panels
static constexpr int kAttnSmem = kQBytes + 4 * kKVBytes + kPBytes + 128;

__global__ void __launch_bounds__(192, 2) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;            // 2 stages
  uint8_t* sV = sK + 2 * kKVBytes;       // 2 stages
  uint8_t* sP = sV + 2 * kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* pv_done = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.S + 127) / 128;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("sdxl_b200: attention smem base not 1024B aligned\n");
      __trap();
    }
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kQBytes);
      tma_load_3d(sQ, &p.tmQ, q_full, p.q_col0 + head * 64, qt * 128, b);
      for (int j = 0; j < nblk; ++j) {
        const int stage = j & 1;
        mbar_wait(&kv_empty[stage], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[stage], 2 * kKVBytes);
        tma_load_3d(sK + stage * kKVBytes, &p.tmK, &kv_full[stage], p.k_col0 + head * 64, j * 128, b);
        tma_load_3d(sV + stage * kKVBytes, &p.tmV, &kv_full[stage], p.v_col0 + head * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc_f16(128, false);
      const uint32_t idesc_pv = make_idesc_f16(64, true);
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      {
        const uint32_t k_addr = smem_u32(sK);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tmem_S, make_sw128_desc(q_addr + k * 32), make_sw128_desc(k_addr + k * 32), idesc_qk, k > 0);
        tc_commit(s_full);
      }
      for (int j = 0; j < nblk; ++j) {
        const int stage = j & 1;
        // P(j) written and S(j) fully consumed by the softmax warps
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + stage * kKVBytes);
#pragma unroll
        for (int t = 0; t < 8; ++t)
          tc_mma_f16(tmem_O, make_sw128_desc(p_addr + (t >> 2) * (128 * 128) + (t & 3) * 32),
                     make_sw128_desc(v_addr + t * 2048), idesc_pv, t > 0);
        tc_commit(pv_done);
        tc_commit(&kv_empty[stage]);
        if (j + 1 < nblk) {
          const int ns = (j + 1) & 1;
          mbar_wait(&kv_full[ns], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(sK + ns * kKVBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_S, make_sw128_desc(q_addr + k * 32), make_sw128_desc(k_addr + k * 32), idesc_qk, k > 0);
          tc_commit(s_full);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row in tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float sl2e = p.scale_log2e;
    float m = -INFINITY, l = 0.f;
    float O[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) O[i] = 0.f;
    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
    const int rx = r & 7;

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * 128;
      // pass 1: row max over valid keys
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (kbase + c + i < p.S) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m, mx);
      const float alpha = exp2f((m - m_new) * sl2e);
      const float mb = m_new * sl2e;
      float sum = 0.f;
      // pass 2: p = exp2(s*scale - m*scale) -> f16 -> swizzled smem (A operand of the PV MMA)
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c, v);
        tmem_ld_wait();
        uint32_t h[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (kbase + c + i < p.S) ? exp2f(fmaf(__uint_as_float(v[i]), sl2e, -mb)) : 0.f;
          float p1 = (kbase + c + i + 1 < p.S) ? exp2f(fmaf(__uint_as_float(v[i + 1]), sl2e, -mb)) : 0.f;
          __half2 t = __floats2half2_rn(p0, p1);
          const float2 back = __half22float2(t);  // sum exactly what the tensor core will see
          sum += back.x + back.y;
          h[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
        }
        uint8_t* panel = prow + (c >> 6) * (128 * 128);
        const int ch0 = (c & 63) >> 3;  // first 16B chunk of this 32-key group inside the 128B row
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(panel + (((ch0 + u) ^ rx) << 4)) =
              make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
      }
      l = l * alpha + sum;
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      // accumulate this block's P V
      mbar_wait(pv_done, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_O + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) O[c + i] = fmaf(O[c + i], alpha, __uint_as_float(v[i]));
      }
    }
    const int t = qt * 128 + r;
    if (t < p.T) {
      const float inv = 1.0f / l;
      __half* o = p.out + ((size_t)b * p.T + t) * p.ldo + head * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 8) {
        uint32_t h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __half2 t2 = __floats2half2_rn(O[c + 2 * i] * inv, O[c + 2 * i + 1] * inv);
          h[i] = *reinterpret_cast<uint32_t*>(&t2);
        }
        *reinterpret_cast<uint4*>(o + c) = make_uint4(h[0], h[1], h[2], h[3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int attention_launch(cudaStream_t st, const AttnParams& p) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  dim3 grid((p.T + 127) / 128, p.n_head, p.B);
  attention_kernel<<<grid, 192, kAttnSmem, st>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace sdxl
