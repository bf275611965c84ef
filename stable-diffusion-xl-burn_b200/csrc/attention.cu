// Fused multi-head attention for the UNet's SpatialTransformer blocks (head dim 64, no mask):
//   out = softmax(q k^T / sqrt(64)) v            (reference src/backend.rs:4-19,32-79,88-128;
//                                                 called from unet/mod.rs:1013-1019)
// Flash-style on tcgen05 tensor cores with TMEM accumulators. One CTA owns TWO 128-query tiles (A, B) of one
// head and streams 128-key blocks once for both; two softmax warpgroups ping-pong against one MMA issuer, so
// the tensor pipe works on tile B while tile A is in the exp/rescale phase and vice versa:
//   S_x = Q_x K^T : A = Q tile (K-major, TMA SW128), B = K tile (K-major)            -> TMEM S_A / S_B (128 cols each)
//   O_x,j = P_x V : A = P_x (f16, written by softmax group x into SW128 smem),
//                   B = V tile (MN-major: keys are the contraction dim)               -> TMEM O_x[j&1] (64 cols each)
// Online softmax (running max / sum, f32) lives in registers of 128 threads per tile (one query row each). The
// per-block P V result is double-buffered in TMEM and folded into the register accumulator one block late, so
// the softmax warps never wait for the tensor pipe in steady state and TMEM never needs a read-modify-write.
// Warp roles (320 threads): warp0 TMA producer, warp1 MMA issuer + TMEM owner, warps 2..5 softmax A, 6..9 softmax B.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace sdxl {

static constexpr int kTileBytes = 128 * 128;        // 128 rows x 64 halves (Q, K or V tile)
static constexpr int kPBytes = 2 * 128 * 128;       // 128 rows x 128 keys, two 64-key swizzle panels
static constexpr int kXchgBytes = 2 * 2 * 2 * 128 * 4;  // SPLIT=2: [slot][group][column half][row] f32
static constexpr int kAttnSmem = 2 * kTileBytes + 4 * kTileBytes + 2 * kPBytes + 256 + kXchgBytes;
// SPLIT = threads per query row in the softmax groups: 1 -> warps 2..5 / 6..9 (320 threads),
// 2 -> warps 2..9 / 10..17 (576 threads): each thread owns 64 of a block's 128 score columns and 32 of the 64 output columns,
// twice as many warps per scheduler to cover the tcgen05.ld latency.
template <int SPLIT> struct AttnCfg { static constexpr int kThreads = 64 + 256 * SPLIT; };

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int SPLIT>
__global__ void __launch_bounds__(AttnCfg<SPLIT>::kThreads, 1) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                       // [2] tiles A, B
  uint8_t* sK = sQ + 2 * kTileBytes;        // [2] stages
  uint8_t* sV = sK + 2 * kTileBytes;        // [2] stages
  uint8_t* sP = sV + 2 * kTileBytes;        // [2] groups
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;     // [2] per group
  uint64_t* p_full = bars + 7;     // [2] per group, 128 arrivals
  uint64_t* pv_done = bars + 9;    // [2] per group
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);
  float* xchg = reinterpret_cast<float*>(sP + 2 * kPBytes + 256);   // SPLIT=2 pair exchange slots

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = (p.S + 127) / 128;
  // Persistent: the (batch, head, 128-query tile) work list is split into contiguous, balanced ranges, one per
  // CTA; inside its range a CTA takes two consecutive tiles of the same head as a ping-pong pair, a single
  // tile otherwise (end of a head or of the range). This removes most of the wave-quantisation loss of a
  // one-pair-per-CTA grid (e.g. 320 tiles on 148 SMs: 1 pair + 1 single instead of 2 full waves).
  const int nqt = (p.T + 127) / 128;
  const long total_tiles = (long)p.B * p.n_head * nqt;
  const int t_begin = (int)(total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)(total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("sdxl_b200: attention smem base not 1024B aligned\n");
      __trap();
    }
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();  // PDL: the prologue above overlapped the previous kernel's tail
  griddep_launch_dependents();

  for (int tile = t_begin; tile < t_end;) {
  const int qt = tile % nqt;
  const int head = (tile / nqt) % p.n_head;
  const int b = tile / (nqt * p.n_head);
  const int row0 = qt * 128;
  const bool hasB = (qt + 1 < nqt) && (tile + 1 < t_end);   // pair with the next tile of the same head
  tile += hasB ? 2 : 1;
  // fresh barriers for every work item (nobody is using them here: see the __syncthreads at the loop end)
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128 * SPLIT);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, (hasB ? 2 : 1) * kTileBytes);
      tma_load_3d(sQ, &p.tmQ, q_full, p.q_col0 + head * 64, row0, b);
      if (hasB) tma_load_3d(sQ + kTileBytes, &p.tmQ, q_full, p.q_col0 + head * 64, row0 + 128, b);
      for (int j = 0; j < nblk; ++j) {
        const int stage = j & 1;
        mbar_wait(&kv_empty[stage], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[stage], 2 * kTileBytes);
        tma_load_3d(sK + stage * kTileBytes, &p.tmK, &kv_full[stage], p.k_col0 + head * 64, j * 128, b);
        tma_load_3d(sV + stage * kTileBytes, &p.tmV, &kv_full[stage], p.v_col0 + head * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-convergent loop, one elected lane issues =====================
    const uint32_t idesc_qk = make_idesc_f16(128, false);
    const uint32_t idesc_pv = make_idesc_f16(64, true);
    const uint32_t dhi = 64u | (1u << 14) | (2u << 29);   // SBO=1024B, version, SWIZZLE_128B
    const uint32_t lo_flag = 1u << 16;                     // LBO(enc)=1
    const uint32_t q_lo0 = ((smem_u32(sQ) >> 4) & 0x3FFFu) | lo_flag;
    const uint32_t k_lo0 = ((smem_u32(sK) >> 4) & 0x3FFFu) | lo_flag;
    const uint32_t v_lo0 = ((smem_u32(sV) >> 4) & 0x3FFFu) | lo_flag;
    const uint32_t p_lo0 = ((smem_u32(sP) >> 4) & 0x3FFFu) | lo_flag;
    constexpr uint32_t kTile16 = kTileBytes >> 4, kP16 = kPBytes >> 4;
    // S_x = Q_x K_stage^T
    auto issue_qk = [&](int x, int stage) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_f16_elect(tmem_base + x * 128, q_lo0 + x * kTile16 + 2 * k, k_lo0 + stage * kTile16 + 2 * k, dhi, idesc_qk, k > 0);
      tc_commit_elect(&s_full[x]);
    };
    // O_x[j&1] = P_x V_j   (P panel t>>2 is 16 KB further, 32 B per 16 keys; V: 16 key rows = 2048 B)
    auto issue_pv = [&](int x, int j) {
      const uint32_t d = tmem_base + 256 + x * 128 + (j & 1) * 64;
#pragma unroll
      for (int t = 0; t < 8; ++t)
        tc_mma_f16_elect(d, p_lo0 + x * kP16 + (t >> 2) * 1024 + (t & 3) * 2, v_lo0 + (j & 1) * kTile16 + t * 128, dhi, idesc_pv,
                         t > 0);
      tc_commit_elect(&pv_done[x]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_qk(0, 0);
    if (hasB) issue_qk(1, 0);
    for (int j = 0; j < nblk; ++j) {
      const bool more = j + 1 < nblk;
      mbar_wait(&p_full[0], j & 1);  // P_A(j) written, S_A(j) consumed
      tc_fence_after();
      issue_pv(0, j);
      if (more) {
        mbar_wait(&kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        issue_qk(0, (j + 1) & 1);
      }
      if (hasB) {
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        issue_pv(1, j);
      }
      tc_commit_elect(&kv_empty[j & 1]);  // K_j / V_j no longer needed once everything issued so far retires
      if (more && hasB) issue_qk(1, (j + 1) & 1);
    }
    // drain: the last stage releases have no consumer; observe them so that no asynchronous arrival is still
    // in flight when the barriers are re-initialised for the next work item
    for (int s = 0; s < 2 && s < nblk; ++s) {
      const int uses = (nblk - s + 1) >> 1;  // blocks j with (j & 1) == s
      mbar_wait(&kv_empty[s], (uses - 1) & 1);
    }
  } else {
    if constexpr (SPLIT == 1) {
    const int x = (warp - 2) >> 2;  // softmax group: 0 = tile A, 1 = tile B
    if (x == 0 || hasB) {
      const int q = warp & 3;
      const int r = q * 32 + lane;  // query row in tile == TMEM lane
      const uint32_t lane_off = (uint32_t)(q * 32) << 16;
      const uint32_t tS = tmem_base + lane_off + x * 128;
      const uint32_t tO = tmem_base + lane_off + 256 + x * 128;
      const float sl2e = p.scale_log2e;
      float m = -INFINITY, m_prev = -INFINITY, l = 0.f, alpha_prev = 0.f;
      float O[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) O[i] = 0.f;
      uint8_t* prow = sP + x * kPBytes + (r >> 3) * 1024 + (r & 7) * 128;
      const int rx = r & 7;

      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&s_full[x], j & 1);
        // P V of block j-1 was issued before Q K^T of block j, so it has retired too. Observe its phase NOW,
        // before this thread's p_full arrival lets the MMA warp issue P V of block j (a waiter must never
        // fall two phases behind an mbarrier).
        if (j > 0) mbar_wait(&pv_done[x], (j - 1) & 1);
        tc_fence_after();
        const int kbase = j * 128;
        const bool ragged = kbase + 128 > p.S;
        // Reference for the exponent. Block 0: exact row max (one extra pass over S). Later blocks: the reference
        // decided at the end of the previous block (lazy max): p = exp2((s - ref) * c) may exceed 1 (f16 P and f32
        // sums have the head-room), and the row's running max is folded in only when it grew by more than 2^8.
        // An overflow-safe redo (warp-uniform, practically never taken) re-runs the block with the exact max.
        float ref = m;
        if (j == 0) {
          float mx = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            uint32_t v[32];
            tmem_ld32(tS + c, v);
            tmem_ld_wait();
            if (!ragged) {
              float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;  // four short chains
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                m0 = fmaxf(m0, __uint_as_float(v[i]));
                m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
                m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
                m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
              }
              mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (kbase + c + i < p.S) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
          }
          ref = mx;
        }
        float alpha, sum, bmax;
        bool redo;
        do {
          alpha = ex2_approx((m_prev - ref) * sl2e);   // rescale of everything accumulated so far (0 for block 0)
          const float mb = ref * sl2e;
          sum = 0.f;
          float b0 = -INFINITY, b1 = -INFINITY;
          // p = exp2(s*scale - ref*scale) -> f16 -> swizzled smem (A operand of the PV MMA); track the block max
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            uint32_t v[32];
            tmem_ld32(tS + c, v);
            tmem_ld_wait();
            uint32_t h[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
              if (ragged) {
                if (kbase + c + i >= p.S) s0 = -INFINITY;
                if (kbase + c + i + 1 >= p.S) s1 = -INFINITY;
              }
              b0 = fmaxf(b0, s0);
              b1 = fmaxf(b1, s1);
              const float p0 = ex2_approx(fmaf(s0, sl2e, -mb));
              const float p1 = ex2_approx(fmaf(s1, sl2e, -mb));
              sum += p0 + p1;
              __half2 t = __floats2half2_rn(p0, p1);
              h[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
            }
            uint8_t* panel = prow + (c >> 6) * (128 * 128);
            const int ch0 = (c & 63) >> 3;  // first 16B chunk of this 32-key group inside the 128B row
#pragma unroll
            for (int u = 0; u < 4; ++u)
              *reinterpret_cast<uint4*>(panel + (((ch0 + u) ^ rx) << 4)) =
                  make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
          }
          bmax = fmaxf(b0, b1);
          // f16 P overflows beyond 2^16: redo the whole block (all lanes: the TMEM loads are warp-collective)
          const bool over = (bmax - ref) * sl2e > 15.0f;
          redo = __any_sync(0xffffffffu, over);
          if (over) ref = bmax;
        } while (redo);
        l = l * alpha + sum;
        m_prev = ref;
        m = ((bmax - ref) * sl2e > 8.0f) ? bmax : ref;   // reference for the next block
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[x]);
        // fold in the PREVIOUS block's P V (already complete: it was issued before this block's Q K^T)
        if (j > 0) {
          const uint32_t to = tO + ((j - 1) & 1) * 64;
#pragma unroll
          for (int c = 0; c < 64; c += 32) {
            uint32_t v[32];
            tmem_ld32(to + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) O[c + i] = fmaf(O[c + i], alpha_prev, __uint_as_float(v[i]));
          }
        }
        alpha_prev = alpha;
      }
      {
        const int j = nblk - 1;
        mbar_wait(&pv_done[x], j & 1);
        tc_fence_after();
        const uint32_t to = tO + (j & 1) * 64;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          uint32_t v[32];
          tmem_ld32(to + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) O[c + i] = fmaf(O[c + i], alpha_prev, __uint_as_float(v[i]));
        }
      }
      const int t = row0 + x * 128 + r;
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
    } else {
    // ---------- SPLIT == 2: two threads per query row ----------
    const int sw = warp - 2;              // 0..15
    const int x = sw >> 3;                // softmax group: 0 = tile A, 1 = tile B
    if (x == 0 || hasB) {
      const int q = warp & 3;             // TMEM lane quarter this warp may access (hardware: warp id % 4)
      const int ch = (sw >> 2) & 1;       // column half: score columns [64 ch, 64 ch + 64), output columns [32 ch, 32 ch + 32)
      const int r = q * 32 + lane;        // query row in tile == TMEM lane
      const uint32_t lane_off = (uint32_t)(q * 32) << 16;
      const uint32_t tS = tmem_base + lane_off + x * 128 + ch * 64;
      const uint32_t tO = tmem_base + lane_off + 256 + x * 128 + ch * 32;
      const float sl2e = p.scale_log2e;
      const int bar_id = 1 + x * 4 + q;   // named barrier of the two warps that share these 32 rows
      int xk = 0;                         // exchange counter (slot = xk & 1)
      // combine a per-thread value with the partner thread of the same row (other column half)
      auto exchange = [&](float v) {
        float* slot = xchg + (((xk & 1) * 2 + x) * 2) * 128;
        slot[ch * 128 + r] = v;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float o = slot[(ch ^ 1) * 128 + r];
        ++xk;
        return o;
      };
      float m = -INFINITY, m_prev = -INFINITY, l = 0.f, alpha_prev = 0.f;
      float O[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) O[i] = 0.f;
      uint8_t* prow = sP + x * kPBytes + ch * (128 * 128) + (r >> 3) * 1024 + (r & 7) * 128;   // my 64-key panel
      const int rx = r & 7;

      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&s_full[x], j & 1);
        if (j > 0) mbar_wait(&pv_done[x], (j - 1) & 1);
        tc_fence_after();
        const int kbase = j * 128 + ch * 64;
        const bool ragged = j * 128 + 128 > p.S;
        float ref = m;
        if (j == 0) {
          float mx = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            uint32_t v[32];
            tmem_ld32(tS + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!ragged || kbase + c + i < p.S) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
          ref = fmaxf(mx, exchange(mx));
        }
        float alpha, sum, bmax;
        bool redo;
        do {
          alpha = ex2_approx((m_prev - ref) * sl2e);
          const float mb = ref * sl2e;
          sum = 0.f;
          float b0 = -INFINITY, b1 = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            uint32_t v[32];
            tmem_ld32(tS + c, v);
            tmem_ld_wait();
            uint32_t h[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
              if (ragged) {
                if (kbase + c + i >= p.S) s0 = -INFINITY;
                if (kbase + c + i + 1 >= p.S) s1 = -INFINITY;
              }
              b0 = fmaxf(b0, s0);
              b1 = fmaxf(b1, s1);
              const float p0 = ex2_approx(fmaf(s0, sl2e, -mb));
              const float p1 = ex2_approx(fmaf(s1, sl2e, -mb));
              sum += p0 + p1;
              __half2 t = __floats2half2_rn(p0, p1);
              h[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
            }
            const int ch0 = c >> 3;  // first 16B chunk of this 32-key group inside my panel's 128B row
#pragma unroll
            for (int u = 0; u < 4; ++u)
              *reinterpret_cast<uint4*>(prow + (((ch0 + u) ^ rx) << 4)) = make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
          }
          const float bh = fmaxf(b0, b1);
          bmax = fmaxf(bh, exchange(bh));   // block max of the whole row: both threads take identical decisions
          const bool over = (bmax - ref) * sl2e > 15.0f;
          redo = __any_sync(0xffffffffu, over);   // same rows in both warps of the pair -> same vote
          if (over) ref = bmax;
        } while (redo);
        l = l * alpha + sum;   // partial sum over my columns; the halves are added once at the end
        m_prev = ref;
        m = ((bmax - ref) * sl2e > 8.0f) ? bmax : ref;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[x]);
        if (j > 0) {
          uint32_t v[32];
          tmem_ld32(tO + ((j - 1) & 1) * 64, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) O[i] = fmaf(O[i], alpha_prev, __uint_as_float(v[i]));
        }
        alpha_prev = alpha;
      }
      {
        const int j = nblk - 1;
        mbar_wait(&pv_done[x], j & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tO + (j & 1) * 64, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) O[i] = fmaf(O[i], alpha_prev, __uint_as_float(v[i]));
      }
      const float lt = l + exchange(l);
      const int t = row0 + x * 128 + r;
      if (t < p.T) {
        const float inv = 1.0f / lt;
        __half* o = p.out + ((size_t)b * p.T + t) * p.ldo + head * 64 + ch * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
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
    }
  }

  tc_fence_before();
  __syncthreads();   // every role is done with this item's barriers, smem and TMEM
  tc_fence_after();
  }  // work items

  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int attention_launch(cudaStream_t st, const AttnParams& p) {
  // SDXL_B200_ATTN_SPLIT=2 selects the two-threads-per-row softmax (16 softmax warps); default 1
  static const int split = (getenv("SDXL_B200_ATTN_SPLIT") && atoi(getenv("SDXL_B200_ATTN_SPLIT")) == 2) ? 2 : 1;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const long tiles = (long)p.B * p.n_head * ((p.T + 127) / 128);
  const long want = (tiles + 1) / 2;  // one pair per CTA when the machine is not full
  dim3 grid((unsigned)(want < num_sms ? (want > 0 ? want : 1) : num_sms));
  if (split == 2) return launch_kernel(attention_kernel<2>, grid, dim3(AttnCfg<2>::kThreads), (size_t)kAttnSmem, st, true, p);
  return launch_kernel(attention_kernel<1>, grid, dim3(AttnCfg<1>::kThreads), (size_t)kAttnSmem, st, true, p);
}

}  // namespace sdxl
