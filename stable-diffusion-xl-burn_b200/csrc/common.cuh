// Shared device-side primitives for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers
// (inline PTX), warp reductions and small math helpers. Everything here is Blackwell-only.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace sdxl {

// ---------------------------------------------------------------------------------------------
// generic helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// erf by Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, far below the f16 rounding of the GEGLU output):
// one MUFU.RCP + one MUFU.EX2 + 8 FMA instead of libdevice erff's ~25 instructions.
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  const float poly =
      t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-ax * ax * 1.4426950408889634f));
  return copysignf(1.0f - poly * e, x);
}
// erf-form GELU (burn::nn::Gelu, reference unet/mod.rs:930,954)
__device__ __forceinline__ float gelu_erf_f(float x) {
  return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752440f));
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");   // "memory": not to be moved across barriers (timeline stamps)
  return t;
}

// Programmatic dependent launch (PDL). wait: block until the preceding kernel in the stream has completed
// and its memory is visible (no-op when the kernel was launched without the PDL attribute).
// launch_dependents: allow the next kernel's CTAs to start their prologue as SM resources free up.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch error) instead of hanging the
// GPU. The bound (~4 s) is far above any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (globaltimer_ns() - t0 > 4000000000ull) {
      printf("sdxl_b200: mbarrier wait timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrive(1) on the mbarrier once all previously issued MMAs of this thread retire.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], f16 inputs, f32 accumulate. One thread issues.
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent issue helpers: executed by ALL lanes of a converged warp with warp-uniform operands; the
// instruction itself is predicated on elect.sync, so exactly one lane issues it. Keeping the surrounding control
// flow convergent lets ptxas hold descriptors / addresses in uniform registers instead of emitting a per-instruction
// ELECT + R2UR.BROADCAST waterfall (which made the single-lane issue loops the bottleneck of the GEMM main loop).
__device__ __forceinline__ void tc_mma_f16_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                      uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One K block (four K=16 MMAs on consecutive 32-byte descriptor steps) + the commit that releases its smem slot,
// under a single elect.sync: the issue loop's cost per K block is what bounds the GEMM main loop.
// `acc_first` = accumulate flag of the first MMA (0 only for the first K block of a tile).
__device__ __forceinline__ void tc_mma4_commit_pair_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                          uint32_t idesc, uint32_t acc_first, uint64_t* bar) {
  // every instruction is predicated on the elect result ONLY: ptxas then keeps the whole sequence on the uniform datapath
  // (UTCHMMA with UR operands back to back); mixing further conditions into the predicate brings back a per-MMA
  // ELECT / R2UR.BROADCAST waterfall loop.
  asm volatile(
      "{\n\t"
      ".reg .pred p, e, t;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 al, bl;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
      "add.u32 al, %1, 2;\n\t"
      "add.u32 bl, %2, 2;\n\t"
      "mov.b64 da, {al, %3};\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, t;\n\t"
      "add.u32 al, %1, 4;\n\t"
      "add.u32 bl, %2, 4;\n\t"
      "mov.b64 da, {al, %3};\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, t;\n\t"
      "add.u32 al, %1, 6;\n\t"
      "add.u32 bl, %2, 6;\n\t"
      "mov.b64 da, {al, %3};\n\t"
      "mov.b64 db, {bl, %3};\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, t;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], %7;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc_first), "r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}"
      ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc_elect(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

// Shared-memory matrix descriptor for a SWIZZLE_128B tile whose rows are 128 bytes (64 halves):
// 8-row swizzle atoms of 1024 B stacked along the outer dimension (SBO = 1024 B). Valid both for
// K-major operands (rows = M/N index, 64 K-elements per row) and MN-major operands (rows = K index,
// 64 MN-elements per row); the major-ness is selected in the instruction descriptor.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);           // start addr, LBO(enc)=1
  const uint32_t hi = 64u /*SBO=1024B>>4*/ | (1u << 14) /*version*/ | (2u << 29) /*SWIZZLE_128B*/;
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// Instruction descriptor: f16 x f16 -> f32, M=128, N=n, A K-major, B K-major or MN-major.
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t n, bool b_mn_major) {
  return (1u << 4) | (b_mn_major ? (1u << 16) : 0u) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// TMEM -> registers: this warp's 32 lanes (lane quarter = warp_id % 4), 16 consecutive columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace sdxl
