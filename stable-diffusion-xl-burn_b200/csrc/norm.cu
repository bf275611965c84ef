// GroupNorm(+SiLU) and LayerNorm for the UNet step. HBM-bound: each input element is read twice
// (stats, apply) / once (LayerNorm, row held in registers) with 128-bit loads; reductions use
// warp shuffles + a small deterministic partial-sum table (no atomics).
//   reference: groupnorm/mod.rs:52-82 (reshape to [B,32,C/32*HW], biased variance, eps inside sqrt,
//   per-channel affine), layernorm/mod.rs:34-49, silu.rs:14-16.
#include "common.cuh"
#include "kernels.h"

namespace sdxl {

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
static constexpr int kGnMaxChunk = 512;

size_t gn_scratch_floats(int B, int n_group) { return (size_t)B * kGnMaxChunk * n_group * 2; }

// ---- stats: grid (nchunk, B); block = V*R threads, V = C/4 float4 columns, R pixel rows in flight
__global__ void gn_stats_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2, int HW,
                                int n_group, int R, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [R][C][2]
  griddep_wait();
  griddep_launch_dependents();
  const int C = C1 + C2;
  const int V = C >> 2;
  const int v = threadIdx.x % V, rr = threadIdx.x / V;
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int per = (HW + nchunk - 1) / nchunk;
  const int p0 = chunk * per;
  const int p1 = min(HW, p0 + per);
  const int c = v * 4;
  const float* src;
  int cc, Cs;
  if (c < C1) { src = x1; cc = c; Cs = C1; } else { src = x2; cc = c - C1; Cs = C2; }
  src += (size_t)b * HW * Cs + cc;
  float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
  int p = p0 + rr;
  for (; p + 7 * R < p1; p += 8 * R) {  // eight independent 128-bit loads in flight per thread
    float4 a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = *reinterpret_cast<const float4*>(src + (size_t)(p + u * R) * Cs);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s.x += a[u].x; s.y += a[u].y; s.z += a[u].z; s.w += a[u].w;
      q.x = fmaf(a[u].x, a[u].x, q.x); q.y = fmaf(a[u].y, a[u].y, q.y);
      q.z = fmaf(a[u].z, a[u].z, q.z); q.w = fmaf(a[u].w, a[u].w, q.w);
    }
  }
  for (; p < p1; p += R) {
    const float4 a = *reinterpret_cast<const float4*>(src + (size_t)p * Cs);
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    q.x = fmaf(a.x, a.x, q.x); q.y = fmaf(a.y, a.y, q.y); q.z = fmaf(a.z, a.z, q.z); q.w = fmaf(a.w, a.w, q.w);
  }
  float* row = sm + ((size_t)rr * C + c) * 2;
  row[0] = s.x; row[1] = q.x; row[2] = s.y; row[3] = q.y; row[4] = s.z; row[5] = q.z; row[6] = s.w; row[7] = q.w;
  __syncthreads();
  if (threadIdx.x < n_group) {
    const int g = threadIdx.x, cpg = C / n_group;
    float S = 0.f, Q = 0.f;
    for (int r = 0; r < R; ++r)
      for (int j = 0; j < cpg; ++j) {
        const float* e = sm + ((size_t)r * C + g * cpg + j) * 2;
        S += e[0];
        Q += e[1];
      }
    float* o = partial + (((size_t)b * nchunk + chunk) * n_group + g) * 2;
    o[0] = S;
    o[1] = Q;
  }
}

// ---- apply: grid (ctas, B); 8 channels per thread
__global__ void gn_apply_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2, int HW,
                                int n_group, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float eps, int silu, const float* __restrict__ partial, int nchunk,
                                __half* __restrict__ y, __half* __restrict__ raw) {
  extern __shared__ float sm[];  // scale[C], shift[C], mean[G], rstd[G]
  griddep_wait();
  griddep_launch_dependents();
  const int C = C1 + C2;
  float* sc = sm;
  float* sh = sm + C;
  float* mean = sh + C;
  float* rstd = mean + n_group;
  const int b = blockIdx.y;
  const int cpg = C / n_group;
  // finalize the statistics: 8 lanes per group walk the chunk partials (fixed order => deterministic)
  for (int g = threadIdx.x >> 3; g < n_group; g += blockDim.x >> 3) {
    const int sub = threadIdx.x & 7;
    double S = 0.0, Q = 0.0;
    for (int k = sub; k < nchunk; k += 8) {
      const float2 e = *reinterpret_cast<const float2*>(partial + (((size_t)b * nchunk + k) * n_group + g) * 2);
      S += (double)e.x;
      Q += (double)e.y;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      S += __shfl_xor_sync(0xffffffffu, S, o);
      Q += __shfl_xor_sync(0xffffffffu, Q, o);
    }
    if (sub == 0) {
      const double n = (double)cpg * HW;
      const double m = S / n;
      double var = Q / n - m * m;
      if (var < 0.0) var = 0.0;
      mean[g] = (float)m;
      rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float s = rstd[g] * gamma[c];
    sc[c] = s;
    sh[c] = beta[c] - mean[g] * s;
  }
  __syncthreads();
  const int V8 = C >> 3;
  const long total = (long)HW * V8;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long idx0 = (long)blockIdx.x * blockDim.x + threadIdx.x; idx0 < total; idx0 += 2 * stride) {
    // two independent items per trip: both 32 B loads are issued before either is consumed
    float4 la[2][2];
    const float* srcs[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long idx = idx0 + u * stride;
      if (idx < total) {
        const int v = (int)(idx % V8);
        const long p = idx / V8;
        const int c = v * 8;
        srcs[u] = (c < C1) ? x1 + ((size_t)b * HW + p) * C1 + c : x2 + ((size_t)b * HW + p) * C2 + (c - C1);
        la[u][0] = *reinterpret_cast<const float4*>(srcs[u]);
        la[u][1] = *reinterpret_cast<const float4*>(srcs[u] + 4);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
    const long idx = idx0 + u * stride;
    if (idx >= total) continue;
    const int v = (int)(idx % V8);
    const long p = idx / V8;
    const int c = v * 8;
    const float4 a0 = la[u][0], a1 = la[u][1];
    float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    uint32_t h[4];
    if (raw) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        h[i] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(raw + ((size_t)b * HW + p) * C + c) = make_uint4(h[0], h[1], h[2], h[3]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = fmaf(f[i], sc[c + i], sh[c + i]);
      if (silu) t = silu_f(t);
      f[i] = t;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      h[i] = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(y + ((size_t)b * HW + p) * C + c) = make_uint4(h[0], h[1], h[2], h[3]);
    }
  }
}

int gn_launch(cudaStream_t st, GnParams& p) {
  const int C = p.C1 + p.C2;
  if ((p.C1 & 7) || (p.C2 & 7) || C % p.n_group || p.n_group > 64) return 3001;
  const int V = C / 4;
  if (V > 1024) return 3002;
  int R = 512 / V;
  if (R < 1) R = 1;
  if (R > 16) R = 16;
  // chunking depends on HW only (never on the batch size): every sample's statistics are summed in the
  // same order whatever B is, which keeps the whole forward batch-invariant bit for bit.
  int nchunk = 148;
  const int max_chunks = cdiv(p.HW, R);
  if (nchunk > max_chunks) nchunk = max_chunks;
  p.nchunk = nchunk;
  const size_t smem1 = (size_t)R * C * 2 * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr = true;
  }
  int e = launch_kernel(gn_stats_kernel, dim3(nchunk, p.B), dim3(V * R), smem1, st, true, p.x1, p.C1, p.x2, p.C2, p.HW,
                        p.n_group, R, p.partial);
  if (e) return e;
  const size_t smem2 = (size_t)(2 * C + 2 * p.n_group) * sizeof(float);
  int ctas = cdiv((long)p.HW * (C / 8), 256 * 4);
  const int cap = (148 * 8) / (p.B > 0 ? p.B : 1);
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return launch_kernel(gn_apply_kernel, dim3(ctas, p.B), dim3(256), smem2, st, true, p.x1, p.C1, p.x2, p.C2, p.HW,
                       p.n_group, p.gamma, p.beta, p.eps, p.silu, (const float*)p.partial, nchunk, p.y, p.raw);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row cached in registers (NV float4 per lane), exact two-pass
// (mean, then centred second moment) as the reference's layernorm() does.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, int rows, int C,
                                 __half* __restrict__ y) {
  griddep_wait();
  griddep_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < C) {
      v[i] = *reinterpret_cast<const float4*>(xr + c);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    } else {
      v[i] = make_float4(0, 0, 0, 0);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < C) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < C) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
      __half2 a = __floats2half2_rn(v[i].x * rstd * g.x + bt.x, v[i].y * rstd * g.y + bt.y);
      __half2 b = __floats2half2_rn(v[i].z * rstd * g.z + bt.z, v[i].w * rstd * g.w + bt.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(y + (size_t)row * C + c) = o;
    }
  }
}

int layernorm_launch(cudaStream_t st, const float* x, const float* gamma, const float* beta, float eps, int rows,
                     int C, __half* y) {
  if (C & 3) return 3003;
  const int nv = cdiv(C, 128);
  const int warps = 8;
  dim3 grid(cdiv(rows, warps));
  if (nv <= 1) return launch_kernel(layernorm_kernel<1>, grid, dim3(warps * 32), (size_t)0, st, true, x, gamma, beta, eps, rows, C, y);
  else if (nv <= 2) return launch_kernel(layernorm_kernel<2>, grid, dim3(warps * 32), (size_t)0, st, true, x, gamma, beta, eps, rows, C, y);
  else if (nv <= 5) return launch_kernel(layernorm_kernel<5>, grid, dim3(warps * 32), (size_t)0, st, true, x, gamma, beta, eps, rows, C, y);
  else if (nv <= 10) return launch_kernel(layernorm_kernel<10>, grid, dim3(warps * 32), (size_t)0, st, true, x, gamma, beta, eps, rows, C, y);
  else if (nv <= 16) return launch_kernel(layernorm_kernel<16>, grid, dim3(warps * 32), (size_t)0, st, true, x, gamma, beta, eps, rows, C, y);
  return 3004;
}

}  // namespace sdxl
