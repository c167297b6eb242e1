// HBM-bound row kernels of the MMinterest step: LayerNorm fwd/bwd, bias-gradient column
// sums, the Linear(d->1) head, the fused focal-loss/diagnostics kernel, global-norm clip +
// AdamW, and the fp32->bf16 parameter cast.  Reductions use warp shuffles; cross-CTA
// reductions are two-stage (fixed order => run-to-run deterministic).
#include "common.cuh"
#include "dropout.cuh"
#include <string.h>

namespace mmi {

constexpr int kMaxVecPerLane = 8;   // d <= 1024
#ifndef MMI_LN_PREFETCH
#define MMI_LN_PREFETCH 0            // L2 prefetch of the next rows in the LayerNorm kernels: measured on B200 at c2 and left off
                                     // (-DMMI_LN_PREFETCH=1: ln_bwd 5.1 -> 5.6 ms/step, ln_fwd 2.35 -> 2.28 ms/step)
#endif
constexpr int kRedCtas = kNumSMs * 4;

// ------------------------------------------------------------------ LayerNorm forward
template <typename T, int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const T* __restrict__ x, int64_t rows, int d,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, T* __restrict__ y, float* __restrict__ stats,
                                                            const DropParams drop) {
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    const T* xr = x + row * (int64_t)d;
    if (MMI_LN_PREFETCH && row + wstride < rows) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + (row + wstride) * (int64_t)d + c * 4));
      }
    }
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      v[i] = (c < nvec) ? load4(xr + c * 4) : make_float4(0, 0, 0, 0);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, e = v[i].w - mean;
        q += a * a + b * b + cc * cc + e * e;
      }
    }
    const float var = warp_sum(q) / (float)d;
    const float rstd = 1.0f / sqrtf(var + eps);
    if (lane == 0 && stats != nullptr) {
      stats[2 * row] = mean;
      stats[2 * row + 1] = rstd;
    }
    T* yr = y + row * (int64_t)d;
    const uint32_t rowh = drop.thr8 ? drop_rowhash(drop.key, (uint64_t)row) : 0u;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + c * 4);
        const float4 b = *reinterpret_cast<const float4*>(beta + c * 4);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (drop.thr8) {                                   // y = dropout(LN(x)): lane owns columns 4c .. 4c+3
          const uint32_t w = drop_keep_word(rowh, (uint32_t)c >> 3, drop.thr8) >> ((c & 7) * 4);
          o.x *= (w & 1u) ? drop.scale : 0.f; o.y *= (w & 2u) ? drop.scale : 0.f;
          o.z *= (w & 4u) ? drop.scale : 0.f; o.w *= (w & 8u) ? drop.scale : 0.f;
        }
        store4(yr + c * 4, o);
      }
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// partial layout: [gridDim.x][2][d]  (dgamma then dbeta)
template <typename T, int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, int64_t rows, int d,
                                                            const float* __restrict__ gamma, const float* __restrict__ stats,
                                                            const T* __restrict__ add, T* __restrict__ dx,
                                                            float* __restrict__ partial, const DropParams dy_drop,
                                                            const DropParams dx_drop, T* __restrict__ dx_dropped) {
  extern __shared__ float sm[];  // [8 warps][2][d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = d >> 2;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = make_float4(0, 0, 0, 0);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; row < rows; row += wstride) {
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    const T* xr = x + row * (int64_t)d;
    const T* dyr = dy + row * (int64_t)d;
    float4 xh[NV], g[NV];
    float c1 = 0.f, c2 = 0.f;
    const uint32_t rowh_y = dy_drop.thr8 ? drop_rowhash(dy_drop.key, (uint64_t)row) : 0u;
    const uint32_t rowh_x = dx_drop.thr8 ? drop_rowhash(dx_drop.key, (uint64_t)row) : 0u;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 xv = load4(xr + c * 4);
        float4 dv = load4(dyr + c * 4);
        if (dy_drop.thr8) {                                // backward of y = dropout(LN(x)): dy <- mask * scale * dy
          const uint32_t w = drop_keep_word(rowh_y, (uint32_t)c >> 3, dy_drop.thr8) >> ((c & 7) * 4);
          dv.x *= (w & 1u) ? dy_drop.scale : 0.f; dv.y *= (w & 2u) ? dy_drop.scale : 0.f;
          dv.z *= (w & 4u) ? dy_drop.scale : 0.f; dv.w *= (w & 8u) ? dy_drop.scale : 0.f;
        }
        const float4 gm = *reinterpret_cast<const float4*>(gamma + c * 4);
        xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        g[i] = make_float4(dv.x * gm.x, dv.y * gm.y, dv.z * gm.z, dv.w * gm.w);
        c1 += g[i].x + g[i].y + g[i].z + g[i].w;
        c2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
        dg[i].x += dv.x * xh[i].x; dg[i].y += dv.y * xh[i].y; dg[i].z += dv.z * xh[i].z; dg[i].w += dv.w * xh[i].w;
        db[i].x += dv.x; db[i].y += dv.y; db[i].z += dv.z; db[i].w += dv.w;
      } else {
        xh[i] = g[i] = make_float4(0, 0, 0, 0);
      }
    }
    c1 = warp_sum(c1) / (float)d;
    c2 = warp_sum(c2) / (float)d;
    T* dxr = dx + row * (int64_t)d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float4 o;
        o.x = rstd * (g[i].x - c1 - xh[i].x * c2);
        o.y = rstd * (g[i].y - c1 - xh[i].y * c2);
        o.z = rstd * (g[i].z - c1 - xh[i].z * c2);
        o.w = rstd * (g[i].w - c1 - xh[i].w * c2);
        if (add != nullptr) {
          const float4 a = load4(add + row * (int64_t)d + c * 4);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        store4(dxr + c * 4, o);
        if (dx_drop.thr8) {                                // gradient into the Linear behind the residual's dropout
          const uint32_t w = drop_keep_word(rowh_x, (uint32_t)c >> 3, dx_drop.thr8) >> ((c & 7) * 4);
          o.x *= (w & 1u) ? dx_drop.scale : 0.f; o.y *= (w & 2u) ? dx_drop.scale : 0.f;
          o.z *= (w & 4u) ? dx_drop.scale : 0.f; o.w *= (w & 8u) ? dx_drop.scale : 0.f;
          store4(dx_dropped + row * (int64_t)d + c * 4, o);
        }
      }
    }
  }
  // CTA reduction of the per-warp column sums (fixed order)
  float* mine = sm + (size_t)warp * 2 * d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      *reinterpret_cast<float4*>(mine + c * 4) = dg[i];
      *reinterpret_cast<float4*>(mine + d + c * 4) = db[i];
    }
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += sm[(size_t)w * 2 * d + c];
    partial[(size_t)blockIdx.x * 2 * d + c] = s;
  }
}

// bf16 fast path (d = 256 * NV): lane owns 8 consecutive columns per 16-byte vector, two rows per warp iteration so
// that 128 B per lane are in flight; optionally also accumulates the column sums of dx (the bias gradient of the
// Linear that produced the LayerNorm input -- saves a separate pass over dx).
// partial layout: [gridDim.x][3][d]  (dgamma, dbeta, dxsum)
__device__ __forceinline__ void bf16x8_to_f32(const uint4& r, float* f) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 f32_to_bf16x8(const float* f) {
  __nv_bfloat162 p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return make_uint4(*reinterpret_cast<uint32_t*>(&p[0]), *reinterpret_cast<uint32_t*>(&p[1]), *reinterpret_cast<uint32_t*>(&p[2]), *reinterpret_cast<uint32_t*>(&p[3]));
}

// keep factors of the 8 consecutive columns (lane + 32 i) * 8 .. + 7 of one row: byte (lane & 3) of keep word (lane + 32 i) >> 2
__device__ __forceinline__ void drop_factors8(const DropParams& dp, uint32_t rowh, int vec, float* kf) {
  const uint32_t w = (drop_keep_word(rowh, (uint32_t)vec >> 2, dp.thr8) >> ((vec & 3) * 8)) & 0xffu;
#pragma unroll
  for (int j = 0; j < 8; ++j) kf[j] = ((w >> j) & 1u) ? dp.scale : 0.f;
}

// LayerNorm forward, bf16 fast path (d = 256 * NV): rows travel through a per-warp cp.async ring (3 stages x 2 rows) so the
// loads of the next two row pairs are in flight while this pair is normalised; 16-byte vectors, gamma / beta in registers.
// (The generic kernel: one row per warp with 8-byte loads, 58 % of the HBM peak at 69 % issue utilisation.)
template <int NV, bool DROP>
__global__ void __launch_bounds__(256) layernorm_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                                 __nv_bfloat16* __restrict__ y, float* __restrict__ stats, const DropParams drop) {
  extern __shared__ float sm[];
  constexpr int d = 256 * NV;
  constexpr int LN_STAGES = 3;
  constexpr uint32_t STAGE_BYTES = 2 * d * 2;            // 2 rows x d bf16
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wstride = (int64_t)gridDim.x * 8;
  uint8_t* ring = reinterpret_cast<uint8_t*>(sm) + (size_t)warp * LN_STAGES * STAGE_BYTES;
  const uint32_t ring_a = static_cast<uint32_t>(__cvta_generic_to_shared(ring));
  float gm[NV][8], bt[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 a = *reinterpret_cast<const float4*>(gamma + (lane + 32 * i) * 8 + j), b = *reinterpret_cast<const float4*>(beta + (lane + 32 * i) * 8 + j);
      gm[i][j] = a.x; gm[i][j + 1] = a.y; gm[i][j + 2] = a.z; gm[i][j + 3] = a.w;
      bt[i][j] = b.x; bt[i][j + 1] = b.y; bt[i][j + 2] = b.z; bt[i][j + 3] = b.w;
    }
  }
  auto issue = [&](int64_t r0, int stage) {
    if (r0 < rows) {
      const int64_t r1 = r0 + wstride < rows ? r0 + wstride : r0;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < NV; ++i)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_a + stage * STAGE_BYTES + (r * NV * 32 + lane + 32 * i) * 16),
                       "l"(reinterpret_cast<const uint4*>(x + (r ? r1 : r0) * (int64_t)d) + lane + 32 * i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int64_t first = (int64_t)blockIdx.x * 8 + warp;
  issue(first, 0);
  issue(first + 2 * wstride, 1);
  int stage = 0;
  for (int64_t row0 = first; row0 < rows; row0 += 2 * wstride) {
    issue(row0 + 4 * wstride, stage >= 1 ? stage - 1 : LN_STAGES - 1);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    const uint8_t* cur = ring + stage * STAGE_BYTES + lane * 16;
    stage = stage + 1 == LN_STAGES ? 0 : stage + 1;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int64_t row = row0 + r * wstride;
      if (row >= rows) break;
      float v[NV][8];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        bf16x8_to_f32(*reinterpret_cast<const uint4*>(cur + (r * NV * 32 + 32 * i) * 16), v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
      const float mean = warp_sum(s) * (1.0f / d);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float a = v[i][j] - mean; q = fmaf(a, a, q); }
      const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / d) + eps);
      if (lane == 0 && stats != nullptr) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
      const uint32_t rowh = (DROP && drop.thr8) ? drop_rowhash(drop.key, (uint64_t)row) : 0u;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * gm[i][j] + bt[i][j];
        if (DROP && drop.thr8) {
          float kf[8];
          drop_factors8(drop, rowh, lane + 32 * i, kf);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= kf[j];
        }
        stg_stream(reinterpret_cast<uint4*>(y + row * (int64_t)d) + lane + 32 * i, f32_to_bf16x8(o));
      }
    }
  }
}

template <int NV, bool WITH_DXSUM, bool DROP>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                    int64_t rows, const float* __restrict__ gamma,
                                                                    const float* __restrict__ stats, const __nv_bfloat16* __restrict__ add,
                                                                    __nv_bfloat16* __restrict__ dx, float* __restrict__ partial,
                                                                    const DropParams dy_drop, const DropParams dx_drop,
                                                                    __nv_bfloat16* __restrict__ dx_dropped) {
  extern __shared__ float sm[];  // [8 warps][3][d]
  constexpr int d = 256 * NV;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wstride = (int64_t)gridDim.x * 8;
  float gm[NV][8], dg[NV][8], db[NV][8], ds[WITH_DXSUM ? NV : 1][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(gamma + (lane + 32 * i) * 8), b = *reinterpret_cast<const float4*>(gamma + (lane + 32 * i) * 8 + 4);
    gm[i][0] = a.x; gm[i][1] = a.y; gm[i][2] = a.z; gm[i][3] = a.w; gm[i][4] = b.x; gm[i][5] = b.y; gm[i][6] = b.z; gm[i][7] = b.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[i][j] = 0.f; db[i][j] = 0.f; if (WITH_DXSUM) ds[i][j] = 0.f; }
  }
  // x / dy rows travel through a per-warp cp.async ring (3 stages x 2 rows x {x, dy}): the loads of the next two row pairs
  // are in flight while this pair is worked on.  With plain loads the kernel alternated between a burst of requests and a
  // stretch of arithmetic (45 % of the HBM peak, 63 % issue) -- the dgamma / dbeta / dx-sum accumulators leave no registers
  // for a software pipeline, shared memory does.  Every lane copies and later reads its OWN 16-byte chunks: no barriers.
  constexpr int LN_STAGES = 3;
  constexpr uint32_t STAGE_BYTES = 2 * 2 * d * 2;        // 2 rows x (x, dy) x d bf16
  uint8_t* ring = reinterpret_cast<uint8_t*>(sm) + (size_t)warp * LN_STAGES * STAGE_BYTES;
  const uint32_t ring_a = static_cast<uint32_t>(__cvta_generic_to_shared(ring));
  auto issue = [&](int64_t r0, int stage) {
    if (r0 < rows) {
      const int64_t r1 = r0 + wstride < rows ? r0 + wstride : r0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int64_t base = (r ? r1 : r0) * (int64_t)d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const uint32_t dst = ring_a + stage * STAGE_BYTES + ((r * 2) * NV * 32 + lane + 32 * i) * 16;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const uint4*>(x + base) + lane + 32 * i) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + NV * 32 * 16), "l"(reinterpret_cast<const uint4*>(dy + base) + lane + 32 * i) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int64_t first = (int64_t)blockIdx.x * 8 + warp;
  issue(first, 0);
  issue(first + 2 * wstride, 1);
  int stage = 0;
  for (int64_t row0 = first; row0 < rows; row0 += 2 * wstride) {
    const int64_t rr[2] = {row0, row0 + wstride};
    const bool in1 = rr[1] < rows;
    float2 st[2];
    issue(row0 + 4 * wstride, stage >= 1 ? stage - 1 : LN_STAGES - 1);     // the stage read in the previous iteration
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    // this lane's chunks of the current row pair stay in shared memory and are read once per pass (registers go to the
    // column-sum accumulators)
    uint8_t* cur = ring + stage * STAGE_BYTES + lane * 16;
    auto XV = [&](int r, int i) -> uint4& { return *reinterpret_cast<uint4*>(cur + ((r * 2) * NV * 32 + 32 * i) * 16); };
    auto DV = [&](int r, int i) -> uint4& { return *reinterpret_cast<uint4*>(cur + ((r * 2 + 1) * NV * 32 + 32 * i) * 16); };
#pragma unroll
    for (int r = 0; r < 2; ++r) st[r] = *reinterpret_cast<const float2*>(stats + 2 * ((r == 0 || in1) ? rr[r] : row0));
    stage = stage + 1 == LN_STAGES ? 0 : stage + 1;
    if (DROP && dy_drop.thr8) {                            // backward of y = dropout(LN(x)): dy <- mask * scale * dy, in registers
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint32_t rowh = drop_rowhash(dy_drop.key, (uint64_t)((r == 0 || in1) ? rr[r] : row0));
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float df[8], kf[8];
          bf16x8_to_f32(DV(r, i), df);
          drop_factors8(dy_drop, rowh, lane + 32 * i, kf);
#pragma unroll
          for (int j = 0; j < 8; ++j) df[j] *= kf[j];
          DV(r, i) = f32_to_bf16x8(df);
        }
      }
    }
    float c1[2] = {0.f, 0.f}, c2[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float xf[8], df[8];
        bf16x8_to_f32(XV(r, i), xf);
        bf16x8_to_f32(DV(r, i), df);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xf[j] - st[r].x) * st[r].y, g = df[j] * gm[i][j];
          c1[r] += g;
          c2[r] = fmaf(g, xh, c2[r]);
          if (r == 0 || in1) { dg[i][j] = fmaf(df[j], xh, dg[i][j]); db[i][j] += df[j]; }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c1[0] += __shfl_xor_sync(0xffffffffu, c1[0], o); c2[0] += __shfl_xor_sync(0xffffffffu, c2[0], o);
      c1[1] += __shfl_xor_sync(0xffffffffu, c1[1], o); c2[1] += __shfl_xor_sync(0xffffffffu, c2[1], o);
    }
    // keep words of the second output: the 4 lanes that share 32 columns need the same word for each (row, vector) pair,
    // 2 NV of them per iteration -- lane k of the group hashes pairs k, k + 4, ... and a shuffle hands them round
    // (every lane hashing all of its own words was a quarter of the kernel's instructions)
    uint32_t kwx[2][NV];
    if (DROP && dx_drop.thr8) {
      constexpr int NW = (2 * NV + 3) / 4;                 // words hashed per lane
      uint32_t mine[NW];
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        const int c = (lane & 3) + 4 * k;                  // pair index r * NV + i (c >= 2 NV: unused)
        const int r = c / NV, i = c - r * NV;
        const uint32_t rowh = drop_rowhash(dx_drop.key, (uint64_t)((r == 1 && in1) ? rr[1] : rr[0]));
        mine[k] = drop_keep_word(rowh, (uint32_t)((lane >> 2) + 8 * i), dx_drop.thr8);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = r * NV + i;
          kwx[r][i] = __shfl_sync(0xffffffffu, mine[c >> 2], (lane & ~3) | (c & 3));
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r == 1 && !in1) break;
      const float m1 = c1[r] * (1.0f / d), m2 = c2[r] * (1.0f / d);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float xf[8], df[8], o[8];
        bf16x8_to_f32(XV(r, i), xf);
        bf16x8_to_f32(DV(r, i), df);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xf[j] - st[r].x) * st[r].y;
          o[j] = st[r].y * (df[j] * gm[i][j] - m1 - xh * m2);
        }
        if (add != nullptr) {
          float af[8];
          bf16x8_to_f32(*(reinterpret_cast<const uint4*>(add + rr[r] * (int64_t)d) + lane + 32 * i), af);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += af[j];
        }
        const uint4 packed = f32_to_bf16x8(o);
        uint4 summed = packed;
        if (DROP && dx_drop.thr8) {                        // second output: the gradient behind the residual's dropout
          const uint32_t w8 = (kwx[r][i] >> ((lane & 3) * 8)) & 0xffu;      // byte (lane & 3) of keep word (lane + 32 i) >> 2
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= ((w8 >> j) & 1u) ? dx_drop.scale : 0.f;
          summed = f32_to_bf16x8(o);
          stg_stream(reinterpret_cast<uint4*>(dx_dropped + rr[r] * (int64_t)d) + lane + 32 * i, summed);
        }
        if (WITH_DXSUM) {                                  // sum what the next kernel will read (the rounded values)
          float of[8];
          bf16x8_to_f32(summed, of);
#pragma unroll
          for (int j = 0; j < 8; ++j) ds[i][j] += of[j];
        }
        stg_stream(reinterpret_cast<uint4*>(dx + rr[r] * (int64_t)d) + lane + 32 * i, packed);
      }
    }
  }
  // CTA reduction of the per-warp column sums (fixed order); the buffer aliases the rings
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  constexpr int NQ = WITH_DXSUM ? 3 : 2;
  float* mine = sm + (size_t)warp * NQ * d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[c + j] = dg[i][j];
      mine[d + c + j] = db[i][j];
      if (WITH_DXSUM) mine[2 * d + c + j] = ds[i][j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NQ * d; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[(size_t)w * NQ * d + c];
    partial[(size_t)blockIdx.x * NQ * d + c] = s;
  }
}

// out_k[c'] += sum_b partial[b][c]  for column c = k * seg + c' (k-th output pointer, may be null).
// block (32, 8): 8 row groups per column reduce through shared memory (fixed order => deterministic).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int nblocks, int ncols, int seg,
                                                              float* __restrict__ out0, float* __restrict__ out1, float* __restrict__ out2) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < ncols) {
    int b = threadIdx.y;
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (; b + 24 < nblocks; b += 32) {
      s += partial[(size_t)b * ncols + c]; s1 += partial[(size_t)(b + 8) * ncols + c];
      s2 += partial[(size_t)(b + 16) * ncols + c]; s3 += partial[(size_t)(b + 24) * ncols + c];
    }
    for (; b < nblocks; b += 8) s += partial[(size_t)b * ncols + c];
    s = (s + s1) + (s2 + s3);
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < ncols) {
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][threadIdx.x];
    const int k = c / seg, cc = c - k * seg;
    float* out = k == 0 ? out0 : (k == 1 ? out1 : out2);
    if (out) out[cc] += s;
  }
}

// ------------------------------------------------------------------ column sums (bias grads)
// block (32, 8): a warp covers 128 columns (4 per lane), 8 warps stride over the rows of this
// CTA's row chunk with 4 independent loads in flight each; partial[gridDim.y][N].
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int64_t M, int N, int64_t ldx, float* __restrict__ partial) {
  __shared__ float4 red[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const bool cin = c < N;
  const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float4 s = make_float4(0, 0, 0, 0);
  if (cin) {
    int64_t r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      const float4 a = load4(x + r * ldx + c), b = load4(x + (r + 8) * ldx + c);
      const float4 cc = load4(x + (r + 16) * ldx + c), d = load4(x + (r + 24) * ldx + c);
      s.x += (a.x + b.x) + (cc.x + d.x); s.y += (a.y + b.y) + (cc.y + d.y);
      s.z += (a.z + b.z) + (cc.z + d.z); s.w += (a.w + b.w) + (cc.w + d.w);
    }
    for (; r < r1; r += 8) {
      const float4 a = load4(x + r * ldx + c);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && cin) {
    for (int w = 1; w < 8; ++w) {
      const float4 o = red[w][threadIdx.x];
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    *reinterpret_cast<float4*>(partial + (size_t)blockIdx.y * N + c) = s;
  }
}

// ------------------------------------------------------------------ head Linear(d -> 1)
template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ x, int64_t rows, int d, const float* __restrict__ w,
                                                       const float* __restrict__ b, const float* __restrict__ add, float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const int nvec = d >> 2;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    const T* xr = x + row * (int64_t)d;
    float s = 0.f;
    for (int c = lane; c < nvec; c += 32) {
      const float4 v = load4(xr + c * 4);
      const float4 ww = *reinterpret_cast<const float4*>(w + c * 4);
      s += v.x * ww.x + v.y * ww.y + v.z * ww.z + v.w * ww.w;
    }
    s = warp_sum(s);
    if (lane == 0) logits[row] = s + (b ? b[0] : 0.f) + (add ? add[row] : 0.f);
  }
}

// partial layout [gridDim.x][d + 1]: dw then db
template <typename T, int NV>
__global__ void __launch_bounds__(256) head_bwd_kernel(const T* __restrict__ x, int64_t rows, int d, const float* __restrict__ w,
                                                       const float* __restrict__ dlogits, const float* __restrict__ gscale,
                                                       T* __restrict__ dx, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [8][d+1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = d >> 2;
  const float gs = gscale ? gscale[0] : 1.0f;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  float4 dw[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dw[i] = make_float4(0, 0, 0, 0);
  float dbias = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; row < rows; row += wstride) {
    const float g = dlogits[row] * gs;
    dbias += g;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 xv = load4(x + row * (int64_t)d + c * 4);
        const float4 ww = *reinterpret_cast<const float4*>(w + c * 4);
        dw[i].x += g * xv.x; dw[i].y += g * xv.y; dw[i].z += g * xv.z; dw[i].w += g * xv.w;
        store4(dx + row * (int64_t)d + c * 4, make_float4(g * ww.x, g * ww.y, g * ww.z, g * ww.w));
      }
    }
  }
  float* mine = sm + (size_t)warp * (d + 1);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      mine[c * 4 + 0] = dw[i].x; mine[c * 4 + 1] = dw[i].y; mine[c * 4 + 2] = dw[i].z; mine[c * 4 + 3] = dw[i].w;
    }
  }
  if (lane == 0) mine[d] = dbias;  // every lane carries the same dbias
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < d + 1; c += blockDim.x) {
    float s = 0.f;
    for (int wi = 0; wi < nw; ++wi) s += sm[(size_t)wi * (d + 1) + c];
    partial[(size_t)blockIdx.x * (d + 1) + c] = s;
  }
}

// ------------------------------------------------------------------ loss (focal / interestBPR) + diagnostics
// One CTA (<= 32 warps), one warp per row, L <= 64.  models/decoder_leave_focal.py:490-572:
//   logits = stage_logits (+ (pos+1) * bias_weight + bias_bias)                      :497-504
//   focal       : my_sigmoid_focal_loss (alpha .5, gamma 2), masked sum / bsz        :35-59, 533-538
//   interestBPR : compute_interest_BPR_all, rows with view_len < L, mean over rows   :163-221
//   mse / mse2 diagnostics (incl. the [B] vs [B,1] broadcast of the reference)        :552-558
//   huber       : huber_loss(sum hazard_masked [B], view_lengths [B,1]) -> [B,B] mean  :61-66, 539-540
//   hazard      : compute_partial_likelihood_loss (rows with view_len == L skipped)   :273-286
//   surviveCE   : BCE-with-logits fed exp(h_t) as the logit, masked batch mean        :68-97
//   interestCE / interestKL : softmax(logits) vs softmax(gt != 0), optional mask      :99-161
// The three survival-chain losses share one gradient path: c_t = d loss / d S_t, then
//   d loss / d x_l = (1 - sigmoid(x_l)) * sum_{t >= l} c_t S_t     (S_t = exp(cumsum log sigmoid)).
// dlogits = d loss / d logits for loss = sum over the losses switched on of weight * value.
struct LossParams {
  const float* logits_in; int64_t* gt; int B, L;
  const float* ep; const float* bias_w; const float* bias_b;
  float inv_bsz, w_focal, w_bpr, bpr_scale;
  int use_focal, use_bpr, rewrite_gt;
  float* logits_out; float* scalars; float* dlogits; float* dbias_w; float* dbias_b;
  int use_huber, use_hazard, use_sce, use_ice, use_ikl, mask_loss, ce_after_focal, kl_after_focal;
  float w_huber, w_hazard, w_sce, w_ice, w_ikl;
};

// huber(e), delta 1 (models/decoder_leave_focal.py:61-66) and its derivative
__device__ __forceinline__ float huber1(float e) { const float a = fabsf(e); return a < 1.f ? 0.5f * e * e : a - 0.5f; }
__device__ __forceinline__ float huber1_grad(float e) { return fabsf(e) < 1.f ? e : (e > 0.f ? 1.f : -1.f); }

__global__ void __launch_bounds__(1024) loss_kernel(const LossParams q) {
  constexpr int NACC = 15;
  __shared__ double red[32][NACC];
  __shared__ int n_bpr_rows_s, n_valid_total_s;
  __shared__ int view_hist_s[66];                        // rows per view length (huber's [B,B] broadcast)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int B = q.B, L = q.L;
  // ---- batch-wide counts that the means need before any gradient: rows in interestBPR (view_len < L), valid
  //      positions (surviveCE), rows per view length (huber)
  if (threadIdx.x == 0) { n_bpr_rows_s = 0; n_valid_total_s = 0; }
  for (int i = threadIdx.x; i < 66; i += blockDim.x) view_hist_s[i] = 0;
  __syncthreads();
  if (q.use_bpr || q.use_sce || q.use_huber) {
    int cnt = 0, nval = 0;
    for (int row = warp; row < B; row += nw) {
      const long long g0 = lane < L ? q.gt[(size_t)row * L + lane] : -2;
      const long long g1 = lane + 32 < L ? q.gt[(size_t)row * L + lane + 32] : -2;
      const int n_view = __popc(__ballot_sync(0xffffffffu, g0 == 1)) + __popc(__ballot_sync(0xffffffffu, g1 == 1));
      nval += __popc(__ballot_sync(0xffffffffu, g0 != -2)) + __popc(__ballot_sync(0xffffffffu, g1 != -2));
      cnt += n_view < L ? 1 : 0;
      if (lane == 0 && q.use_huber) atomicAdd(&view_hist_s[n_view], 1);
    }
    if (lane == 0 && cnt) atomicAdd(&n_bpr_rows_s, cnt);
    if (lane == 0 && nval) atomicAdd(&n_valid_total_s, nval);
  }
  __syncthreads();
  const int n_bpr = n_bpr_rows_s;
  const float inv_nvalid_total = 1.0f / (float)n_valid_total_s;
  const float inv_nbpr = 1.0f / (float)n_bpr;            // 0 rows: inf -> loss = 0 * inf = NaN like torch's empty mean
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
  for (int row = warp; row < B; row += nw) {
    float x[2], sig[2], logp[2];
    long long g[2];
    bool in[2], valid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int l = lane + 32 * h;
      in[h] = l < L;
      x[h] = in[h] ? q.logits_in[(size_t)row * L + l] : 0.f;
      if (in[h] && q.bias_w != nullptr) x[h] += (float)(l + 1) * q.bias_w[l] + q.bias_b[l];
      if (in[h] && q.logits_out != nullptr) q.logits_out[(size_t)row * L + l] = x[h];
      g[h] = in[h] ? q.gt[(size_t)row * L + l] : -2;
      valid[h] = in[h] && g[h] != -2;
      sig[h] = 1.0f / (1.0f + expf(-x[h]));
      logp[h] = in[h] ? logf(sig[h]) : 0.f;
    }
    // inclusive scan of log p over the row: h_t = cumsum(log sigmoid(x))  (:506-510)
    float sc0 = logp[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, sc0, o);
      if (lane >= o) sc0 += n;
    }
    const float tot0 = __shfl_sync(0xffffffffu, sc0, 31);
    float sc1 = logp[1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, sc1, o);
      if (lane >= o) sc1 += n;
    }
    sc1 += tot0;
    const float surv[2] = {expf(sc0), expf(sc1)};
    const float sm0 = valid[0] ? surv[0] : 0.f, sm1 = valid[1] ? surv[1] : 0.f;
    const float s_row = warp_sum(sm0 + sm1);
    const int n_valid = __popc(__ballot_sync(0xffffffffu, valid[0])) + __popc(__ballot_sync(0xffffffffu, valid[1]));
    const int n_view = __popc(__ballot_sync(0xffffffffu, in[0] && g[0] == 1)) + __popc(__ballot_sync(0xffffffffu, in[1] && g[1] == 1));
    const int n_nonneg_orig = __popc(__ballot_sync(0xffffffffu, in[0] && g[0] >= 0)) + __popc(__ballot_sync(0xffffffffu, in[1] && g[1] >= 0));
    // mse2: survival_masked[i, durations[i]-1] = 1 (python index -1 wraps to L-1)  (:554-555)
    const int pos = (n_valid - 1 + L) % L;
    const float at_pos = (pos < 32) ? __shfl_sync(0xffffffffu, sm0, pos) : __shfl_sync(0xffffffffu, sm1, pos - 32);
    const float s2_row = s_row - at_pos + 1.0f;
    // after the in-place rewrite gt in {1,0,-2}: (gt>=0).sum() == n_valid; otherwise count of {1,0}
    const float v2 = (q.use_focal && q.rewrite_gt) ? (float)n_valid : (float)n_nonneg_orig;
    float dl[2] = {0.f, 0.f};
    float lsum = 0.f;
    if (q.use_focal) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int l = lane + 32 * h;
        if (!in[h]) continue;
        const long long gn = (g[h] > 0) ? 1 : (g[h] == -1 ? 0 : g[h]);
        if (q.rewrite_gt) q.gt[(size_t)row * L + l] = gn;
        if (valid[h]) {
          const float t = (float)gn;
          const float e = q.ep[l];
          const float p = sig[h] * e;
          const float ce = fmaxf(x[h], 0.f) - x[h] * t + log1pf(expf(-fabsf(x[h])));
          const float pt = p * t + (1.f - p) * (1.f - t);
          const float om = 1.f - pt;
          lsum += 0.5f * ce * om * om;
          const float dp = e * sig[h] * (1.f - sig[h]);
          dl[h] = 0.5f * ((sig[h] - t) * om * om - ce * 2.f * om * (2.f * t - 1.f) * dp) * q.w_focal * q.inv_bsz;
        }
      }
      lsum = warp_sum(lsum);
    }
    // ---- interestBPR (:163-221): pos = logits[row, view_len]; the other L-1 logits (pad positions included) are negatives
    float bpr_row = 0.f;
    if (q.use_bpr && n_view < L) {                       // warp-uniform
      const float xpos = (n_view < 32) ? __shfl_sync(0xffffffffu, x[0], n_view) : __shfl_sync(0xffffffffu, x[1], n_view - 32);
      bool neg[2];
      float mx = -INFINITY;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        neg[h] = in[h] && (lane + 32 * h) != n_view;
        if (neg[h]) mx = fmaxf(mx, x[h]);
      }
      mx = warp_max(mx);
      float e[2], gsig[2], se = 0.f, sa = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        e[h] = neg[h] ? expf(x[h] - mx) : 0.f;
        gsig[h] = neg[h] ? 1.0f / (1.0f + expf(-(x[h] - xpos))) : 0.f;
        se += e[h];
        sa += e[h] * gsig[h];
      }
      se = warp_sum(se);
      sa = warp_sum(sa);
      const float A = sa / se;                           // sum_k softmax_k * sigmoid(neg_k - pos)
      const float Ac = fminf(fmaxf(A, 1e-8f), 1.0f - 1e-8f);
      bpr_row = -logf(Ac);
      const float dA = (A > 1e-8f && A < 1.0f - 1e-8f) ? -1.0f / A : 0.f;   // clamp passes no gradient outside its range
      const float coef = dA * q.w_bpr * q.bpr_scale * inv_nbpr;
      float dpos = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (neg[h]) {
          const float sk = e[h] / se, gg = gsig[h] * (1.0f - gsig[h]);
          dl[h] += coef * (sk * gg + sk * (gsig[h] - A));
          dpos -= sk * gg;
        }
      }
      dpos = warp_sum(dpos) * coef;
      if (n_view < 32) { if (lane == n_view) dl[0] += dpos; }
      else if (lane == n_view - 32) dl[1] += dpos;
    }
    // ---- survival-chain losses: c[h] = d loss / d S_t at this lane's positions (zero where gt == -2)
    float c[2] = {0.f, 0.f};
    float huber_row = 0.f, hazard_row = 0.f, sce_row = 0.f;
    if (q.use_huber) {                                   // mean_ij huber(H_j - v_i): row j against the histogram of v
      const float H = warp_sum((valid[0] ? 1.f - surv[0] : 0.f) + (valid[1] ? 1.f - surv[1] : 0.f));
      float hv = 0.f, hg = 0.f;
      for (int v = lane; v <= L; v += 32) {
        const float n = (float)view_hist_s[v];
        hv += n * huber1(H - (float)v);
        hg += n * huber1_grad(H - (float)v);
      }
      huber_row = warp_sum(hv);
      const float gH = warp_sum(hg) * q.w_huber * q.bpr_scale / ((float)B * (float)B);
#pragma unroll
      for (int h = 0; h < 2; ++h) if (valid[h]) c[h] -= gH;          // H = sum_valid (1 - S_t)
    }
    if (q.use_hazard && n_view < L) {                    // reference: `if observed_time==40: continue` (L = 40)
      const float hm[2] = {valid[0] ? 1.f - surv[0] : 0.f, valid[1] ? 1.f - surv[1] : 0.f};
      const float at = (n_view < 32) ? __shfl_sync(0xffffffffu, hm[0], n_view) : __shfl_sync(0xffffffffu, hm[1], n_view - 32);
      const float risk = warp_sum((lane >= n_view ? hm[0] : 0.f) + (lane + 32 >= n_view ? hm[1] : 0.f));
      hazard_row = -(logf(at + 1e-6f) - logf(risk + 1e-6f));
      const float coef = q.w_hazard * q.inv_bsz;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int l = lane + 32 * h;
        if (valid[h]) c[h] += coef * ((l == n_view ? 1.f / (at + 1e-6f) : 0.f) - (l >= n_view ? 1.f / (risk + 1e-6f) : 0.f));
      }
    }
    if (q.use_sce) {
      const float coef = q.w_sce * q.bpr_scale * inv_nvalid_total;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!valid[h]) continue;
        const float S = surv[h], y = (g[h] == 1) ? 1.f : 0.f;
        sce_row += fmaxf(S, 0.f) - S * y + log1pf(expf(-fabsf(S)));
        c[h] += coef * (1.0f / (1.0f + expf(-S)) - y);
      }
      sce_row = warp_sum(sce_row);
    }
    if (q.use_huber || q.use_hazard || q.use_sce) {      // suffix sums of c_t S_t over t >= l
      float u0 = c[0] * surv[0], u1 = c[1] * surv[1];
      if (!in[0]) u0 = 0.f;
      if (!in[1]) u1 = 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float a0 = __shfl_down_sync(0xffffffffu, u0, o), a1 = __shfl_down_sync(0xffffffffu, u1, o);
        if (lane + o < 32) { u0 += a0; u1 += a1; }
      }
      u0 += __shfl_sync(0xffffffffu, u1, 0);
      dl[0] += (1.f - sig[0]) * u0;
      dl[1] += (1.f - sig[1]) * u1;
    }
    // ---- interestCE / interestKL (:99-161): q = softmax(x) over all L positions, target = softmax(gt != 0)
    float ice_row = 0.f, ikl_row = 0.f;
    if (q.use_ice || q.use_ikl) {
      float mx = warp_max(fmaxf(in[0] ? x[0] : -INFINITY, in[1] ? x[1] : -INFINITY));
      float e[2] = {in[0] ? expf(x[0] - mx) : 0.f, in[1] ? expf(x[1] - mx) : 0.f};
      const float se = warp_sum(e[0] + e[1]);
      float qs[2], logq[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) { qs[h] = e[h] / se; logq[h] = in[h] ? logf(qs[h]) : 0.f; }   // log(softmax), as written
      const float inv_nvalid_row = 1.0f / (float)n_valid;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        if (!(which == 0 ? q.use_ice : q.use_ikl)) continue;
        const bool rewritten = which == 0 ? q.ce_after_focal : q.kl_after_focal;   // focal earlier in the list rewrote gt
        bool nl[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) nl[h] = in[h] && (rewritten ? (g[h] > 0 || (g[h] < 0 && g[h] != -1)) : (g[h] != 0));
        const int n1 = __popc(__ballot_sync(0xffffffffu, nl[0])) + __popc(__ballot_sync(0xffffffffu, nl[1]));
        const float t_hi = n1 > 0 ? 1.f : 0.f;                                       // softmax subtracts the row max first
        const float Z = (float)n1 * expf(1.f - t_hi) + (float)(L - n1) * expf(0.f - t_hi);
        float w[2], W = 0.f, A = 0.f, T = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float t = in[h] ? expf((nl[h] ? 1.f : 0.f) - t_hi) / Z : 0.f;
          w[h] = q.mask_loss ? (valid[h] ? t * inv_nvalid_row : 0.f) : t;
          W += w[h];
          A += w[h] * logq[h];
          T += in[h] ? w[h] * logf(t) : 0.f;
        }
        W = warp_sum(W); A = warp_sum(A); T = warp_sum(T);
        const float coef = (which == 0 ? q.w_ice : q.w_ikl) * q.inv_bsz;
        if (which == 0) ice_row = -A; else ikl_row = T - A;
#pragma unroll
        for (int h = 0; h < 2; ++h) if (in[h]) dl[h] += coef * (W * qs[h] - w[h]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (in[h]) q.dlogits[(size_t)row * L + lane + 32 * h] = dl[h];
    if (lane == 0) {
      acc[10] += huber_row; acc[11] += hazard_row; acc[12] += sce_row; acc[13] += ice_row; acc[14] += ikl_row;
      acc[0] += lsum;
      acc[1] += s_row; acc[2] += (double)s_row * s_row;
      acc[3] += n_view; acc[4] += (double)n_view * n_view;
      acc[5] += s2_row; acc[6] += (double)s2_row * s2_row;
      acc[7] += v2; acc[8] += (double)v2 * v2;
      acc[9] += bpr_row;
    }
  }
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NACC; ++i) red[warp][i] = acc[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[NACC];
    for (int i = 0; i < NACC; ++i) {
      t[i] = 0.0;
      for (int w = 0; w < nw; ++w) t[i] += red[w][i];
    }
    const double b = (double)B;
    const double focal = t[0] * (double)q.inv_bsz;
    const double bpr = q.use_bpr ? t[9] * (double)q.bpr_scale / (double)n_bpr : 0.0;
    // nn.MSELoss()([B], [B,1]) broadcasts to [B,B]: mean_ij (s_j - v_i)^2  (:552)
    const double mse = t[2] / b - 2.0 * (t[1] / b) * (t[3] / b) + t[4] / b;
    const double mse2 = t[6] / b - 2.0 * (t[5] / b) * (t[7] / b) + t[8] / b;
    q.scalars[0] = (float)focal;
    q.scalars[1] = (float)mse;
    q.scalars[2] = (float)mse2;
    const double huber = q.use_huber ? t[10] * (double)q.bpr_scale / (b * b) : 0.0;
    const double hazard = q.use_hazard ? t[11] * (double)q.inv_bsz : 0.0;
    const double sce = q.use_sce ? t[12] * (double)q.bpr_scale / (double)n_valid_total_s : 0.0;
    const double ice = q.use_ice ? t[13] * (double)q.inv_bsz : 0.0;
    const double ikl = q.use_ikl ? t[14] * (double)q.inv_bsz : 0.0;
    q.scalars[3] = (float)((q.use_focal ? focal * (double)q.w_focal : 0.0) + (q.use_bpr ? bpr * (double)q.w_bpr : 0.0) +
                           huber * (double)q.w_huber + hazard * (double)q.w_hazard + sce * (double)q.w_sce +
                           ice * (double)q.w_ice + ikl * (double)q.w_ikl);
    q.scalars[4] = (float)bpr;
    q.scalars[5] = (float)huber; q.scalars[6] = (float)hazard; q.scalars[7] = (float)sce;
    q.scalars[8] = (float)ice; q.scalars[9] = (float)ikl;
  }
  // ---- learnable position bias gradients: d bias_bias[l] = sum_b dlogits[b,l], d bias_weight[l] = (l+1) * that
  if (q.dbias_w != nullptr || q.dbias_b != nullptr) {
    __threadfence_block();
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < B; ++r) s += q.dlogits[(size_t)r * L + l];
      if (q.dbias_b) q.dbias_b[l] += s;
      if (q.dbias_w) q.dbias_w[l] += s * (float)(l + 1);
    }
  }
}

// ------------------------------------------------------------------ clip + AdamW
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partial) {
  __shared__ double red[8];
  float s = 0.f;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  int iter = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(g + i * 4);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    if (++iter == 64) { acc += s; s = 0.f; iter = 0; }
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += g[i] * g[i];
  acc += s;
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void clip_coef_kernel(const double* __restrict__ partial, int n, float max_norm, float* __restrict__ norm_out) {
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) t += partial[i];
  t = warp_sum_d(t);
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(t);
    float coef = max_norm / (norm + 1e-6f);  // torch.nn.utils.clip_grad_norm_
    coef = coef > 1.0f ? 1.0f : coef;
    norm_out[0] = norm;
    norm_out[1] = coef;
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float lr_wd, float beta1, float beta2,
                                                    float eps, float step_size, float bc2_sqrt, const float* __restrict__ norm_out,
                                                    __nv_bfloat16* __restrict__ bf16_out) {
  const float coef = norm_out ? norm_out[1] : 1.0f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.0f - lr_wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (bf16_out) bf16_out[i] = __float2bfloat16_rn(pi);
  }
}

// ------------------------------------------------------------------ cast
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void cast_bf16_transpose_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows, int64_t cols) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[c * rows + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

static int grid_for_rows(int64_t rows, int warps_per_cta, int cap) {
  int64_t ctas = (rows + warps_per_cta - 1) / warps_per_cta;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return (int)ctas;
}


// ------------------------------------------------------------------ dropout test hook
__global__ void dropout_mask_kernel(const DropParams dp, int64_t row0, int64_t rows, int cols, uint32_t group0, uint8_t* __restrict__ mask) {
  const int64_t n = rows * (int64_t)cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    const uint32_t w = dp.thr8 ? drop_keep_word(drop_rowhash(dp.key, (uint64_t)(row0 + r)), group0 + ((uint32_t)c >> 5), dp.thr8) : 0xffffffffu;
    mask[i] = (w >> (c & 31)) & 1u;
  }
}

}  // namespace mmi

using namespace mmi;

extern "C" int mmi_dropout_mask(const mmi_dropout* drop, int64_t row0, int64_t rows, int cols, uint32_t group0, uint8_t* mask,
                                mmi_stream_t stream) {
  MMI_CHECK_ARG(drop && mask && rows >= 0 && cols > 0, "dropout_mask: bad arguments");
  MMI_CHECK_ARG(drop->thr8 < 256u, "dropout_mask: thr8 must be < 256");
  if (rows == 0) return MMI_OK;
  dropout_mask_kernel<<<kNumSMs * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(make_drop(*drop), row0, rows, cols, group0, mask);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_layernorm_fwd(const void* x, int dtype, int64_t rows, int d, const float* gamma, const float* beta, float eps,
                                 void* y, float* stats, mmi_stream_t stream) {
  return mmi_layernorm_fwd_drop(x, dtype, rows, d, gamma, beta, eps, y, stats, nullptr, stream);
}

extern "C" int mmi_layernorm_fwd_drop(const void* x, int dtype, int64_t rows, int d, const float* gamma, const float* beta, float eps,
                                      void* y, float* stats, const mmi_dropout* drop, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const DropParams dp = drop ? make_drop(*drop) : drop_off();
  MMI_CHECK_ARG(dp.thr8 < 256u, "layernorm_fwd: dropout thr8 must be < 256");
  MMI_CHECK_ARG(x && y && gamma && beta, "layernorm_fwd: null pointer");
  MMI_CHECK_ARG(d % 4 == 0 && d <= 128 * kMaxVecPerLane && d > 0, "layernorm: d=%d must be a multiple of 4 and <= %d", d, 128 * kMaxVecPerLane);
  if (rows == 0) return MMI_OK;
  if (dtype == MMI_BF16 && (d == 256 || d == 512 || d == 768 || d == 1024) &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0) {
    // bandwidth path: 16-byte vectors through a cp.async ring, two rows per warp and stage
    const int grid16 = grid_for_rows((rows + 1) / 2, 8, kNumSMs * 4);
    const size_t smem = (size_t)8 * 3 * 2 * d * 2;
#define MMI_LN_FWD16(NV_)                                                                                                     \
  do {                                                                                                                        \
    if (dp.thr8) {                                                                                                            \
      if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_fwd_bf16_kernel<NV_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      layernorm_fwd_bf16_kernel<NV_, true><<<grid16, 256, smem, st>>>((const __nv_bfloat16*)x, rows, gamma, beta, eps, (__nv_bfloat16*)y, stats, dp); \
    } else {                                                                                                                  \
      if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_fwd_bf16_kernel<NV_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      layernorm_fwd_bf16_kernel<NV_, false><<<grid16, 256, smem, st>>>((const __nv_bfloat16*)x, rows, gamma, beta, eps, (__nv_bfloat16*)y, stats, dp); \
    }                                                                                                                         \
  } while (0)
    if (d == 256) MMI_LN_FWD16(1); else if (d == 512) MMI_LN_FWD16(2); else if (d == 768) MMI_LN_FWD16(3); else MMI_LN_FWD16(4);
#undef MMI_LN_FWD16
    MMI_CHECK_LAUNCH();
    return MMI_OK;
  }
  const int grid = grid_for_rows(rows, 8, kNumSMs * 8);
  const int nv = d <= 128 ? 1 : (d <= 256 ? 2 : (d <= 512 ? 4 : 8));
#define MMI_LN_FWD(T_, NV_) layernorm_fwd_kernel<T_, NV_><<<grid, 256, 0, st>>>((const T_*)x, rows, d, gamma, beta, eps, (T_*)y, stats, dp)
#define MMI_LN_FWD_NV(T_) do { if (nv == 1) MMI_LN_FWD(T_, 1); else if (nv == 2) MMI_LN_FWD(T_, 2); else if (nv == 4) MMI_LN_FWD(T_, 4); else MMI_LN_FWD(T_, 8); } while (0)
  if (dtype == MMI_F32) MMI_LN_FWD_NV(float);
  else if (dtype == MMI_BF16) MMI_LN_FWD_NV(__nv_bfloat16);
  else { set_error("layernorm_fwd: bad dtype %d", dtype); return MMI_EINVAL; }
#undef MMI_LN_FWD_NV
#undef MMI_LN_FWD
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int64_t mmi_layernorm_bwd_workspace(int d) { return (int64_t)kRedCtas * 3 * d; }

extern "C" int mmi_layernorm_bwd(const void* dy, const void* x, int dtype, int64_t rows, int d, const float* gamma, const float* stats,
                                 const void* add, void* dx, float* dgamma, float* dbeta, float* dxsum, float* workspace,
                                 mmi_stream_t stream) {
  return mmi_layernorm_bwd_drop(dy, x, dtype, rows, d, gamma, stats, add, dx, dgamma, dbeta, dxsum, workspace, nullptr, nullptr, nullptr, stream);
}

extern "C" int mmi_layernorm_bwd_drop(const void* dy, const void* x, int dtype, int64_t rows, int d, const float* gamma,
                                      const float* stats, const void* add, void* dx, float* dgamma, float* dbeta, float* dxsum,
                                      float* workspace, const mmi_dropout* dy_drop, const mmi_dropout* dx_drop, void* dx_dropped,
                                      mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(dy && x && gamma && stats && dx && workspace, "layernorm_bwd: null pointer");
  const DropParams dyd = dy_drop ? make_drop(*dy_drop) : drop_off();
  DropParams dxd = dx_drop ? make_drop(*dx_drop) : drop_off();
  MMI_CHECK_ARG(dyd.thr8 < 256u && dxd.thr8 < 256u, "layernorm_bwd: dropout thr8 must be < 256");
  MMI_CHECK_ARG(!(dx_drop && !dx_dropped), "layernorm_bwd: dx_drop needs the dx_dropped output");
  if (dx_drop && dxd.thr8 == 0u) {   // site switched off: the dropped gradient IS dx (keep the two-output contract with thr8 = 1 .. 255 only)
    set_error("layernorm_bwd: dx_drop with thr8 = 0; pass NULL and use dx");
    return MMI_EINVAL;
  }
  const bool any_drop = dyd.thr8 != 0u || dxd.thr8 != 0u;
  MMI_CHECK_ARG(d % 4 == 0 && d <= 128 * kMaxVecPerLane && d > 0, "layernorm: d=%d must be a multiple of 4 and <= %d", d, 128 * kMaxVecPerLane);
  if (rows == 0) return MMI_OK;
  const bool al16 = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
                      reinterpret_cast<uintptr_t>(add) | reinterpret_cast<uintptr_t>(dx_dropped)) & 15) == 0;
  if (dtype == MMI_BF16 && (d == 256 || d == 512 || d == 768 || d == 1024) && al16) {
    // bandwidth path: 16-byte vectors, two rows per warp in flight, dx column sums fused
    const int grid = grid_for_rows((rows + 1) / 2, 8, kNumSMs * 2);
    const int nq = dxsum ? 3 : 2;
    const size_t ring = (size_t)8 * 3 * 2 * 2 * d * 2;          // 8 warps x 3 stages x 2 rows x (x, dy) x d bf16
    const size_t red = (size_t)8 * nq * d * sizeof(float);
    const size_t smem = ring > red ? ring : red;
#define MMI_LN_BWD16_K(NV_, DS_, DR_)                                                                                          \
  do {                                                                                                                        \
    if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_bwd_bf16_kernel<NV_, DS_, DR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    layernorm_bwd_bf16_kernel<NV_, DS_, DR_><<<grid, 256, smem, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, rows, gamma, stats, (const __nv_bfloat16*)add, (__nv_bfloat16*)dx, workspace, dyd, dxd, (__nv_bfloat16*)dx_dropped); \
  } while (0)
#define MMI_LN_BWD16(NV_)                                                                                                     \
  do {                                                                                                                        \
    if (dxsum) { if (any_drop) MMI_LN_BWD16_K(NV_, true, true); else MMI_LN_BWD16_K(NV_, true, false); }                      \
    else { if (any_drop) MMI_LN_BWD16_K(NV_, false, true); else MMI_LN_BWD16_K(NV_, false, false); }                          \
  } while (0)
    if (d == 256) MMI_LN_BWD16(1); else if (d == 512) MMI_LN_BWD16(2); else if (d == 768) MMI_LN_BWD16(3); else MMI_LN_BWD16(4);
#undef MMI_LN_BWD16_K
#undef MMI_LN_BWD16
    MMI_CHECK_LAUNCH();
    reduce_partials_kernel<<<(nq * d + 31) / 32, dim3(32, 8), 0, st>>>(workspace, grid, nq * d, d, dgamma, dbeta, dxsum);
    MMI_CHECK_LAUNCH();
    return MMI_OK;
  }
  const int grid = grid_for_rows(rows, 8, kRedCtas);
  const size_t smem = (size_t)8 * 2 * d * sizeof(float);
  const int nv = d <= 128 ? 1 : (d <= 256 ? 2 : (d <= 512 ? 4 : 8));
#define MMI_LN_BWD(T_, NV_)                                                                                                   \
  do {                                                                                                                        \
    if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_bwd_kernel<T_, NV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    layernorm_bwd_kernel<T_, NV_><<<grid, 256, smem, st>>>((const T_*)dy, (const T_*)x, rows, d, gamma, stats, (const T_*)add, (T_*)dx, workspace, dyd, dxd, (T_*)dx_dropped); \
  } while (0)
#define MMI_LN_BWD_NV(T_) do { if (nv == 1) MMI_LN_BWD(T_, 1); else if (nv == 2) MMI_LN_BWD(T_, 2); else if (nv == 4) MMI_LN_BWD(T_, 4); else MMI_LN_BWD(T_, 8); } while (0)
  if (dtype == MMI_F32) MMI_LN_BWD_NV(float);
  else if (dtype == MMI_BF16) MMI_LN_BWD_NV(__nv_bfloat16);
  else { set_error("layernorm_bwd: bad dtype %d", dtype); return MMI_EINVAL; }
#undef MMI_LN_BWD_NV
#undef MMI_LN_BWD
  MMI_CHECK_LAUNCH();
  reduce_partials_kernel<<<(2 * d + 31) / 32, dim3(32, 8), 0, st>>>(workspace, grid, 2 * d, d, dgamma, dbeta, nullptr);
  MMI_CHECK_LAUNCH();
  if (dxsum) {   // generic path: separate column-sum pass over dx
    const int rc = mmi_colsum_acc(dxd.thr8 ? dx_dropped : dx, dtype, rows, d, d, dxsum, workspace, mmi_layernorm_bwd_workspace(d), stream);
    if (rc) return rc;
  }
  return MMI_OK;
}

extern "C" int mmi_colsum_acc(const void* x, int dtype, int64_t M, int N, int64_t ldx, float* out, float* workspace,
                              int64_t workspace_floats, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(x && out && workspace, "colsum: null pointer");
  MMI_CHECK_ARG(N % 4 == 0 && N > 0, "colsum: N=%d must be a multiple of 4", N);
  if (M == 0) return MMI_OK;
  const int gx = (N / 4 + 31) / 32;
  int64_t gy = (4 * kNumSMs + gx - 1) / gx;
  if (gy > (M + 63) / 64) gy = (M + 63) / 64;
  if (gy * N > workspace_floats) gy = workspace_floats / N;
  MMI_CHECK_ARG(gy >= 1, "colsum: workspace too small (%lld floats for N=%d)", (long long)workspace_floats, N);
  dim3 grid(gx, (unsigned)gy);
  dim3 block(32, 8);
  if (dtype == MMI_F32) colsum_kernel<float><<<grid, block, 0, st>>>((const float*)x, M, N, ldx, workspace);
  else if (dtype == MMI_BF16) colsum_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)x, M, N, ldx, workspace);
  else { set_error("colsum: bad dtype %d", dtype); return MMI_EINVAL; }
  MMI_CHECK_LAUNCH();
  reduce_partials_kernel<<<(N + 31) / 32, dim3(32, 8), 0, st>>>(workspace, (int)gy, N, N, out, nullptr, nullptr);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_head_fwd(const void* x, int dtype, int64_t rows, int d, const float* w, const float* b, const float* add,
                            float* logits, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(x && w && logits, "head_fwd: null pointer");
  MMI_CHECK_ARG(d % 4 == 0 && d > 0, "head: d=%d must be a multiple of 4", d);
  if (rows == 0) return MMI_OK;
  const int grid = grid_for_rows(rows, 8, kNumSMs * 8);
  if (dtype == MMI_F32) head_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, rows, d, w, b, add, logits);
  else if (dtype == MMI_BF16) head_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, rows, d, w, b, add, logits);
  else { set_error("head_fwd: bad dtype %d", dtype); return MMI_EINVAL; }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int64_t mmi_head_bwd_workspace(int d) { return (int64_t)kRedCtas * (d + 1); }

extern "C" int mmi_head_bwd(const void* x, int dtype, int64_t rows, int d, const float* w, const float* dlogits, const float* gscale,
                            void* dx, float* dw, float* db, float* workspace, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(x && w && dlogits && dx && dw && workspace, "head_bwd: null pointer");   // db may be NULL (bias owned by another head)
  MMI_CHECK_ARG(d % 4 == 0 && d <= 128 * kMaxVecPerLane && d > 0, "head: d=%d must be a multiple of 4 and <= %d", d, 128 * kMaxVecPerLane);
  if (rows == 0) return MMI_OK;
  const int grid = grid_for_rows(rows, 8, kRedCtas);
  const size_t smem = (size_t)8 * (d + 1) * sizeof(float);
  if (dtype == MMI_F32) {
    if (d <= 512) head_bwd_kernel<float, 4><<<grid, 256, smem, st>>>((const float*)x, rows, d, w, dlogits, gscale, (float*)dx, workspace);
    else head_bwd_kernel<float, 8><<<grid, 256, smem, st>>>((const float*)x, rows, d, w, dlogits, gscale, (float*)dx, workspace);
  } else if (dtype == MMI_BF16) {
    if (d <= 512) head_bwd_kernel<__nv_bfloat16, 4><<<grid, 256, smem, st>>>((const __nv_bfloat16*)x, rows, d, w, dlogits, gscale, (__nv_bfloat16*)dx, workspace);
    else head_bwd_kernel<__nv_bfloat16, 8><<<grid, 256, smem, st>>>((const __nv_bfloat16*)x, rows, d, w, dlogits, gscale, (__nv_bfloat16*)dx, workspace);
  } else { set_error("head_bwd: bad dtype %d", dtype); return MMI_EINVAL; }
  MMI_CHECK_LAUNCH();
  reduce_partials_kernel<<<(d + 1 + 31) / 32, dim3(32, 8), 0, st>>>(workspace, grid, d + 1, d, dw, db, nullptr);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_loss_fwd_bwd(const mmi_loss_args* a, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(a && a->logits && a->gt && a->scalars && a->dlogits, "loss: null pointer");
  MMI_CHECK_ARG(a->L > 0 && a->L <= 64 && a->B > 0, "loss: need 0 < L <= 64 (got %d), B > 0 (got %d)", a->L, a->B);
  MMI_CHECK_ARG(!a->use_focal || a->exposure_prob, "loss: focal needs exposure_prob");
  MMI_CHECK_ARG((a->bias_weight == nullptr) == (a->bias_bias == nullptr), "loss: bias_weight and bias_bias go together");
  LossParams q;
  q.logits_in = a->logits; q.gt = a->gt; q.B = a->B; q.L = a->L;
  q.ep = a->exposure_prob; q.bias_w = a->bias_weight; q.bias_b = a->bias_bias;
  q.inv_bsz = a->inv_bsz; q.w_focal = a->w_focal; q.w_bpr = a->w_bpr; q.bpr_scale = a->bpr_scale;
  q.use_focal = a->use_focal; q.use_bpr = a->use_bpr; q.rewrite_gt = a->rewrite_gt;
  q.logits_out = a->logits_out; q.scalars = a->scalars; q.dlogits = a->dlogits; q.dbias_w = a->dbias_weight; q.dbias_b = a->dbias_bias;
  q.use_huber = a->use_huber; q.use_hazard = a->use_hazard; q.use_sce = a->use_surviveCE; q.use_ice = a->use_interestCE;
  q.use_ikl = a->use_interestKL; q.mask_loss = a->mask_loss; q.ce_after_focal = a->ce_after_focal; q.kl_after_focal = a->kl_after_focal;
  q.w_huber = a->w_huber; q.w_hazard = a->w_hazard; q.w_sce = a->w_surviveCE; q.w_ice = a->w_interestCE; q.w_ikl = a->w_interestKL;
  const int threads = a->B >= 32 ? 1024 : 32 * a->B;
  loss_kernel<<<1, threads, 0, st>>>(q);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_focal_loss_fwd_bwd(const float* logits, int64_t* gt, int B, int L, const float* exposure_prob, float inv_bsz,
                                      float weight, int rewrite_gt, float* scalars, float* dlogits, mmi_stream_t stream) {
  mmi_loss_args a;
  memset(&a, 0, sizeof(a));
  a.logits = logits; a.gt = gt; a.B = B; a.L = L; a.exposure_prob = exposure_prob;
  a.inv_bsz = inv_bsz; a.w_focal = weight; a.use_focal = 1; a.rewrite_gt = rewrite_gt; a.bpr_scale = 1.0f;
  a.scalars = scalars; a.dlogits = dlogits;
  return mmi_loss_fwd_bwd(&a, stream);
}

extern "C" int64_t mmi_clip_adamw_workspace(int64_t n) { (void)n; return 2 * (int64_t)kRedCtas + 4; }

extern "C" int mmi_clip_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, float max_norm, int step, float* norm_out, void* bf16_out,
                              float* workspace, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && norm_out && workspace, "clip_adamw: null pointer");
  MMI_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0 && (reinterpret_cast<uintptr_t>(grads) & 15) == 0, "clip_adamw: alignment");
  MMI_CHECK_ARG(step >= 1, "clip_adamw: step must be >= 1");
  if (n == 0) return MMI_OK;
  double* partial = reinterpret_cast<double*>(workspace);
  int64_t g64 = (n / 4 + 255) / 256;
  const int grid = (int)(g64 > kRedCtas ? kRedCtas : (g64 < 1 ? 1 : g64));
  sumsq_kernel<<<grid, 256, 0, st>>>(grads, n, partial);
  MMI_CHECK_LAUNCH();
  clip_coef_kernel<<<1, 32, 0, st>>>(partial, grid, max_norm, norm_out);
  MMI_CHECK_LAUNCH();
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int64_t ga = (n + 255) / 256;
  const int grid_a = (int)(ga > kNumSMs * 8 ? kNumSMs * 8 : ga);
  adamw_kernel<<<grid_a, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr * weight_decay, beta1, beta2, eps, (float)(lr / bc1),
                                       (float)sqrt(bc2), max_norm > 0 ? norm_out : nullptr, (__nv_bfloat16*)bf16_out);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_cast_bf16(const float* src, void* dst, int64_t rows, int64_t cols, int transpose, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(src && dst, "cast_bf16: null pointer");
  if (rows * cols == 0) return MMI_OK;
  if (!transpose) {
    int64_t g = (rows * cols + 255) / 256;
    cast_bf16_kernel<<<(int)(g > kNumSMs * 8 ? kNumSMs * 8 : g), 256, 0, st>>>(src, (__nv_bfloat16*)dst, rows * cols);
  } else {
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    cast_bf16_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(src, (__nv_bfloat16*)dst, rows, cols);
  }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}
