// MMI_IMPL_TC attention (bf16, head dim 32): the candidate x history attention of
// models/encoder.py:44-73,138-161 on the 5th-gen tensor cores.
//
// Per CTA (serial pipeline, 2-4 CTAs per SM overlap each other's MMA / softmax phases):
//   fwd      : 128 queries x 64-key tiles.  S = Q K^T (tcgen05.mma 128x64x16, fp32 in TMEM) ->
//              128 threads (one per TMEM lane = query row) read S with tcgen05.ld, apply the
//              reference's "set to -10000 then / sqrt(dh)" mask, online softmax in the exp2
//              domain, write bf16 P into swizzled smem -> O_tile = P V (tcgen05.mma 128x32x16) ->
//              rescale-and-accumulate O in registers.  The two key blocks ([Qa Ka^T | Qb Kb^T])
//              stream through the same running max / sum: one joint softmax, never materialised.
//   bwd dq   : same tiling; recomputes S, dP = dO V^T, dS = P (dP - delta) scale, dQ += dS K.
//   bwd dk/dv: 128 keys x 64-query tiles, transposed formulation (thread = key row):
//              S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q, accumulated in TMEM.
// Operands arrive by TMA (64 B rows, SWIZZLE_64B) straight from the fused-projection buffers:
// head h of a [tokens, ld] tensor is the 32-column box at column h*32.
#include "common.cuh"
#include "tc_common.cuh"

namespace mmi {
namespace tc {

constexpr int DH = 32;
constexpr int QT = 128;                 // rows per CTA (TMEM lanes)
constexpr int NKT = 64;                 // columns per tile
constexpr int ATT_THREADS = 192;        // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 softmax
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr uint32_t TILE64 = NKT * DH * 2;    // 4 KB : 64 rows x 64 B
constexpr uint32_t TILE128 = QT * DH * 2;    // 8 KB : 128 rows x 64 B
constexpr uint32_t PBYTES = QT * NKT * 2;    // 16 KB: 128 rows x 128 B (SWIZZLE_128B, K-major)

struct AttnTcParams {
  int B, H, Lq, nblk;
  int Lk[2];
  const uint8_t* mask_q;
  const uint8_t* mask_k[2];
  __nv_bfloat16* out; int64_t ldo;
  float* lse;
  const __nv_bfloat16* dout; int64_t lddo;
  float* delta;
  __nv_bfloat16* dq[2]; int64_t lddq[2];
  __nv_bfloat16* dk; int64_t lddk;
  __nv_bfloat16* dv; int64_t lddv;
  int which;
  float scale;        // 1/sqrt(dh)
  float scale_log2;   // scale * log2(e)
  float fill_log2;    // -10000 * scale * log2(e)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 64 bf16 (32 packed words) of row `row` into a [rows x 128 B] SWIZZLE_128B K-major tile
__device__ __forceinline__ void write_row_sw128_half(uint8_t* tile, int row, int half, const uint32_t (&w)[16]) {
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int chunk = half * 4 + v;
    *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
  }
}
// bit c of the result = mask[base + c] != 0 for c < count (c in 0..31), whole warp participates
__device__ __forceinline__ uint32_t mask_bits32(const uint8_t* mask, int64_t base, int off, int count, int lane) {
  const int c = off + lane;
  const bool v = (c < count) ? (mask[base + c] != 0) : false;
  return __ballot_sync(0xffffffffu, v);
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

constexpr uint32_t IDESC_S = make_idesc(QT, NKT, false, false);    // 128 x 64, A/B K-major
constexpr uint32_t IDESC_O = make_idesc(QT, DH, false, true);      // 128 x 32, A K-major (P), B MN-major

__device__ __forceinline__ uint64_t desc_k64(uint32_t addr, int kstep) { return make_smem_desc(addr + kstep * 32, 16, 512, 4); }      // K-major SW64
__device__ __forceinline__ uint64_t desc_mn64(uint32_t addr, int kstep) { return make_smem_desc(addr + kstep * 1024, 512, 512, 4); }  // MN-major SW64, 16 rows/step
__device__ __forceinline__ uint64_t desc_p128(uint32_t addr, int kstep) { return make_smem_desc(addr + kstep * 32, 16, 1024, 2); }   // K-major SW128

// ====================================================================================== forward
__global__ void __launch_bounds__(ATT_THREADS, 3)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                   const __grid_constant__ CUtensorMap tmKa, const __grid_constant__ CUtensorMap tmKb,
                   const __grid_constant__ CUtensorMap tmVa, const __grid_constant__ CUtensorMap tmVb, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                       // [2][8 KB]
  uint8_t* sKV = sQ + 2 * TILE128;          // [2 stages][K 4 KB | V 4 KB]
  uint8_t* sP = sKV + 2 * 2 * TILE64;       // 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + PBYTES);
  uint64_t* bar_q = bars;                   // 1
  uint64_t* full_kv = bars + 1;             // [2]
  uint64_t* empty_kv = bars + 3;            // [2]
  uint64_t* s_ready = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* pv_ready = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  int ntile[2] = {(p.Lk[0] + NKT - 1) / NKT, p.nblk > 1 ? (p.Lk[1] + NKT - 1) / NKT : 0};
  const int ntiles = ntile[0] + ntile[1];

  if (warp == 0 && lane == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&full_kv[s], 1); mbar_init(&empty_kv[s], 1); }
    mbar_init(s_ready, 1); mbar_init(p_ready, 128); mbar_init(pv_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + NKT;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_q, (p.nblk > 1 ? 2 : 1) * TILE128);
      tma_load_2d(&tmQa, bar_q, sQ, h * DH, b * p.Lq + q0);
      if (p.nblk > 1) tma_load_2d(&tmQb, bar_q, sQ + TILE128, h * DH, b * p.Lq + q0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1, blk = j < ntile[0] ? 0 : 1, kt = blk ? j - ntile[0] : j;
        mbar_wait(&empty_kv[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&full_kv[s], 2 * TILE64);
        uint8_t* dst = sKV + s * 2 * TILE64;
        const int row = b * p.Lk[blk] + kt * NKT;
        tma_load_2d(blk ? &tmKb : &tmKa, &full_kv[s], dst, h * DH, row);
        tma_load_2d(blk ? &tmVb : &tmVa, &full_kv[s], dst + TILE64, h * DH, row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(bar_q, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1, blk = j < ntile[0] ? 0 : 1;
        mbar_wait(&full_kv[s], (j >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t aQ = smem_u32(sQ + blk * TILE128), aK = smem_u32(sKV + s * 2 * TILE64), aV = aK + TILE64;
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tS, desc_k64(aQ, k), desc_k64(aK, k), IDESC_S, k);
        umma_commit(s_ready);
        mbar_wait(p_ready, j & 1);
        tcgen05_fence_after();
        const uint32_t aP = smem_u32(sP);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tO, desc_p128(aP, k), desc_mn64(aV, k), IDESC_O, k);
        umma_commit(pv_ready);
        umma_commit(&empty_kv[s]);
      }
    }
    __syncwarp();
  } else {
    const int qd = warp & 3, row = qd * 32 + lane;
    const int qi = q0 + row;
    const bool q_in = qi < p.Lq;
    const bool mq = q_in ? (p.mask_q[(int64_t)b * p.Lq + qi] != 0) : false;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    float m = -INFINITY, l = 0.f, o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int blk = j < ntile[0] ? 0 : 1, kt = blk ? j - ntile[0] : j;
      const int k0 = kt * NKT, nvalid = min(NKT, p.Lk[blk] - k0);
      const int64_t mbase = (int64_t)b * p.Lk[blk] + k0;
      const uint32_t w[2] = {mask_bits32(p.mask_k[blk], mbase, 0, nvalid, lane), mask_bits32(p.mask_k[blk], mbase, 32, nvalid, lane)};
      mbar_wait(s_ready, j & 1);
      tcgen05_fence_after();
      // pass 1: row max of the masked, scaled logits (log2 domain)
      float mx = -INFINITY;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (hf * 32 < nvalid) {
          uint32_t r[32];
          tmem_ld_32x32(tS + lane_addr + hf * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (hf * 32 + c < nvalid) {
              const float x = (mq && ((w[hf] >> c) & 1u)) ? __uint_as_float(r[c]) * p.scale_log2 : p.fill_log2;
              mx = fmaxf(mx, x);
            }
          }
        }
      }
      const float m_new = fmaxf(m, mx);
      const float alpha = ex2(m - m_new);
      float rowsum = 0.f;
      // pass 2: p = exp2(x - m_new) -> bf16 -> swizzled smem (A operand of the PV MMA)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t pk[16];
        if (hf * 32 < nvalid) {
          uint32_t r[32];
          tmem_ld_32x32(tS + lane_addr + hf * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float pv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cc = c + e;
              float x = (mq && ((w[hf] >> cc) & 1u)) ? __uint_as_float(r[cc]) * p.scale_log2 : p.fill_log2;
              pv[e] = (hf * 32 + cc < nvalid) ? ex2(x - m_new) : 0.f;
            }
            rowsum += pv[0] + pv[1];
            pk[c >> 1] = pack_bf16(pv[0], pv[1]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) pk[c] = 0u;
        }
        write_row_sw128_half(sP, row, hf, pk);
      }
      l = l * alpha + rowsum;
      m = m_new;
      tcgen05_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_ready);
      mbar_wait(pv_ready, j & 1);
      tcgen05_fence_after();
      {
        uint32_t r[32];
        tmem_ld_32x32(tO + lane_addr, r);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DH; ++d) o[d] = fmaf(o[d], alpha, __uint_as_float(r[d]));
      }
      tcgen05_fence_before();
    }
    if (q_in) {
      const float inv = 1.0f / l;
      __nv_bfloat16* dst = p.out + ((int64_t)b * p.Lq + qi) * p.ldo + h * DH;
#pragma unroll
      for (int d = 0; d < DH; d += 8) {
        *reinterpret_cast<uint4*>(dst + d) = make_uint4(pack_bf16(o[d] * inv, o[d + 1] * inv), pack_bf16(o[d + 2] * inv, o[d + 3] * inv),
                                                        pack_bf16(o[d + 4] * inv, o[d + 5] * inv), pack_bf16(o[d + 6] * inv, o[d + 7] * inv));
      }
      p.lse[((int64_t)b * p.H + h) * p.Lq + qi] = m * kLn2 + logf(l);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
  }
}

// ====================================================================================== backward: dQ
// TMEM: S [0,64) | dP [64,128) | dQ [128,160)  -> 256 columns
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                      const __grid_constant__ CUtensorMap tmKa, const __grid_constant__ CUtensorMap tmKb,
                      const __grid_constant__ CUtensorMap tmVa, const __grid_constant__ CUtensorMap tmVb,
                      const __grid_constant__ CUtensorMap tmdO, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                       // [2][8 KB]
  uint8_t* sdO = sQ + 2 * TILE128;          // 8 KB
  uint8_t* sKV = sdO + TILE128;             // [2][K 4 KB | V 4 KB]
  uint8_t* sdS = sKV + 2 * 2 * TILE64;      // 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + PBYTES);
  uint64_t* bar_q = bars;
  uint64_t* full_kv = bars + 1;
  uint64_t* empty_kv = bars + 3;
  uint64_t* sdp_ready = bars + 5;
  uint64_t* ds_ready = bars + 6;
  uint64_t* ds_free = bars + 7;
  uint64_t* dq_ready = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  int ntile[2] = {(p.Lk[0] + NKT - 1) / NKT, p.nblk > 1 ? (p.Lk[1] + NKT - 1) / NKT : 0};
  const int ntiles = ntile[0] + ntile[1];

  if (warp == 0 && lane == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&full_kv[s], 1); mbar_init(&empty_kv[s], 1); }
    mbar_init(sdp_ready, 1); mbar_init(ds_ready, 128); mbar_init(ds_free, 1); mbar_init(dq_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + NKT, tdQ = tmem + 2 * NKT;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_q, (p.nblk > 1 ? 3 : 2) * TILE128);
      tma_load_2d(&tmQa, bar_q, sQ, h * DH, b * p.Lq + q0);
      if (p.nblk > 1) tma_load_2d(&tmQb, bar_q, sQ + TILE128, h * DH, b * p.Lq + q0);
      tma_load_2d(&tmdO, bar_q, sdO, h * DH, b * p.Lq + q0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1, blk = j < ntile[0] ? 0 : 1, kt = blk ? j - ntile[0] : j;
        mbar_wait(&empty_kv[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&full_kv[s], 2 * TILE64);
        uint8_t* dst = sKV + s * 2 * TILE64;
        const int row = b * p.Lk[blk] + kt * NKT;
        tma_load_2d(blk ? &tmKb : &tmKa, &full_kv[s], dst, h * DH, row);
        tma_load_2d(blk ? &tmVb : &tmVa, &full_kv[s], dst + TILE64, h * DH, row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(bar_q, 0);
      const uint32_t adO = smem_u32(sdO), adS = smem_u32(sdS);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1, blk = j < ntile[0] ? 0 : 1, kt = blk ? j - ntile[0] : j;
        mbar_wait(&full_kv[s], (j >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t aQ = smem_u32(sQ + blk * TILE128), aK = smem_u32(sKV + s * 2 * TILE64), aV = aK + TILE64;
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tS, desc_k64(aQ, k), desc_k64(aK, k), IDESC_S, k);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tdP, desc_k64(adO, k), desc_k64(aV, k), IDESC_S, k);
        umma_commit(sdp_ready);
        mbar_wait(ds_ready, j & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdQ, desc_p128(adS, k), desc_mn64(aK, k), IDESC_O, (kt > 0 || k > 0) ? 1u : 0u);
        umma_commit(ds_free);
        umma_commit(&empty_kv[s]);
        if (kt == ntile[blk] - 1) umma_commit(dq_ready);
      }
    }
    __syncwarp();
  } else {
    const int qd = warp & 3, row = qd * 32 + lane;
    const int qi = q0 + row;
    const bool q_in = qi < p.Lq;
    const bool mq = q_in ? (p.mask_q[(int64_t)b * p.Lq + qi] != 0) : false;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    float delta = 0.f, lse2 = 0.f;
    if (q_in) {
      const __nv_bfloat16* orow = p.out + ((int64_t)b * p.Lq + qi) * p.ldo + h * DH;
      const __nv_bfloat16* dorow = p.dout + ((int64_t)b * p.Lq + qi) * p.lddo + h * DH;
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        const float4 a = load4(orow + d), g = load4(dorow + d);
        delta += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
      }
      const int64_t li = ((int64_t)b * p.H + h) * p.Lq + qi;
      lse2 = p.lse[li] * kLog2e;
      p.delta[li] = delta;
    }
    int blk_phase = 0;
    for (int j = 0; j < ntiles; ++j) {
      const int blk = j < ntile[0] ? 0 : 1, kt = blk ? j - ntile[0] : j;
      const int k0 = kt * NKT, nvalid = min(NKT, p.Lk[blk] - k0);
      const int64_t mbase = (int64_t)b * p.Lk[blk] + k0;
      const uint32_t w[2] = {mask_bits32(p.mask_k[blk], mbase, 0, nvalid, lane), mask_bits32(p.mask_k[blk], mbase, 32, nvalid, lane)};
      mbar_wait(sdp_ready, j & 1);
      tcgen05_fence_after();
      if (j > 0) mbar_wait(ds_free, (j - 1) & 1);   // the previous dQ MMA has finished reading sdS
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t pk[16];
        if (hf * 32 < nvalid) {   // warp-uniform: tcgen05.ld is .sync.aligned
          uint32_t rs[32], rp[32];
          tmem_ld_32x32(tS + lane_addr + hf * 32, rs);
          tmem_ld_32x32(tdP + lane_addr + hf * 32, rp);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float ds[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cc = c + e;
              const bool valid = mq && (((w[hf] >> cc) & 1u) != 0);   // overwritten (masked) logits pass no gradient
              const float pr = ex2(__uint_as_float(rs[cc]) * p.scale_log2 - lse2);
              ds[e] = valid ? pr * (__uint_as_float(rp[cc]) - delta) * p.scale : 0.f;
            }
            pk[c >> 1] = pack_bf16(ds[0], ds[1]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) pk[c] = 0u;
        }
        write_row_sw128_half(sdS, row, hf, pk);
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(ds_ready);
      if (kt == ntile[blk] - 1) {  // dQ of this key block is complete
        mbar_wait(dq_ready, blk_phase);
        blk_phase ^= 1;
        tcgen05_fence_after();
        uint32_t r[32];
        tmem_ld_32x32(tdQ + lane_addr, r);
        tmem_ld_wait();
        if (q_in && p.dq[blk] != nullptr) {
          __nv_bfloat16* dst = p.dq[blk] + ((int64_t)b * p.Lq + qi) * p.lddq[blk] + h * DH;
#pragma unroll
          for (int d = 0; d < DH; d += 8) {
            *reinterpret_cast<uint4*>(dst + d) =
                make_uint4(pack_bf16(__uint_as_float(r[d]), __uint_as_float(r[d + 1])), pack_bf16(__uint_as_float(r[d + 2]), __uint_as_float(r[d + 3])),
                           pack_bf16(__uint_as_float(r[d + 4]), __uint_as_float(r[d + 5])), pack_bf16(__uint_as_float(r[d + 6]), __uint_as_float(r[d + 7])));
          }
        }
        tcgen05_fence_before();
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

// ====================================================================================== backward: dK, dV
// CTA = 128 keys of one key block; loops over 64-query tiles.  thread = key row.
// TMEM: S^T [0,64) | dP^T [64,128) | dK [128,160) | dV [160,192)  -> 256 columns
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;                        // 8 KB (128 keys)
  uint8_t* sV = sK + TILE128;                // 8 KB
  uint8_t* sQdO = sV + TILE128;              // [2 stages][Q 4 KB | dO 4 KB]
  uint8_t* sPT = sQdO + 2 * 2 * TILE64;      // 16 KB
  uint8_t* sdST = sPT + PBYTES;              // 16 KB
  float* sLse = reinterpret_cast<float*>(sdST + PBYTES);   // [2][64]
  float* sDelta = sLse + 2 * NKT;                          // [2][64]
  uint32_t* sMq = reinterpret_cast<uint32_t*>(sDelta + 2 * NKT);  // [2][2] bit masks (valid & in-range queries)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMq + 4);
  uint64_t* bar_kv = bars;
  uint64_t* full_q = bars + 1;
  uint64_t* empty_q = bars + 3;
  uint64_t* sdp_ready = bars + 5;
  uint64_t* pds_ready = bars + 6;
  uint64_t* pds_free = bars + 7;
  uint64_t* dkv_ready = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * QT;
  const int blk = p.which;
  const int Lk = p.Lk[blk];
  const int nq = (p.Lq + NKT - 1) / NKT;

  if (warp == 0 && lane == 0) {
    mbar_init(bar_kv, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&full_q[s], 1); mbar_init(&empty_q[s], 1); }
    mbar_init(sdp_ready, 1); mbar_init(pds_ready, 128); mbar_init(pds_free, 1); mbar_init(dkv_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tST = tmem, tdPT = tmem + NKT, tdK = tmem + 2 * NKT, tdV = tmem + 2 * NKT + DH;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_kv, 2 * TILE128);
      tma_load_2d(&tmK, bar_kv, sK, h * DH, b * Lk + k0);
      tma_load_2d(&tmV, bar_kv, sV, h * DH, b * Lk + k0);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        mbar_wait(&empty_q[s], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&full_q[s], 2 * TILE64);
        uint8_t* dst = sQdO + s * 2 * TILE64;
        const int row = b * p.Lq + i * NKT;
        tma_load_2d(&tmQ, &full_q[s], dst, h * DH, row);
        tma_load_2d(&tmdO, &full_q[s], dst + TILE64, h * DH, row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(bar_kv, 0);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aPT = smem_u32(sPT), adST = smem_u32(sdST);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        mbar_wait(&full_q[s], (i >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t aQ = smem_u32(sQdO + s * 2 * TILE64), adO = aQ + TILE64;
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tST, desc_k64(aK, k), desc_k64(aQ, k), IDESC_S, k);      // S^T  = K Q^T
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tdPT, desc_k64(aV, k), desc_k64(adO, k), IDESC_S, k);    // dP^T = V dO^T
        umma_commit(sdp_ready);
        mbar_wait(pds_ready, i & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdV, desc_p128(aPT, k), desc_mn64(adO, k), IDESC_O, (i > 0 || k > 0) ? 1u : 0u);  // dV += P^T dO
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdK, desc_p128(adST, k), desc_mn64(aQ, k), IDESC_O, (i > 0 || k > 0) ? 1u : 0u);  // dK += dS^T Q
        umma_commit(pds_free);
        umma_commit(&empty_q[s]);
      }
      umma_commit(dkv_ready);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3, row = qd * 32 + lane;
    const int kj = k0 + row;
    const bool k_in = kj < Lk;
    const bool mk = k_in ? (p.mask_k[blk][(int64_t)b * Lk + kj] != 0) : false;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int st = threadIdx.x - 64;   // 0..127 among the softmax threads
    for (int i = 0; i < nq; ++i) {
      const int s = i & 1, qbase = i * NKT, nvalid = min(NKT, p.Lq - qbase);
      // per-query vectors of this tile (stage s is free: its previous user, tile i-2, was consumed before pds_ready(i-2))
      if (st < NKT) {
        const bool in = st < nvalid;
        const int64_t li = ((int64_t)b * p.H + h) * p.Lq + qbase + st;
        sLse[s * NKT + st] = in ? p.lse[li] * kLog2e : 0.f;
        sDelta[s * NKT + st] = in ? p.delta[li] : 0.f;
      }
      if (warp == 2) {
        const uint32_t w0 = mask_bits32(p.mask_q, (int64_t)b * p.Lq + qbase, 0, nvalid, lane);
        const uint32_t w1 = mask_bits32(p.mask_q, (int64_t)b * p.Lq + qbase, 32, nvalid, lane);
        if (lane == 0) { sMq[s * 2] = w0; sMq[s * 2 + 1] = w1; }
      }
      named_bar_sync(1, 128);
      const uint32_t wq[2] = {sMq[s * 2], sMq[s * 2 + 1]};
      mbar_wait(sdp_ready, i & 1);
      tcgen05_fence_after();
      if (i > 0) mbar_wait(pds_free, (i - 1) & 1);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t pp[16], pd[16];
        if (hf * 32 < nvalid) {   // warp-uniform: tcgen05.ld is .sync.aligned
          uint32_t rs[32], rp[32];
          tmem_ld_32x32(tST + lane_addr + hf * 32, rs);
          tmem_ld_32x32(tdPT + lane_addr + hf * 32, rp);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float pr[2], ds[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cc = c + e, qc = hf * 32 + cc;
              const bool valid = mk && (((wq[hf] >> cc) & 1u) != 0);
              const float x = valid ? __uint_as_float(rs[cc]) * p.scale_log2 : p.fill_log2;
              pr[e] = (qc < nvalid) ? ex2(x - sLse[s * NKT + qc]) : 0.f;
              ds[e] = valid ? pr[e] * (__uint_as_float(rp[cc]) - sDelta[s * NKT + qc]) * p.scale : 0.f;
            }
            pp[c >> 1] = pack_bf16(pr[0], pr[1]);
            pd[c >> 1] = pack_bf16(ds[0], ds[1]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) { pp[c] = 0u; pd[c] = 0u; }
        }
        write_row_sw128_half(sPT, row, hf, pp);
        write_row_sw128_half(sdST, row, hf, pd);
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(pds_ready);
    }
    mbar_wait(dkv_ready, 0);
    tcgen05_fence_after();
    uint32_t rk[32], rv[32];
    tmem_ld_32x32(tdK + lane_addr, rk);
    tmem_ld_32x32(tdV + lane_addr, rv);
    tmem_ld_wait();
    if (k_in) {
      if (p.dk != nullptr) {
        __nv_bfloat16* dst = p.dk + ((int64_t)b * Lk + kj) * p.lddk + h * DH;
#pragma unroll
        for (int d = 0; d < DH; d += 8)
          *reinterpret_cast<uint4*>(dst + d) =
              make_uint4(pack_bf16(__uint_as_float(rk[d]), __uint_as_float(rk[d + 1])), pack_bf16(__uint_as_float(rk[d + 2]), __uint_as_float(rk[d + 3])),
                         pack_bf16(__uint_as_float(rk[d + 4]), __uint_as_float(rk[d + 5])), pack_bf16(__uint_as_float(rk[d + 6]), __uint_as_float(rk[d + 7])));
      }
      if (p.dv != nullptr) {
        __nv_bfloat16* dst = p.dv + ((int64_t)b * Lk + kj) * p.lddv + h * DH;
#pragma unroll
        for (int d = 0; d < DH; d += 8)
          *reinterpret_cast<uint4*>(dst + d) =
              make_uint4(pack_bf16(__uint_as_float(rv[d]), __uint_as_float(rv[d + 1])), pack_bf16(__uint_as_float(rv[d + 2]), __uint_as_float(rv[d + 3])),
                         pack_bf16(__uint_as_float(rv[d + 4]), __uint_as_float(rv[d + 5])), pack_bf16(__uint_as_float(rv[d + 6]), __uint_as_float(rv[d + 7])));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

// ====================================================================================== host
static bool map_rows(const void* ptr, int64_t ld, int64_t rows, int width, uint32_t box_rows, CUtensorMap* m) {
  return get_tensor_map(ptr, (uint64_t)width, (uint64_t)rows, (uint64_t)ld, DH, box_rows, CU_TENSOR_MAP_SWIZZLE_64B, m);
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { set_error("attn_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return MMI_ECUDA; }
  return MMI_OK;
}

}  // namespace tc

int attn_tc(int kind, const mmi_attn_args* a, int which, cudaStream_t st) {
  using namespace tc;
  MMI_CHECK_ARG(a->dtype == MMI_BF16 && a->dh == DH, "attn_tc: bf16 with head dim 32 only (got dtype %d, dh %d)", a->dtype, a->dh);
  MMI_CHECK_ARG(a->nblk >= 1 && a->nblk <= 2 && a->B > 0 && a->H > 0 && a->Lq > 0, "attn_tc: bad sizes");
  MMI_CHECK_ARG(a->mask_q && a->out && a->lse, "attn_tc: null pointer");
  const int width = a->H * DH;
  AttnTcParams p{};
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.nblk = a->nblk;
  p.mask_q = a->mask_q;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out); p.ldo = a->ldo; p.lse = a->lse;
  p.dout = reinterpret_cast<const __nv_bfloat16*>(a->dout); p.lddo = a->lddo; p.delta = a->delta;
  p.scale = 1.0f / sqrtf((float)DH);
  p.scale_log2 = p.scale * kLog2e;
  p.fill_log2 = -10000.0f * p.scale * kLog2e;
  for (int i = 0; i < a->nblk; ++i) {
    const mmi_attn_block& s = a->blk[i];
    MMI_CHECK_ARG(s.q && s.k && s.v && s.mask_k && s.Lk > 0, "attn_tc: block %d has null pointer / Lk<=0", i);
    MMI_CHECK_ARG(s.ldq % 8 == 0 && s.ldk % 8 == 0 && s.ldv % 8 == 0, "attn_tc: leading dims must be multiples of 8 (TMA)");
    p.Lk[i] = s.Lk; p.mask_k[i] = s.mask_k;
    p.dq[i] = reinterpret_cast<__nv_bfloat16*>(s.dq); p.lddq[i] = s.lddq;
  }
  if (a->nblk == 1) { p.Lk[1] = 0; p.mask_k[1] = p.mask_k[0]; }
  const int64_t q_rows = (int64_t)a->B * a->Lq;
  static bool cfg_done[3] = {false, false, false};
  if (kind == 0 || kind == 1) {
    CUtensorMap mQ[2], mK[2], mV[2], mdO;
    for (int i = 0; i < 2; ++i) {
      const mmi_attn_block& s = a->blk[i < a->nblk ? i : 0];
      const int64_t k_rows = (int64_t)a->B * s.Lk;
      if (!map_rows(s.q, s.ldq, q_rows, width, QT, &mQ[i])) return MMI_ECUDA;
      if (!map_rows(s.k, s.ldk, k_rows, width, NKT, &mK[i])) return MMI_ECUDA;
      if (!map_rows(s.v, s.ldv, k_rows, width, NKT, &mV[i])) return MMI_ECUDA;
    }
    dim3 grid((a->Lq + QT - 1) / QT, a->H, a->B);
    if (kind == 0) {
      const size_t smem = 2 * TILE128 + 4 * TILE64 + PBYTES + 256 + 1024;
      if (!cfg_done[0]) { int rc = set_smem(attn_fwd_tc_kernel, smem); if (rc) return rc; cfg_done[0] = true; }
      attn_fwd_tc_kernel<<<grid, ATT_THREADS, smem, st>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], p);
    } else {
      MMI_CHECK_ARG(a->dout && a->delta, "attn_tc bwd: null dout/delta");
      MMI_CHECK_ARG(a->lddo % 8 == 0, "attn_tc: lddo must be a multiple of 8");
      if (!map_rows(a->dout, a->lddo, q_rows, width, QT, &mdO)) return MMI_ECUDA;
      const size_t smem = 3 * TILE128 + 4 * TILE64 + PBYTES + 256 + 1024;
      if (!cfg_done[1]) { int rc = set_smem(attn_bwd_dq_tc_kernel, smem); if (rc) return rc; cfg_done[1] = true; }
      attn_bwd_dq_tc_kernel<<<grid, ATT_THREADS, smem, st>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], mdO, p);
    }
  } else {
    MMI_CHECK_ARG(which >= 0 && which < a->nblk, "attn_tc dkv: bad block index %d", which);
    MMI_CHECK_ARG(a->dout && a->delta, "attn_tc bwd: null dout/delta");
    const mmi_attn_block& s = a->blk[which];
    p.which = which;
    p.dk = reinterpret_cast<__nv_bfloat16*>(s.dk); p.lddk = s.lddk;
    p.dv = reinterpret_cast<__nv_bfloat16*>(s.dv); p.lddv = s.lddv;
    const int64_t k_rows = (int64_t)a->B * s.Lk;
    CUtensorMap mQ, mK, mV, mdO;
    if (!map_rows(s.q, s.ldq, q_rows, width, NKT, &mQ)) return MMI_ECUDA;
    if (!map_rows(a->dout, a->lddo, q_rows, width, NKT, &mdO)) return MMI_ECUDA;
    if (!map_rows(s.k, s.ldk, k_rows, width, QT, &mK)) return MMI_ECUDA;
    if (!map_rows(s.v, s.ldv, k_rows, width, QT, &mV)) return MMI_ECUDA;
    dim3 grid((s.Lk + QT - 1) / QT, a->H, a->B);
    const size_t smem = 2 * TILE128 + 4 * TILE64 + 2 * PBYTES + 4 * NKT * 4 + 16 + 256 + 1024;
    if (!cfg_done[2]) { int rc = set_smem(attn_bwd_dkv_tc_kernel, smem); if (rc) return rc; cfg_done[2] = true; }
    attn_bwd_dkv_tc_kernel<<<grid, ATT_THREADS, smem, st>>>(mQ, mK, mV, mdO, p);
  }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

}  // namespace mmi
