// MMI_IMPL_TC attention (bf16, head dim 32): the candidate x history attention of
// models/encoder.py:44-73,138-161 on the 5th-gen tensor cores.
//
// With dh = 32 the tensor pipe needs 128 cycles per 128 x 64 score tile while the exponentials
// need 512 (16 MUFU/clk/SM), so these kernels are built around the SOFTMAX threads, not the MMA:
//   * every score tile is produced ahead of time: S (and dP) are double-buffered in TMEM and the
//     single MMA thread issues tile j+2 as soon as the softmax threads have pulled tile j into
//     registers; P / dS staging in shared memory is double-buffered too, so the softmax warps of
//     the two resident CTAs never wait for the tensor pipe in steady state;
//   * forward is two-pass: pass 1 takes the exact row maximum (one FMNMX per score, no MUFU),
//     pass 2 recomputes S, exponentiates against the final maximum and lets P V accumulate in
//     TMEM across all key tiles -- no running rescale, no TMEM round trip per tile;
//   * per score the fast path (tile without masked keys) is FFMA + EX2 + FADD + half a pack in
//     forward, FFMA + EX2 + FFMA + FMUL + pack in backward; the reference's "set to -10000, then
//     / sqrt(dh)" masking (padded queries get a uniform softmax, masked logits pass no gradient)
//     is handled exactly on a slow path chosen per warp and tile.
//
//   fwd      : CTA = 128 queries x (b, h); 64-key tiles; the two key blocks ([Qa Ka^T | Qb Kb^T])
//              share one softmax.  TMEM: S[2] 128 cols | O 32 cols.
//   bwd dq   : same rows, 32-key tiles: S, dP = dO V^T -> dS = P (dP - delta) scale -> dQ += dS K.
//              TMEM: S[2] | dP[2] | dQa | dQb (192 cols).
//   bwd dk/dv: CTA = 128 keys of one block, 32-query tiles, transposed formulation (thread = key):
//              S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q.  TMEM: S^T[2] | dP^T[2] | dK | dV.
// Operands arrive by TMA (64 B rows, SWIZZLE_64B) straight from the fused-projection buffers:
// head h of a [tokens, ld] tensor is the 32-column box at column h*32.
#include <stdlib.h>

#include "common.cuh"
#include "dropout.cuh"
#include "tc_common.cuh"

namespace mmi {
namespace tc {

constexpr int DH = 32;
constexpr int QT = 128;                 // rows per CTA (TMEM lanes)
constexpr int NT = 32;                  // key / query columns per score tile
constexpr int KV_STAGES = 4;
constexpr int ATT_THREADS = 192;        // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 softmax
constexpr int ATT_TMEM_COLS = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr uint32_t TILE32 = 32 * DH * 2;     // 2 KB : 32 rows x 64 B
constexpr uint32_t TILE128 = QT * DH * 2;    // 8 KB : 128 rows x 64 B

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
  float* dbq[2]; float* dbk; float* dbv;   // optional fused bias-gradient accumulators [H*dh]
  int which;
  float scale;        // 1/sqrt(dh)
  float scale_log2;   // scale * log2(e)
  float fill_log2;    // -10000 * scale * log2(e)
  DropParams drop;    // logits dropout (kernels instantiated with DROP = true): after the -10000 fill, before the scale
};
// Logits dropout (models/encoder.py:145-150) in the three kernels, DROP = true instantiations only:
//   raw logit v = valid ? q.k : -10000;  v <- keep ? v * drop.scale : 0;  softmax over v / sqrt(dh)
// so a dropped logit is 0 whether it was masked or not (the reference's order of operations), and only logits that are
// both valid and kept pass a gradient (times drop.scale).  keep(query row, key) comes from the keep word of
// (row = (b*H + h)*Lq + q, group = (blk << 20) + (k >> 5)): fwd / dq threads own a query row and generate one word per
// 32 keys; dk/dv threads own a key, so lane c generates the word of query c of the tile and the warp shuffles them.
__device__ __forceinline__ uint32_t attn_group(int blk, int k32) { return (static_cast<uint32_t>(blk) << 20) + static_cast<uint32_t>(k32); }
// 32 x 32 bit-matrix transpose across a warp: lane l holds row l on entry and column l on exit (bit c of the result of
// lane l = bit l of the input of lane c).  Five butterfly stages (one shuffle each) instead of 32 shuffles.
__device__ __forceinline__ uint32_t warp_bit_transpose32(uint32_t x, int lane) {
  uint32_t m = 0x0000ffffu;
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
    x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
    m ^= m << (j >> 1);
  }
  return x;
}

#ifdef MMI_ATTN_TRACE
// debug build only: clock64 stamps of one CTA (blockIdx.z == gridDim.z / 2, x == 0, y == 0), read back by mmi_debug_trace
__device__ long long g_trace[8192];
#define TRACE_ON (blockIdx.z == gridDim.z / 2 && blockIdx.x == 0 && blockIdx.y == (gridDim.z == 1 ? gridDim.y / 2 : 0))
#define TRACE(slot) do { if (TRACE_ON && lane == 0) g_trace[(slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2: two results per FMA-pipe issue slot) and a polynomial exp2.
// Measured on B200 (tools/micro/pipe_rates.cu): MUFU.EX2 16 results/clk/SM, FFMA2 126 results/clk/SM.  With one ex2 per
// score the softmax loops are MUFU-bound while the warps compute, so a fixed share of the scores of every tile takes
// its exponential on the FMA / ALU pipes instead: round-to-nearest split x = n + f (magic-number add), 2^f by a
// degree-3 minimax polynomial on [-1/2, 1/2] (relative error 7.5e-5, bf16 keeps 3.9e-3), 2^n added into the exponent.
__device__ __forceinline__ float2 ex2_mufu2(float2 x) { return make_float2(ex2(x.x), ex2(x.y)); }
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  constexpr float kMagic = 12582912.0f;                        // 1.5 * 2^23: x + kMagic holds round(x) in its low mantissa bits
  x.x = fmaxf(x.x, -126.0f); x.y = fmaxf(x.y, -126.0f);        // keep the biased exponent non-negative (2^-126 ~ 0)
  const float2 t = add2(x, splat2(kMagic));
  const float2 n = add2(t, splat2(-kMagic));
  const float2 f = fma2(n, splat2(-1.0f), x);
  float2 q = fma2(f, splat2(0.0551716685295105f), splat2(0.2426111251115799f));
  q = fma2(q, f, splat2(0.6932609677314758f));
  q = fma2(q, f, splat2(0.9999280571937561f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
}
// which score pairs of a 32-score tile take the polynomial: 6 of 16 (3 of every 8), spread so both pipes stay busy
#ifndef MMI_POLY
#define MMI_POLY 0   // measured (tools/attn_bench.py, c2 usr fwd): 0 -> 1.99 ms, 3 of 8 -> 2.02 ms, 4 of 8 -> 2.06 ms: the loops are
#endif               // not MUFU-bound end to end (hand-off latency is), so the polynomial's extra issue slots cost more than they free
__device__ __forceinline__ constexpr bool poly_pair(int pair) {
  return MMI_POLY == 0 ? false
       : MMI_POLY == 2 ? ((pair & 7) == 2 || (pair & 7) == 6)
       : MMI_POLY == 3 ? ((pair & 7) == 2 || (pair & 7) == 5 || (pair & 7) == 7)
       : MMI_POLY == 4 ? ((pair & 1) == 1)
       : ((pair & 7) != 0 && (pair & 7) != 4 && (pair & 7) != 6);   // 5 of 8
}

// 32 bf16 = the whole row `row` of a [rows x 64 B] SWIZZLE_64B K-major tile (address bits [4,6) ^= bits [7,9))
// row_addr = shared-space address of the row (tile + row * 64); swz = (row >> 1) & 3
__device__ __forceinline__ void write_row_sw64(uint32_t row_addr, uint32_t swz, const uint32_t (&w)[16]) {
#pragma unroll
  for (uint32_t v = 0; v < 4; ++v) sts_u4(row_addr + ((v ^ swz) << 4), w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
}
// bit c of the result = mask[base + off + c] != 0 for off + c < count, whole warp participates
__device__ __forceinline__ uint32_t mask_bits32(const uint8_t* mask, int64_t base, int off, int count, int lane) {
  const int c = off + lane;
  const bool v = (c < count) ? (mask[base + c] != 0) : false;
  return __ballot_sync(0xffffffffu, v);
}
__device__ __forceinline__ uint32_t range_bits32(int off, int count) {   // bit c set iff off + c < count
  const int n = count - off;
  return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
}

// Key-validity bit words of every key tile of the CTA, built ONCE (all warps, before the role split) so that the
// softmax loop reads them from shared memory instead of paying a global-load latency per tile.
// words_per_tile = tile_cols / 32; word w of tile j covers keys [j_k0 + 32 w, +32); bits past Lk are 0.
// entry j = {valid bits, in-range bits} of the 32 keys of tile j
__device__ __forceinline__ void build_key_bits(uint2* kb, const AttnTcParams& p, int b, int nt0, int T, int warp, int lane, int nwarps) {
  for (int j = warp; j < T; j += nwarps) {
    const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j;
    const int Lk = blk ? p.Lk[1] : p.Lk[0];
    const uint32_t bits = mask_bits32(blk ? p.mask_k[1] : p.mask_k[0], (int64_t)b * Lk, kt * NT, Lk, lane);
    if (lane == 0) kb[j] = make_uint2(bits, range_bits32(kt * NT, Lk));
  }
}

constexpr uint32_t IDESC_S32 = make_idesc(QT, NT, false, false);   // 128 x 32, A/B K-major
constexpr uint32_t IDESC_O = make_idesc(QT, DH, false, true);          // 128 x 32, A K-major, B MN-major

__device__ __forceinline__ uint64_t desc_k64(uint32_t addr, int kstep) { return make_smem_desc(addr + kstep * 32, 16, 512, 4); }      // K-major SW64
__device__ __forceinline__ uint64_t desc_mn64(uint32_t addr, int kstep) { return make_smem_desc(addr + kstep * 1024, 512, 512, 4); }  // MN-major SW64, 16 rows/step

// barrier slots shared by the three kernels
struct Bars {
  uint64_t once;                 // one-shot operand load (Q / K,V of the CTA's rows)
  uint64_t kv_full[KV_STAGES];
  uint64_t kv_empty[KV_STAGES];
  uint64_t a_ready[2];           // MMA  -> softmax : S (and dP) tile in TMEM
  uint64_t s_free[2];            // softmax -> MMA  : tile pulled into registers
  uint64_t p_ready[2];           // softmax -> MMA  : P / dS staged in shared memory
  uint64_t p_free[2];            // MMA  -> softmax : staging buffer consumed
  uint64_t done;                 // MMA  -> softmax : accumulators complete
  uint32_t tmem_slot;
};

__device__ __forceinline__ void init_bars(Bars* bars, int softmax_threads) {
  mbar_init(&bars->once, 1);
  for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
  for (int b = 0; b < 2; ++b) {
    mbar_init(&bars->a_ready[b], 1);
    mbar_init(&bars->s_free[b], softmax_threads);
    mbar_init(&bars->p_ready[b], softmax_threads);
    mbar_init(&bars->p_free[b], 1);
  }
  mbar_init(&bars->done, 1);
  fence_barrier_init();
}
__device__ __forceinline__ uint32_t tmem_setup(Bars* bars, int warp, uint32_t cols = ATT_TMEM_COLS) {
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  return bars->tmem_slot;
}
__device__ __forceinline__ void tmem_teardown(uint32_t tmem, int warp, uint32_t cols = ATT_TMEM_COLS) {
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
  }
}
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, const uint32_t (&r)[32], float mul) {
#pragma unroll
  for (int d = 0; d < DH; d += 8)
    *reinterpret_cast<uint4*>(dst + d) =
        make_uint4(pack_bf16x2(__uint_as_float(r[d]) * mul, __uint_as_float(r[d + 1]) * mul), pack_bf16x2(__uint_as_float(r[d + 2]) * mul, __uint_as_float(r[d + 3]) * mul),
                   pack_bf16x2(__uint_as_float(r[d + 4]) * mul, __uint_as_float(r[d + 5]) * mul), pack_bf16x2(__uint_as_float(r[d + 6]) * mul, __uint_as_float(r[d + 7]) * mul));
}

// Column sums of a [32 lanes x 32 values] register tile by a transposing butterfly (31 shuffles instead of 160):
// lane i ends with sum over lanes of v[i].  Used by the backward epilogues for the fused bias gradients.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}
__device__ __forceinline__ void add_bias_grad(float* db, const uint32_t (&r)[32], bool row_in, int lane) {
  float f[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) f[c] = row_in ? __uint_as_float(r[c]) : 0.f;
  const float s = warp_colsum32(f, lane);
  atomicAdd(db + lane, s);
}

// ====================================================================================== forward
// smem: Q [2][8 KB] | K,V ring [4][2 KB + 2 KB] | P [2][8 KB] | barriers | key bits      (~50 KB: four CTAs per SM)
// TMEM: S[2] @0,32 | O @64                                                             (128 columns allocated)
//
// Small CTAs on purpose: with dh = 32 every hand-off (mbarrier wait ~100+ cycles even when already complete,
// tcgen05.ld, fence.proxy.async) costs about as much as the arithmetic of a tile, so the kernel is bound by the
// latency of ONE CTA's softmax -> MMA -> softmax chain, not by a pipe.  Four resident CTAs (16 softmax warps, 4 per
// scheduler) overlap those chains and each other's prologue / epilogue.
//
// Single pass over the 32-key tiles with a LAZY running maximum: thread = query row keeps a reference maximum m
// (log2 domain) taken from its first tile and only moves it when a later tile exceeds it by more than kTau;
// P = exp2(x - m) <= 2^kTau is exact in bf16's exponent range, so the result equals the two-pass softmax up to
// rounding.  When m moves, the warp rescales its rows of O in TMEM (tcgen05.ld / .st) after the P V products
// issued so far have retired -- rare on real data, never on the critical path.
constexpr float kTau = 8.0f;
constexpr int FWD_STAGES = 4;

__global__ void __launch_bounds__(ATT_THREADS, 4)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                   const __grid_constant__ CUtensorMap tmKa, const __grid_constant__ CUtensorMap tmKb,
                   const __grid_constant__ CUtensorMap tmVa, const __grid_constant__ CUtensorMap tmVb, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + 2 * TILE128;
  uint8_t* sP = sKV + FWD_STAGES * 2 * TILE32;
  Bars* bars = reinterpret_cast<Bars*>(sP + 2 * TILE128);
  uint2* kbits = reinterpret_cast<uint2*>(bars + 1);

  const int warp = (int)uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;   // warp index in a uniform register
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int nt0 = (p.Lk[0] + NT - 1) / NT, nt1 = p.nblk > 1 ? (p.Lk[1] + NT - 1) / NT : 0;
  const int T = nt0 + nt1;
  const int rows_valid = min(QT, p.Lq - q0);
  const int nact = (rows_valid + 31) >> 5;                  // softmax warps with at least one real query

  auto load_kv = [&](int j) {                            // one elected lane: K, V of tile j into ring stage j % FWD_STAGES
    const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j, st = j % FWD_STAGES;
    mbar_expect_tx(&bars->kv_full[st], 2 * TILE32);
    uint8_t* dst = sKV + st * 2 * TILE32;
    const int row = b * (blk ? p.Lk[1] : p.Lk[0]) + kt * NT;
    tma_load_2d(blk ? &tmKb : &tmKa, &bars->kv_full[st], dst, h * DH, row);
    tma_load_2d(blk ? &tmVb : &tmVa, &bars->kv_full[st], dst + TILE32, h * DH, row);
  };
  // The thread that initialises the barriers also launches the first loads (Q and the whole K/V ring) right away:
  // their latency overlaps the mask reads and the TMEM allocation below instead of following them.
  if (warp == 0 && elect_one()) {
    init_bars(bars, 32 * nact);
    mbar_expect_tx(&bars->once, (p.nblk > 1 ? 2 : 1) * TILE128);
    tma_load_2d(&tmQa, &bars->once, sQ, h * DH, b * p.Lq + q0);
    if (p.nblk > 1) tma_load_2d(&tmQb, &bars->once, sQ + TILE128, h * DH, b * p.Lq + q0);
    for (int j = 0; j < min(T, FWD_STAGES); ++j) load_kv(j);
  }
  build_key_bits(kbits, p, b, nt0, T, warp, lane, ATT_THREADS / 32);
  if (warp == 4) TRACE(4090);
  const uint32_t tmem = tmem_setup(bars, warp, 128);
  const uint32_t tO = tmem + 2 * NT;
  if (warp == 4) TRACE(4091);

  if (warp == 0) {
    // warp-uniform producer loop: all lanes wait for the free stage, one elected lane issues the TMA
    for (int j = FWD_STAGES; j < T; ++j) {
      mbar_wait_bg(&bars->kv_empty[j % FWD_STAGES], ((j / FWD_STAGES) & 1) ^ 1);
      if (elect_one()) load_kv(j);
      __syncwarp();
    }
  } else if (warp == 1) {
    // warp-uniform MMA loop: all lanes wait, one elected lane issues
    const uint32_t tS = uniform(tmem), tOu = uniform(tO);
    mbar_wait(&bars->once, 0);
    auto issue_pv = [&](int u) {                         // O += P(u) V(u)
      const int pb = u & 1, st = u % FWD_STAGES;
      mbar_wait_bg(&bars->p_ready[pb], (u >> 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t aP = smem_u32(sP + pb * TILE128), aV = smem_u32(sKV + st * 2 * TILE32 + TILE32);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tOu, desc_k64(aP, k), desc_mn64(aV, k), IDESC_O, (u > 0 || k > 0) ? 1u : 0u);
        umma_commit(&bars->p_free[pb]);
        umma_commit(&bars->kv_empty[st]);
      }
      __syncwarp();
    };
    for (int j = 0; j < T; ++j) {
      const int blk = j < nt0 ? 0 : 1, st = j % FWD_STAGES, sb = j & 1;
      mbar_wait_bg(&bars->kv_full[st], (j / FWD_STAGES) & 1);
      if (j >= 2) mbar_wait_bg(&bars->s_free[sb], ((j >> 1) - 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t aQ = smem_u32(sQ + blk * TILE128), aK = smem_u32(sKV + st * 2 * TILE32);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tS + sb * NT, desc_k64(aQ, k), desc_k64(aK, k), IDESC_S32, k);
        umma_commit(&bars->a_ready[sb]);                 // also covers P V (j-2): P buffer sb is free once this fires
      }
      __syncwarp();
      if (j >= 1) issue_pv(j - 1);
    }
    issue_pv(T - 1);
    if (elect_one()) umma_commit(&bars->done);
    __syncwarp();
  } else if ((warp & 3) < nact) {
    const int qd = warp & 3, row = qd * 32 + lane;
    const int qi = q0 + row;
    const bool q_in = qi < p.Lq;
    // rows past Lq compute on whatever the TMA box held (finite) and are never stored
    const bool mq = q_in ? (p.mask_q[(int64_t)b * p.Lq + qi] != 0) : true;
    const bool warp_all_mq = __all_sync(0xffffffffu, mq);
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const float scale_t = mq ? p.scale_log2 : 0.f;       // x = s * scale_t + base_t  (a padded query sees `fill` everywhere)
    const float base_t = mq ? 0.f : p.fill_log2;
    float m = 0.f, l0 = 0.f, l1 = 0.f;
    // shared-space addresses of everything the loop touches, computed once
    const uint32_t a_ready_a = smem_u32(&bars->a_ready[0]), s_free_a = smem_u32(&bars->s_free[0]), p_ready_a = smem_u32(&bars->p_ready[0]);
    const uint32_t kb_a = smem_u32(kbits), prow_a = smem_u32(sP) + row * 64, swz = (row >> 1) & 3;
    const uint32_t tS_row = tmem + lane_addr;
    for (int j = 0; j < T; ++j) {
      const uint32_t sb = j & 1;
      const uint2 kb = lds_u2(kb_a + j * 8);
      const uint32_t wv = kb.x, wr = kb.y;
      if (warp == 4) TRACE(j * 8 + 0);
      mbar_wait_a(a_ready_a + sb * 8, (j >> 1) & 1);     // S(j) ready; P buffer sb consumed by P V (j-2)
      if (warp == 4) TRACE(j * 8 + 1);
      tcgen05_fence_after();
      uint32_t r[32];
      tmem_ld_32x32(tS_row + sb * NT, r);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive_a(s_free_a + sb * 8);
      if (warp == 4) TRACE(j * 8 + 2);
      // ---- tile maximum in the log2 domain
      float t;
      {
        float mx = -INFINITY;
        if (wv == 0xffffffffu) {
#pragma unroll
          for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(r[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) if ((wv >> c) & 1u) mx = fmaxf(mx, __uint_as_float(r[c]));
        }
        t = (wr & ~wv) != 0u ? p.fill_log2 : -INFINITY;  // a masked key inside the tile sits at `fill`
        if (mx > -INFINITY) t = fmaxf(t, mx * p.scale_log2);
        if (!mq) t = p.fill_log2;
      }
      if (j == 0) m = t;
      const bool move = j > 0 && t > m + kTau;
      if (__any_sync(0xffffffffu, move)) {
        // move the reference maximum of some rows: rescale l and the warp's rows of O once P V (0..j-1) have retired
        // (tcgen05.ld / .st are warp-collective, so every lane takes part; lanes that keep m multiply by 1)
        mbar_wait(&bars->p_free[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tcgen05_fence_after();
        const float f = move ? ex2(m - t) : 1.0f;
        uint32_t ro[32];
        tmem_ld_32x32(tO + lane_addr, ro);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) ro[c] = __float_as_uint(__uint_as_float(ro[c]) * f);
        tmem_st_32x32(tO + lane_addr, ro);
        tmem_st_wait();
        tcgen05_fence_before();
        l0 *= f; l1 *= f;
        if (move) m = t;
      }
      const float nb_t = base_t - m;                     // x - m = s * scale_t + nb_t
      uint32_t pk[16];
      if (warp_all_mq && wv == 0xffffffffu) {
        float2 l2 = make_float2(l0, l1);
        const float2 sc2 = splat2(scale_t), nb2 = splat2(nb_t);
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float2 x = fma2(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), sc2, nb2);
          const float2 e = poly_pair(c >> 1) ? ex2_poly2(x) : ex2_mufu2(x);
          l2 = add2(l2, e);
          pk[c >> 1] = pack_bf16x2(e.x, e.y);
        }
        l0 = l2.x; l1 = l2.y;
      } else {
        const float pm = ex2(p.fill_log2 - m);           // probability weight of a masked key
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float e[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int cc = c + u;
            const float ev = ex2(fmaf(__uint_as_float(r[cc]), scale_t, nb_t));
            e[u] = ((wv >> cc) & 1u) ? ev : (((wr >> cc) & 1u) ? pm : 0.f);
          }
          l0 += e[0]; l1 += e[1];
          pk[c >> 1] = pack_bf16x2(e[0], e[1]);
        }
      }
      if (warp == 4) TRACE(j * 8 + 3);
      write_row_sw64(prow_a + sb * TILE128, swz, pk);
      fence_proxy_async_smem();
      mbar_arrive_a(p_ready_a + sb * 8);
      if (warp == 4) TRACE(j * 8 + 4);
    }
    if (warp == 4) TRACE(4093);
    mbar_wait(&bars->done, 0);
    tcgen05_fence_after();
    uint32_t ro[32];
    tmem_ld_32x32(tO + lane_addr, ro);
    tmem_ld_wait();
    if (q_in) {
      const float l = l0 + l1;
      store_row32_bf16(p.out + ((int64_t)b * p.Lq + qi) * p.ldo + h * DH, ro, 1.0f / l);
      p.lse[((int64_t)b * p.H + h) * p.Lq + qi] = m * kLn2 + logf(l);
    }
    if (warp == 4) TRACE(4092);
  }
  tmem_teardown(tmem, warp, 128);
}

// ====================================================================================== forward, 64-key tiles
// Same algorithm with HALF the hand-offs per score: one barrier round trip (a_ready wait, s_free / p_ready arrivals,
// fence.proxy.async) now covers 64 keys.  The CUDA-event trace of the 32-key kernel (tools/attn_trace.py) shows ~550 of
// the ~1350 cycles a softmax warp spends per tile going to those hand-offs, not to arithmetic.  TMEM stays at 128
// columns per CTA (S single-buffered 64 | O 32) and shared memory at ~49 KB, so four CTAs still fit on an SM:
//   * S is single-buffered: the softmax thread pulls the tile into registers in two 32-column halves and frees the
//     buffer after the second pull; the MMA warp refills it while the second half is being exponentiated;
//   * K and V ride in separate 2-stage rings: a K stage is free as soon as its S product retired (early), a V stage
//     once its P V product retired, so the next K tile is always in flight one whole tile ahead;
//   * P (bf16, 128 B rows, SWIZZLE_128B) is single-buffered: half 0 is written ~a half tile after p_ready of the
//     previous tile, by which time its P V product has long retired.
// smem: Q [2][8 KB] | K [2][4 KB] | V [2][4 KB] | P 16 KB | barriers | key bits
#ifndef MMI_NS_TMA
#define MMI_NS_TMA 64
#endif
#ifndef MMI_NS_MMA
#define MMI_NS_MMA 32
#endif
constexpr int NT64 = 64;
constexpr uint32_t TILE64 = NT64 * DH * 2;       // 4 KB : 64 rows x 64 B
constexpr uint32_t PTILE64 = QT * NT64 * 2;      // 16 KB: 128 rows x 128 B
constexpr uint32_t IDESC_S64 = make_idesc(QT, NT64, false, false);
__device__ __forceinline__ uint64_t desc_k128(uint32_t addr, int kstep) { return make_smem_desc(addr + kstep * 32, 16, 1024, 2); }   // K-major SW128

// entry j = {valid bits of keys 0-31, 32-63, in-range bits of keys 0-31, 32-63} of 64-key tile j
__device__ __forceinline__ void build_key_bits64(uint4* kb, const AttnTcParams& p, int b, int nt0, int T, int warp, int lane, int nwarps) {
  for (int j = warp; j < T; j += nwarps) {
    const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j;
    const int Lk = blk ? p.Lk[1] : p.Lk[0];
    const uint8_t* mk = blk ? p.mask_k[1] : p.mask_k[0];
    const uint32_t lo = mask_bits32(mk, (int64_t)b * Lk, kt * NT64, Lk, lane), hi = mask_bits32(mk, (int64_t)b * Lk, kt * NT64 + 32, Lk, lane);
    if (lane == 0) kb[j] = make_uint4(lo, hi, range_bits32(kt * NT64, Lk), range_bits32(kt * NT64 + 32, Lk));
  }
}
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

template <bool DROP>
__global__ void __launch_bounds__(ATT_THREADS, 4)
attn_fwd_tc64_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                     const __grid_constant__ CUtensorMap tmKa, const __grid_constant__ CUtensorMap tmKb,
                     const __grid_constant__ CUtensorMap tmVa, const __grid_constant__ CUtensorMap tmVb,
                     const __grid_constant__ CUtensorMap tmO, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * TILE128;
  uint8_t* sV = sK + 2 * TILE64;
  uint8_t* sP = sV + 2 * TILE64;                          // 32 KB from the base: 1024-aligned
  Bars* bars = reinterpret_cast<Bars*>(sP + PTILE64);
  uint4* kbits = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(bars + 1) + 15) & ~static_cast<uintptr_t>(15));

  const int warp = (int)uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int nt0 = (p.Lk[0] + NT64 - 1) / NT64, nt1 = p.nblk > 1 ? (p.Lk[1] + NT64 - 1) / NT64 : 0;
  const int T = nt0 + nt1;
  const int rows_valid = min(QT, p.Lq - q0);
  const int nact = (rows_valid + 31) >> 5;
  // barrier slots: kv_full / kv_empty [0,1] = K ring, [2,3] = V ring; a_ready[0], s_free[0], p_ready[0], p_free[0]
  auto load_k = [&](int j) {
    const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j, st = j & 1;
    mbar_expect_tx(&bars->kv_full[st], TILE64);
    tma_load_2d(blk ? &tmKb : &tmKa, &bars->kv_full[st], sK + st * TILE64, h * DH, b * (blk ? p.Lk[1] : p.Lk[0]) + kt * NT64);
  };
  auto load_v = [&](int j) {
    const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j, st = j & 1;
    mbar_expect_tx(&bars->kv_full[2 + st], TILE64);
    tma_load_2d(blk ? &tmVb : &tmVa, &bars->kv_full[2 + st], sV + st * TILE64, h * DH, b * (blk ? p.Lk[1] : p.Lk[0]) + kt * NT64);
  };
  if (warp == 0 && elect_one()) {
    init_bars(bars, 32 * nact);
    mbar_expect_tx(&bars->once, (p.nblk > 1 ? 2 : 1) * TILE128);
    tma_load_2d(&tmQa, &bars->once, sQ, h * DH, b * p.Lq + q0);
    if (p.nblk > 1) tma_load_2d(&tmQb, &bars->once, sQ + TILE128, h * DH, b * p.Lq + q0);
    for (int j = 0; j < min(T, 2); ++j) { load_k(j); load_v(j); }
  }
  build_key_bits64(kbits, p, b, nt0, T, warp, lane, ATT_THREADS / 32);
  const uint32_t tmem = tmem_setup(bars, warp, 128);
  const uint32_t tO = tmem + NT64;

  if (warp == 0) {
    for (int j = 2; j < T; ++j) {
      const uint32_t par = ((j >> 1) & 1) ^ 1;
      mbar_wait_bg(&bars->kv_empty[j & 1], par, MMI_NS_TMA);
      if (elect_one()) load_k(j);
      __syncwarp();
      mbar_wait_bg(&bars->kv_empty[2 + (j & 1)], par, MMI_NS_TMA);
      if (elect_one()) load_v(j);
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t tS = uniform(tmem), tOu = uniform(tO);
    mbar_wait(&bars->once, 0);
    const uint32_t aP = smem_u32(sP);
    auto issue_pv = [&](int u) {                         // O += P(u) V(u)
      const int st = u & 1;
      mbar_wait_bg(&bars->kv_full[2 + st], (u >> 1) & 1, MMI_NS_MMA);
      mbar_wait_bg(&bars->p_ready[0], u & 1, MMI_NS_MMA);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t aV = smem_u32(sV + st * TILE64);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tOu, desc_k128(aP, k), desc_mn64(aV, k), IDESC_O, (u > 0 || k > 0) ? 1u : 0u);
        umma_commit(&bars->p_free[0]);
        umma_commit(&bars->kv_empty[2 + st]);
      }
      __syncwarp();
    };
    for (int j = 0; j < T; ++j) {
      const int blk = j < nt0 ? 0 : 1, st = j & 1;
      mbar_wait_bg(&bars->kv_full[st], (j >> 1) & 1, MMI_NS_MMA);
      if (j >= 1) mbar_wait_bg(&bars->s_free[0], (j - 1) & 1, MMI_NS_MMA);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t aQ = smem_u32(sQ + blk * TILE128), aK = smem_u32(sK + st * TILE64);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tS, desc_k64(aQ, k), desc_k64(aK, k), IDESC_S64, k);
        umma_commit(&bars->kv_empty[st]);
        umma_commit(&bars->a_ready[0]);
      }
      __syncwarp();
      if (j >= 1) issue_pv(j - 1);
    }
    issue_pv(T - 1);
    if (elect_one()) umma_commit(&bars->done);
    __syncwarp();
  } else if ((warp & 3) < nact) {
    const int qd = warp & 3, row = qd * 32 + lane;
    const int qi = q0 + row;
    const bool q_in = qi < p.Lq;
    const bool mq = q_in ? (p.mask_q[(int64_t)b * p.Lq + qi] != 0) : true;
    const bool warp_all_mq = __all_sync(0xffffffffu, mq);
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    // DROP: the fill and the dropout are applied to the raw scores in registers, after which every in-range key is an
    // ordinary logit with scale scale_log2 * drop.scale (padded query rows included)
    const float scale_d = p.scale_log2 * p.drop.scale;
    const float scale_t = DROP ? scale_d : (mq ? p.scale_log2 : 0.f);
    const float base_t = DROP ? 0.f : (mq ? 0.f : p.fill_log2);
    const uint32_t rowh = DROP ? drop_rowhash(p.drop.key, (uint64_t)(((int64_t)b * p.H + h) * p.Lq + qi)) : 0u;
    float m = 0.f, l0 = 0.f, l1 = 0.f;
    const uint32_t a_ready_a = smem_u32(&bars->a_ready[0]), s_free_a = smem_u32(&bars->s_free[0]), p_ready_a = smem_u32(&bars->p_ready[0]);
    const uint32_t p_free_a = smem_u32(&bars->p_free[0]);
    const uint32_t kb_a = smem_u32(kbits), prow_a = smem_u32(sP) + row * 128, swz = row & 7;
    const uint32_t tS_row = tmem + lane_addr;
    for (int j = 0; j < T; ++j) {
      const uint4 kb = lds_u4(kb_a + j * 16);
      mbar_wait_a(a_ready_a, j & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const uint32_t wr = hf ? kb.w : kb.z;
        const uint32_t wv = DROP ? wr : (hf ? kb.y : kb.x);              // DROP: every in-range key carries a logit
        uint32_t r[32];
        tmem_ld_32x32(tS_row + hf * 32, r);
        tmem_ld_wait();
        if (hf == 1) {                                   // the whole tile is in registers: the MMA warp may refill S
          tcgen05_fence_before();
          mbar_arrive_a(s_free_a);
        }
        if constexpr (DROP) {
          const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j;
          const uint32_t kw = drop_keep_word(rowh, attn_group(blk, kt * 2 + hf), p.drop.thr8);
          const uint32_t vb = mq ? (hf ? kb.y : kb.x) : 0u;             // logits that are not overwritten with -10000
          if (vb == 0xffffffffu) {
#pragma unroll
            for (int c = 0; c < 32; ++c) r[c] = ((kw >> c) & 1u) ? r[c] : 0u;
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const uint32_t v = ((vb >> c) & 1u) ? r[c] : __float_as_uint(-10000.0f);
              r[c] = ((kw >> c) & 1u) ? v : 0u;
            }
          }
        }
        // ---- maximum of this half in the log2 domain
        float t;
        {
          float mx = -INFINITY;
          if (wv == 0xffffffffu) {
#pragma unroll
            for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(r[c]));
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) if ((wv >> c) & 1u) mx = fmaxf(mx, __uint_as_float(r[c]));
          }
          if constexpr (DROP) {
            t = mx > -INFINITY ? mx * scale_d : -INFINITY;
          } else {
            t = (wr & ~wv) != 0u ? p.fill_log2 : -INFINITY;
            if (mx > -INFINITY) t = fmaxf(t, mx * p.scale_log2);
            if (!mq) t = p.fill_log2;
          }
        }
        const bool first = j == 0 && hf == 0;
        if (first) m = t;
        const bool move = !first && t > m + kTau;
        if (__any_sync(0xffffffffu, move)) {
          // rare: move the reference maximum of some rows -- rescale l, the rows of O (once every P V product issued so
          // far has retired) and, in the second half, the half-0 probabilities this thread already staged
          if (j >= 1) mbar_wait(&bars->p_free[0], (j - 1) & 1);
          tcgen05_fence_after();
          const float f = move ? ex2(m - t) : 1.0f;
          if (j >= 1) {
            uint32_t ro[32];
            tmem_ld_32x32(tO + lane_addr, ro);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) ro[c] = __float_as_uint(__uint_as_float(ro[c]) * f);
            tmem_st_32x32(tO + lane_addr, ro);
            tmem_st_wait();
            tcgen05_fence_before();
          }
          if (hf == 1) {
#pragma unroll
            for (uint32_t v = 0; v < 4; ++v) {
              const uint32_t a = prow_a + ((v ^ swz) << 4);
              uint4 w = lds_u4(a);
              uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
                ww[i] = pack_bf16x2(__uint_as_float(ww[i] << 16) * f, __uint_as_float(ww[i] & 0xffff0000u) * f);
              sts_u4(a, ww[0], ww[1], ww[2], ww[3]);
            }
          }
          l0 *= f; l1 *= f;
          if (move) m = t;
        }
        const float nb_t = base_t - m;
        uint32_t pk[16];
        if ((DROP || warp_all_mq) && wv == 0xffffffffu) {
          float2 l2 = make_float2(l0, l1);
          const float2 sc2 = splat2(scale_t), nb2 = splat2(nb_t);
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float2 x = fma2(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), sc2, nb2);
            const float2 e = poly_pair(c >> 1) ? ex2_poly2(x) : ex2_mufu2(x);
            l2 = add2(l2, e);
            pk[c >> 1] = pack_bf16x2(e.x, e.y);
          }
          l0 = l2.x; l1 = l2.y;
        } else {
          const float pm = ex2(p.fill_log2 - m);         // (DROP: wv == wr, never selected)
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float e[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int cc = c + u;
              const float ev = ex2(fmaf(__uint_as_float(r[cc]), scale_t, nb_t));
              e[u] = ((wv >> cc) & 1u) ? ev : (((wr >> cc) & 1u) ? pm : 0.f);
            }
            l0 += e[0]; l1 += e[1];
            pk[c >> 1] = pack_bf16x2(e[0], e[1]);
          }
        }
        if (hf == 0 && j >= 1) mbar_wait_a(p_free_a, (j - 1) & 1);      // P V (j-1) has consumed the staging tile
#pragma unroll
        for (uint32_t v = 0; v < 4; ++v)
          sts_u4(prow_a + (((hf * 4 + v) ^ swz) << 4), pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive_a(p_ready_a);
    }
    mbar_wait(&bars->done, 0);
    tcgen05_fence_after();
    uint32_t ro[32];
    tmem_ld_32x32(tO + lane_addr, ro);
    tmem_ld_wait();
    {
      // O rows leave by TMA (3-D map [B, Lq, H dh]: rows past Lq of this batch item are clipped): thread-per-row global
      // stores put 32 different lines into every STG.  The warp's 32 x 64 B slice is staged in the Q buffer, which no
      // product reads any more (`done` covers every MMA of the CTA).
      const float inv = 1.0f / (l0 + l1);
      const uint32_t out_a = smem_u32(sQ) + (uint32_t)qd * (32 * DH * 2), orow_a = out_a + lane * 64, oswz = (lane >> 1) & 3;
#pragma unroll
      for (uint32_t v = 0; v < 4; ++v)
        sts_u4(orow_a + ((v ^ oswz) << 4),
               pack_bf16x2(__uint_as_float(ro[8 * v]) * inv, __uint_as_float(ro[8 * v + 1]) * inv),
               pack_bf16x2(__uint_as_float(ro[8 * v + 2]) * inv, __uint_as_float(ro[8 * v + 3]) * inv),
               pack_bf16x2(__uint_as_float(ro[8 * v + 4]) * inv, __uint_as_float(ro[8 * v + 5]) * inv),
               pack_bf16x2(__uint_as_float(ro[8 * v + 6]) * inv, __uint_as_float(ro[8 * v + 7]) * inv));
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        tma_store_3d(&tmO, reinterpret_cast<const void*>(sQ + qd * (32 * DH * 2)), h * DH, q0 + qd * 32, b);
        bulk_commit();
      }
      __syncwarp();
      if (q_in) p.lse[((int64_t)b * p.H + h) * p.Lq + qi] = m * kLn2 + logf(l0 + l1);
      if (lane == 0) bulk_wait_read0();                    // the staging tile must outlive the store's read
      __syncwarp();
    }
  }
  tmem_teardown(tmem, warp, 128);
}

// ====================================================================================== backward: dQ
// smem: Q [2][8 KB] | dO 8 KB | K,V ring [3][2 KB + 2 KB] | dS [2][8 KB] | barriers | key bits   (~53 KB: four CTAs / SM)
// TMEM: S @0 | dP @32 | dQa @64 | dQb @96                                                      (128 columns)
// S / dP are single-buffered: the softmax threads pull a tile into registers first thing (s_free), so the MMA warp
// refills the buffer while they compute.  dS staging is double-buffered; buffer (j & 1) is known to be free when
// a_ready(j) fires because that commit also covers the dQ product of tile j - 2.
constexpr int BWD_STAGES = 3;

template <bool DROP>
__global__ void __launch_bounds__(ATT_THREADS, 4)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                      const __grid_constant__ CUtensorMap tmKa, const __grid_constant__ CUtensorMap tmKb,
                      const __grid_constant__ CUtensorMap tmVa, const __grid_constant__ CUtensorMap tmVb,
                      const __grid_constant__ CUtensorMap tmdO, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + 2 * TILE128;
  uint8_t* sKV = sdO + TILE128;
  uint8_t* sdS = sKV + BWD_STAGES * 2 * TILE32;           // 36 KB from the base: still 1024-aligned
  Bars* bars = reinterpret_cast<Bars*>(sdS + 2 * TILE128);
  uint2* kbits = reinterpret_cast<uint2*>(bars + 1);

  const int warp = (int)uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;   // warp index in a uniform register
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int nt0 = (p.Lk[0] + NT - 1) / NT, nt1 = p.nblk > 1 ? (p.Lk[1] + NT - 1) / NT : 0;
  const int T = nt0 + nt1;
  const int rows_valid = min(QT, p.Lq - q0);
  const int nact = (rows_valid + 31) >> 5;

  auto load_kv = [&](int j) {                            // one elected lane: K, V of tile j into ring stage j % BWD_STAGES
    const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j, st = j % BWD_STAGES;
    mbar_expect_tx(&bars->kv_full[st], 2 * TILE32);
    uint8_t* dst = sKV + st * 2 * TILE32;
    const int row = b * (blk ? p.Lk[1] : p.Lk[0]) + kt * NT;
    tma_load_2d(blk ? &tmKb : &tmKa, &bars->kv_full[st], dst, h * DH, row);
    tma_load_2d(blk ? &tmVb : &tmVa, &bars->kv_full[st], dst + TILE32, h * DH, row);
  };
  if (warp == 0 && elect_one()) {                        // first loads leave before the mask reads / TMEM allocation
    init_bars(bars, 32 * nact);
    mbar_expect_tx(&bars->once, (p.nblk > 1 ? 3 : 2) * TILE128);
    tma_load_2d(&tmQa, &bars->once, sQ, h * DH, b * p.Lq + q0);
    if (p.nblk > 1) tma_load_2d(&tmQb, &bars->once, sQ + TILE128, h * DH, b * p.Lq + q0);
    tma_load_2d(&tmdO, &bars->once, sdO, h * DH, b * p.Lq + q0);
    for (int j = 0; j < min(T, BWD_STAGES); ++j) load_kv(j);
  }
  build_key_bits(kbits, p, b, nt0, T, warp, lane, ATT_THREADS / 32);
  // per-row operands of the softmax threads (mask byte, lse, the O and dO rows behind delta) are requested BEFORE the TMEM
  // allocation / CTA-wide sync and consumed after it, so their global-load latency is not exposed
  const int qi_pre = q0 + (warp & 3) * 32 + lane;
  const bool pre_in = warp >= 2 && qi_pre < p.Lq;
  uint8_t mq_pre = 0;
  float lse_pre = 0.f;
  uint2 o_pre[DH / 4], do_pre[DH / 4];
  if (pre_in) {
    mq_pre = p.mask_q[(int64_t)b * p.Lq + qi_pre];
    lse_pre = p.lse[((int64_t)b * p.H + h) * p.Lq + qi_pre];
    const uint2* orow = reinterpret_cast<const uint2*>(p.out + ((int64_t)b * p.Lq + qi_pre) * p.ldo + h * DH);
    const uint2* dorow = reinterpret_cast<const uint2*>(p.dout + ((int64_t)b * p.Lq + qi_pre) * p.lddo + h * DH);
#pragma unroll
    for (int d = 0; d < DH / 4; ++d) { o_pre[d] = orow[d]; do_pre[d] = dorow[d]; }
  }
  const uint32_t tmem = tmem_setup(bars, warp, 128);
  const uint32_t tdP = tmem + NT, tdQ = tmem + 2 * NT;

  if (warp == 0) {
    for (int j = BWD_STAGES; j < T; ++j) {
      mbar_wait_bg(&bars->kv_empty[j % BWD_STAGES], ((j / BWD_STAGES) & 1) ^ 1);
      if (elect_one()) load_kv(j);
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t tS = uniform(tmem), tdPu = uniform(tdP), tdQu = uniform(tdQ);
    mbar_wait(&bars->once, 0);
    const uint32_t adO = smem_u32(sdO);
    auto issue_dq = [&](int u) {                       // dQ[blk(u)] += dS(u) K(u)
      const int pb = u & 1, st = u % BWD_STAGES, blk = u < nt0 ? 0 : 1, kt = blk ? u - nt0 : u;
      mbar_wait_bg(&bars->p_ready[pb], (u >> 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t adS = smem_u32(sdS + pb * TILE128), aK = smem_u32(sKV + st * 2 * TILE32);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tdQu + blk * DH, desc_k64(adS, k), desc_mn64(aK, k), IDESC_O, (kt > 0 || k > 0) ? 1u : 0u);
        umma_commit(&bars->kv_empty[st]);
      }
      __syncwarp();
    };
    for (int j = 0; j < T; ++j) {
      const int blk = j < nt0 ? 0 : 1, st = j % BWD_STAGES;
      mbar_wait_bg(&bars->kv_full[st], (j / BWD_STAGES) & 1);
      if (j >= 1) mbar_wait_bg(&bars->s_free[0], (j - 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t aQ = smem_u32(sQ + blk * TILE128), aK = smem_u32(sKV + st * 2 * TILE32), aV = aK + TILE32;
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tS, desc_k64(aQ, k), desc_k64(aK, k), IDESC_S32, k);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tdPu, desc_k64(adO, k), desc_k64(aV, k), IDESC_S32, k);
        umma_commit(&bars->a_ready[0]);                  // also covers dQ (j-2): dS buffer (j & 1) is free once this fires
      }
      __syncwarp();
      if (j >= 1) issue_dq(j - 1);
    }
    issue_dq(T - 1);
    if (elect_one()) umma_commit(&bars->done);
    __syncwarp();
  } else if ((warp & 3) < nact) {
    const int qd = warp & 3, row = qd * 32 + lane;
    const int qi = q0 + row;
    const bool q_in = qi < p.Lq;
    const bool mq = q_in ? (mq_pre != 0) : false;
    const bool warp_all_mq = __all_sync(0xffffffffu, mq);
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    float delta = 0.f, nlse2 = -INFINITY;               // rows past Lq: p = exp2(-inf) = 0
    if (q_in) {
#pragma unroll
      for (int d = 0; d < DH / 4; ++d) {
        const uint2 a = o_pre[d], g = do_pre[d];
        delta += __uint_as_float(a.x << 16) * __uint_as_float(g.x << 16) + __uint_as_float(a.x & 0xffff0000u) * __uint_as_float(g.x & 0xffff0000u) +
                 __uint_as_float(a.y << 16) * __uint_as_float(g.y << 16) + __uint_as_float(a.y & 0xffff0000u) * __uint_as_float(g.y & 0xffff0000u);
      }
      nlse2 = -lse_pre * kLog2e;
      p.delta[((int64_t)b * p.H + h) * p.Lq + qi] = delta;
    }
    const float nds = -delta * p.scale;                  // dS = P * (dP * scale + nds)
    const uint32_t a_ready_a = smem_u32(&bars->a_ready[0]), s_free_a = smem_u32(&bars->s_free[0]), p_ready_a = smem_u32(&bars->p_ready[0]);
    const uint32_t kb_a = smem_u32(kbits), dsrow_a = smem_u32(sdS) + row * 64, swz = (row >> 1) & 3;
    const uint32_t rowh = DROP ? drop_rowhash(p.drop.key, (uint64_t)(((int64_t)b * p.H + h) * p.Lq + qi)) : 0u;
    for (int j = 0; j < T; ++j) {
      const uint32_t pb = j & 1;
      const uint32_t wv = lds_u1(kb_a + j * 8);
      const bool fast = !DROP && warp_all_mq && wv == 0xffffffffu;
      uint32_t ve = 0u;                                  // DROP: logits of this row that are valid AND kept (the only ones with a gradient)
      if constexpr (DROP) {
        const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j;
        ve = (mq ? wv : 0u) & drop_keep_word(rowh, attn_group(blk, kt), p.drop.thr8);
      }
      mbar_wait_a(a_ready_a, j & 1);                     // S(j), dP(j) ready; dS buffer pb consumed by dQ (j-2)
      tcgen05_fence_after();
      uint32_t pk[16];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t rs[16], rp[16];
        tmem_ld_32x32b_x16(tmem + lane_addr + hf * 16, rs);
        tmem_ld_32x32b_x16(tdP + lane_addr + hf * 16, rp);
        tmem_ld_wait();
        if (hf == 1) {                                   // the whole tile is in registers: hand the buffers back
          tcgen05_fence_before();
          mbar_arrive_a(s_free_a);
        }
        if constexpr (DROP) {
          // surviving logit = q.k * drop.scale * scale; every other entry gets dS = 0 by a SELECT (its P may be inf / NaN)
          const float ds_ = p.drop.scale;
          const float2 sl2 = splat2(p.scale_log2 * ds_), nl2 = splat2(nlse2), sc2 = splat2(p.scale * ds_), nd2 = splat2(nds * ds_);
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float2 x = fma2(make_float2(__uint_as_float(rs[c]), __uint_as_float(rs[c + 1])), sl2, nl2);
            const float2 pr = ex2_mufu2(x);
            const float2 ds = mul2(pr, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, nd2));
            const float d0 = ((ve >> (hf * 16 + c)) & 1u) ? ds.x : 0.f, d1 = ((ve >> (hf * 16 + c + 1)) & 1u) ? ds.y : 0.f;
            pk[hf * 8 + (c >> 1)] = pack_bf16x2(d0, d1);
          }
        } else if (fast) {
          const float2 sl2 = splat2(p.scale_log2), nl2 = splat2(nlse2), sc2 = splat2(p.scale), nd2 = splat2(nds);
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float2 x = fma2(make_float2(__uint_as_float(rs[c]), __uint_as_float(rs[c + 1])), sl2, nl2);
            const float2 pr = poly_pair(c >> 1) ? ex2_poly2(x) : ex2_mufu2(x);
            const float2 ds = mul2(pr, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, nd2));
            pk[hf * 8 + (c >> 1)] = pack_bf16x2(ds.x, ds.y);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float ds[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int cc = c + t;
              const bool valid = mq && (((wv >> (hf * 16 + cc)) & 1u) != 0);   // overwritten (masked) logits pass no gradient
              const float pr = ex2(fmaf(__uint_as_float(rs[cc]), p.scale_log2, nlse2));
              ds[t] = valid ? pr * fmaf(__uint_as_float(rp[cc]), p.scale, nds) : 0.f;
            }
            pk[hf * 8 + (c >> 1)] = pack_bf16x2(ds[0], ds[1]);
          }
        }
      }
      write_row_sw64(dsrow_a + pb * TILE128, swz, pk);
      fence_proxy_async_smem();
      mbar_arrive_a(p_ready_a + pb * 8);
    }
    mbar_wait(&bars->done, 0);
    tcgen05_fence_after();
    uint32_t ra[32];
    tmem_ld_32x32(tdQ + lane_addr, ra);
    tmem_ld_wait();
    if (q_in && p.dq[0] != nullptr) store_row32_bf16(p.dq[0] + ((int64_t)b * p.Lq + qi) * p.lddq[0] + h * DH, ra, 1.0f);
    if (p.dbq[0] != nullptr) add_bias_grad(p.dbq[0] + h * DH, ra, q_in, lane);
    if (p.nblk > 1) {
      tmem_ld_32x32(tdQ + DH + lane_addr, ra);
      tmem_ld_wait();
      if (q_in && p.dq[1] != nullptr) store_row32_bf16(p.dq[1] + ((int64_t)b * p.Lq + qi) * p.lddq[1] + h * DH, ra, 1.0f);
      if (p.dbq[1] != nullptr) add_bias_grad(p.dbq[1] + h * DH, ra, q_in, lane);
    }
  }
  tmem_teardown(tmem, warp, 128);
}

// ====================================================================================== backward: dK, dV
// CTA = 128 keys of one key block; loops over 32-query tiles.  thread = key row.
// smem: K 8 KB | V 8 KB | Q,dO ring [3][2 KB + 2 KB] | P^T 8 KB | dS^T 8 KB | per-query vectors | barriers   (~46 KB)
// TMEM: S^T @0 | dP^T @32 | dK @64 | dV @96                                                                (128 columns)
// Everything single-buffered: S^T / dP^T are pulled into registers at the top of a tile (s_free) and refilled
// while the threads compute; P^T / dS^T are rewritten only after the dK / dV products of the previous tile have
// retired (p_free), which happened long before in steady state.
struct QVec {
  float nlse2[NT];         // -lse * log2(e)          (+/-inf tricks: -inf for queries past Lq => P = 0)
  float nds[NT];           // -delta * scale
  uint32_t mq;             // valid-query bits
  uint32_t pad[3];
  uint32_t rh[NT];         // DROP: dropout row hash of each query (drop_rowhash of (b*H + h)*Lq + q)
};
constexpr uint32_t QVEC_RH_OFF = 2 * NT * 4 + 16;

template <bool DROP>
__global__ void __launch_bounds__(ATT_THREADS, 4)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE128;
  uint8_t* sQdO = sV + TILE128;
  uint8_t* sPT = sQdO + 4 * 2 * TILE32;                   // ring region sized for 4 stages keeps the tiles 1024-aligned
  uint8_t* sdST = sPT + TILE128;
  QVec* qv = reinterpret_cast<QVec*>(sdST + TILE128);
  Bars* bars = reinterpret_cast<Bars*>(qv + BWD_STAGES);

  const int warp = (int)uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;   // warp index in a uniform register
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * QT;
  const int blk = p.which;
  const int Lk = (blk ? p.Lk[1] : p.Lk[0]);
  const int T = (p.Lq + NT - 1) / NT;
  const int rows_valid = min(QT, Lk - k0);
  const int nact = (rows_valid + 31) >> 5;

  // Producer state (warp 0).  The per-query vectors (one query per lane) are fetched one tile AHEAD so that their
  // global-load latency overlaps the wait for a free stage instead of pacing the whole pipeline.
  float lse_n = 0.f, delta_n = 0.f;
  uint8_t mq_n = 0;
  auto fetch = [&](int i) {
    const int qi = i * NT + lane;
    if (qi < p.Lq) {
      const int64_t li = ((int64_t)b * p.H + h) * p.Lq + qi;
      lse_n = p.lse[li]; delta_n = p.delta[li]; mq_n = p.mask_q[(int64_t)b * p.Lq + qi];
    } else { lse_n = INFINITY; delta_n = 0.f; mq_n = 0; }
  };
  auto produce = [&](int i, bool wait) {                 // whole warp 0: per-query vectors + Q, dO of query tile i into stage i % BWD_STAGES
    const int st = i % BWD_STAGES;
    const float lse_c = lse_n, delta_c = delta_n;
    const bool mq_c = mq_n != 0;
    if (wait && i + 1 < T) fetch(i + 1);                 // steady state: the next tile's vectors are requested one tile ahead
    if (wait) mbar_wait_bg(&bars->kv_empty[st], ((i / BWD_STAGES) & 1) ^ 1);
    qv[st].nlse2[lane] = -lse_c * kLog2e;                // queries past Lq: -inf => P = 0
    qv[st].nds[lane] = -delta_c * p.scale;
    if constexpr (DROP) qv[st].rh[lane] = drop_rowhash(p.drop.key, (uint64_t)(((int64_t)b * p.H + h) * p.Lq + i * NT + lane));
    const uint32_t mqb = __ballot_sync(0xffffffffu, mq_c);
    if (lane == 0) qv[st].mq = mqb;
    __syncwarp();
    if (elect_one()) {
      mbar_expect_tx(&bars->kv_full[st], 2 * TILE32);
      uint8_t* dst = sQdO + st * 2 * TILE32;
      const int row = b * p.Lq + i * NT;
      tma_load_2d(&tmQ, &bars->kv_full[st], dst, h * DH, row);
      tma_load_2d(&tmdO, &bars->kv_full[st], dst + TILE32, h * DH, row);
    }
    __syncwarp();
  };
  if (warp == 0) {
    // K, V of the CTA's rows AND the first query tiles (one per ring stage) leave before the TMEM allocation and the
    // CTA-wide sync: every load of the prologue is in flight at once (candidate-side launches have only 2 query tiles)
    fetch(0);
    if (elect_one()) {
      init_bars(bars, 32 * nact);
      mbar_expect_tx(&bars->once, 2 * TILE128);
      tma_load_2d(&tmK, &bars->once, sK, h * DH, b * Lk + k0);
      tma_load_2d(&tmV, &bars->once, sV, h * DH, b * Lk + k0);
    }
    __syncwarp();
    // the ring starts empty: its first stages are produced without waiting; their per-query vectors are requested back
    // to back (three load latencies in parallel, not in series) before any of them is consumed
    float l3[BWD_STAGES], d3[BWD_STAGES];
    uint8_t m3[BWD_STAGES];
    l3[0] = lse_n; d3[0] = delta_n; m3[0] = mq_n;
#pragma unroll
    for (int i = 1; i < BWD_STAGES; ++i) {
      if (i < T) fetch(i);
      l3[i] = lse_n; d3[i] = delta_n; m3[i] = mq_n;
    }
    if (BWD_STAGES < T) fetch(BWD_STAGES);               // leaves (lse_n, delta_n, mq_n) = tile BWD_STAGES for the steady-state loop
    const float l_s = lse_n, d_s = delta_n;
    const uint8_t m_s = mq_n;
#pragma unroll
    for (int i = 0; i < BWD_STAGES; ++i) {
      if (i < T) {
        lse_n = l3[i]; delta_n = d3[i]; mq_n = m3[i];
        produce(i, false);
      }
    }
    lse_n = l_s; delta_n = d_s; mq_n = m_s;
  }
  // the softmax threads' key-mask byte: requested before the TMEM allocation / CTA-wide sync, consumed after it
  const int kj_pre = k0 + (warp & 3) * 32 + lane;
  const uint8_t mk_pre = (warp >= 2 && kj_pre < Lk) ? (blk ? p.mask_k[1] : p.mask_k[0])[(int64_t)b * Lk + kj_pre] : (uint8_t)0;
  if (warp == 2) TRACE(4090);
  const uint32_t tmem = tmem_setup(bars, warp, 128);
  const uint32_t tdPT = tmem + NT, tdK = tmem + 2 * NT, tdV = tmem + 2 * NT + DH;
  if (warp == 2) TRACE(4091);

  if (warp == 0) {
    for (int i = BWD_STAGES; i < T; ++i) produce(i, true);
  } else if (warp == 1) {
    const uint32_t tS = uniform(tmem), tdPTu = uniform(tdPT), tdKu = uniform(tdK), tdVu = uniform(tdV);
    mbar_wait(&bars->once, 0);
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
    const uint32_t aPT = smem_u32(sPT), adST = smem_u32(sdST);
    auto issue_dkv = [&](int u) {
      const int st = u % BWD_STAGES;
      mbar_wait_bg(&bars->p_ready[0], u & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t aQ = smem_u32(sQdO + st * 2 * TILE32), adO = aQ + TILE32;
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tdVu, desc_k64(aPT, k), desc_mn64(adO, k), IDESC_O, (u > 0 || k > 0) ? 1u : 0u);   // dV += P^T dO
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tdKu, desc_k64(adST, k), desc_mn64(aQ, k), IDESC_O, (u > 0 || k > 0) ? 1u : 0u);   // dK += dS^T Q
        umma_commit(&bars->p_free[0]);
        umma_commit(&bars->kv_empty[st]);
      }
      __syncwarp();
    };
    for (int i = 0; i < T; ++i) {
      const int st = i % BWD_STAGES;
      mbar_wait_bg(&bars->kv_full[st], (i / BWD_STAGES) & 1);
      if (i >= 1) mbar_wait_bg(&bars->s_free[0], (i - 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t aQ = smem_u32(sQdO + st * 2 * TILE32), adO = aQ + TILE32;
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tS, desc_k64(aK, k), desc_k64(aQ, k), IDESC_S32, k);        // S^T  = K Q^T
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tdPTu, desc_k64(aV, k), desc_k64(adO, k), IDESC_S32, k);    // dP^T = V dO^T
        umma_commit(&bars->a_ready[0]);
      }
      __syncwarp();
      if (i >= 1) issue_dkv(i - 1);
    }
    issue_dkv(T - 1);
    if (elect_one()) umma_commit(&bars->done);
    __syncwarp();
  } else if ((warp & 3) < nact) {
    const int qd = warp & 3, row = qd * 32 + lane;
    const int kj = k0 + row;
    const bool k_in = kj < Lk;
    const bool mk = k_in ? (mk_pre != 0) : false;
    const bool warp_all_mk = __all_sync(0xffffffffu, mk);
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t a_ready_a = smem_u32(&bars->a_ready[0]), s_free_a = smem_u32(&bars->s_free[0]), p_ready_a = smem_u32(&bars->p_ready[0]);
    const uint32_t p_free_a = smem_u32(&bars->p_free[0]), kv_full_a = smem_u32(&bars->kv_full[0]), qv_a = smem_u32(qv);
    const uint32_t ptrow_a = smem_u32(sPT) + row * 64, dstrow_a = smem_u32(sdST) + row * 64, swz = (row >> 1) & 3;
    int st = 0, st_phase = 0;
    for (int i = 0; i < T; ++i) {
      if (warp == 2) TRACE(i * 8 + 0);
      mbar_wait_a(kv_full_a + st * 8, st_phase);         // acquire the loader's per-query vectors
      if (warp == 2) TRACE(i * 8 + 1);
      const uint32_t qva = qv_a + st * (uint32_t)sizeof(QVec);
      const uint32_t wq = lds_u1(qva + 2 * NT * 4);
      const bool fast = !DROP && warp_all_mk && wq == 0xffffffffu;
      uint32_t kq = 0u;                                  // DROP: bit c = this thread's key is kept for query c of the tile
      if constexpr (DROP) {
        // lane c generates the keep word of (query c, this warp's 32 keys); the transpose hands every thread (= key) its
        // own bit of all 32 words
        const uint32_t Wq = drop_keep_word(lds_u1(qva + QVEC_RH_OFF + lane * 4), attn_group(blk, (k0 >> 5) + qd), p.drop.thr8);
        kq = warp_bit_transpose32(Wq, lane);
      }
      if (++st == BWD_STAGES) { st = 0; st_phase ^= 1; }
      mbar_wait_a(a_ready_a, i & 1);
      tcgen05_fence_after();
      if (warp == 2) TRACE(i * 8 + 2);
      uint32_t pp[16], pd[16];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t rs[16], rp[16];
        tmem_ld_32x32b_x16(tmem + lane_addr + hf * 16, rs);
        tmem_ld_32x32b_x16(tdPT + lane_addr + hf * 16, rp);
        tmem_ld_wait();
        if (hf == 1) {
          tcgen05_fence_before();
          mbar_arrive_a(s_free_a);
        }
        if constexpr (DROP) {
          // column cc = query cc of the tile.  raw logit: valid ? q.k : -10000, then keep ? * drop.scale : 0; only logits that
          // are valid AND kept pass a gradient (a SELECT: the P of the others may be anything)
          const float ds_ = p.drop.scale;
          const float2 sl2 = splat2(p.scale_log2 * ds_), sc2 = splat2(p.scale * ds_), dsc2 = splat2(ds_);
          const uint32_t kqh = kq >> (hf * 16), wqh = wq >> (hf * 16);
          if (warp_all_mk && wq == 0xffffffffu) {         // no masked key / padded query in this tile: one select per operand
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              const float2 nl = lds_f2(qva + (hf * 16 + c) * 4), nd = lds_f2(qva + (NT + hf * 16 + c) * 4);
              const bool k0_ = (kqh >> c) & 1u, k1_ = (kqh >> (c + 1)) & 1u;
              const float s0 = k0_ ? __uint_as_float(rs[c]) : 0.f, s1 = k1_ ? __uint_as_float(rs[c + 1]) : 0.f;
              const float2 pr = ex2_mufu2(fma2(make_float2(s0, s1), sl2, nl));
              const float2 ds = mul2(pr, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, mul2(nd, dsc2)));
              pp[hf * 8 + (c >> 1)] = pack_bf16x2(pr.x, pr.y);
              pd[hf * 8 + (c >> 1)] = pack_bf16x2(k0_ ? ds.x : 0.f, k1_ ? ds.y : 0.f);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              const float2 nl = lds_f2(qva + (hf * 16 + c) * 4), nd = lds_f2(qva + (NT + hf * 16 + c) * 4);
              const bool k0_ = (kqh >> c) & 1u, k1_ = (kqh >> (c + 1)) & 1u;
              const bool v0 = mk && ((wqh >> c) & 1u), v1 = mk && ((wqh >> (c + 1)) & 1u);
              const float s0 = k0_ ? (v0 ? __uint_as_float(rs[c]) : -10000.0f) : 0.f, s1 = k1_ ? (v1 ? __uint_as_float(rs[c + 1]) : -10000.0f) : 0.f;
              const float2 pr = ex2_mufu2(fma2(make_float2(s0, s1), sl2, nl));
              const float2 ds = mul2(pr, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, mul2(nd, dsc2)));
              pp[hf * 8 + (c >> 1)] = pack_bf16x2(pr.x, pr.y);
              pd[hf * 8 + (c >> 1)] = pack_bf16x2((k0_ && v0) ? ds.x : 0.f, (k1_ && v1) ? ds.y : 0.f);
            }
          }
        } else if (fast) {                               // branch hoisted out of the element loops: straight-line FFMA / EX2 code
#pragma unroll
          for (int c = 0; c < 16; c += 4) {
            const float4 nl = lds_f4(qva + (hf * 16 + c) * 4), nd = lds_f4(qva + (NT + hf * 16 + c) * 4);
            const float2 sl2 = splat2(p.scale_log2), sc2 = splat2(p.scale);
            const float2 xa = fma2(make_float2(__uint_as_float(rs[c]), __uint_as_float(rs[c + 1])), sl2, make_float2(nl.x, nl.y));
            const float2 xb = fma2(make_float2(__uint_as_float(rs[c + 2]), __uint_as_float(rs[c + 3])), sl2, make_float2(nl.z, nl.w));
            const float2 pa = poly_pair(c >> 1) ? ex2_poly2(xa) : ex2_mufu2(xa), pb = poly_pair((c >> 1) + 1) ? ex2_poly2(xb) : ex2_mufu2(xb);
            const float2 da = mul2(pa, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, make_float2(nd.x, nd.y)));
            const float2 db = mul2(pb, fma2(make_float2(__uint_as_float(rp[c + 2]), __uint_as_float(rp[c + 3])), sc2, make_float2(nd.z, nd.w)));
            pp[hf * 8 + (c >> 1)] = pack_bf16x2(pa.x, pa.y);
            pp[hf * 8 + (c >> 1) + 1] = pack_bf16x2(pb.x, pb.y);
            pd[hf * 8 + (c >> 1)] = pack_bf16x2(da.x, da.y);
            pd[hf * 8 + (c >> 1) + 1] = pack_bf16x2(db.x, db.y);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; c += 4) {
            const float4 nl = lds_f4(qva + (hf * 16 + c) * 4), nd = lds_f4(qva + (NT + hf * 16 + c) * 4);
            const float nlv[4] = {nl.x, nl.y, nl.z, nl.w}, ndv[4] = {nd.x, nd.y, nd.z, nd.w};
            float pr[4], ds[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int cc = c + t;
              const bool valid = mk && (((wq >> (hf * 16 + cc)) & 1u) != 0);
              const float x = valid ? __uint_as_float(rs[cc]) * p.scale_log2 : p.fill_log2;
              pr[t] = ex2(x + nlv[t]);                   // queries past Lq: nlse2 = -inf => 0
              ds[t] = valid ? pr[t] * fmaf(__uint_as_float(rp[cc]), p.scale, ndv[t]) : 0.f;
            }
            pp[hf * 8 + (c >> 1)] = pack_bf16x2(pr[0], pr[1]);
            pp[hf * 8 + (c >> 1) + 1] = pack_bf16x2(pr[2], pr[3]);
            pd[hf * 8 + (c >> 1)] = pack_bf16x2(ds[0], ds[1]);
            pd[hf * 8 + (c >> 1) + 1] = pack_bf16x2(ds[2], ds[3]);
          }
        }
      }
      if (warp == 2) TRACE(i * 8 + 3);
      if (i >= 1) mbar_wait_a(p_free_a, (i - 1) & 1);    // dK / dV products of tile i-1 have consumed the staging tiles
      write_row_sw64(ptrow_a, swz, pp);
      write_row_sw64(dstrow_a, swz, pd);
      fence_proxy_async_smem();
      mbar_arrive_a(p_ready_a);
      if (warp == 2) TRACE(i * 8 + 4);
    }
    mbar_wait(&bars->done, 0);
    tcgen05_fence_after();
    if (warp == 2) TRACE(4093);
    uint32_t rk[32];
    tmem_ld_32x32(tdK + lane_addr, rk);
    tmem_ld_wait();
    if (k_in && p.dk != nullptr) store_row32_bf16(p.dk + ((int64_t)b * Lk + kj) * p.lddk + h * DH, rk, 1.0f);
    if (p.dbk != nullptr) add_bias_grad(p.dbk + h * DH, rk, k_in, lane);
    tmem_ld_32x32(tdV + lane_addr, rk);
    tmem_ld_wait();
    if (k_in && p.dv != nullptr) store_row32_bf16(p.dv + ((int64_t)b * Lk + kj) * p.lddv + h * DH, rk, 1.0f);
    if (p.dbv != nullptr) add_bias_grad(p.dbv + h * DH, rk, k_in, lane);
    if (warp == 2) TRACE(4094);
  }
  tmem_teardown(tmem, warp, 128);
  if (warp == 2) TRACE(4092);
}

#include "attention_bwd_fused.cuh"
#include "attention_bwd_allkeys.cuh"

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

#ifdef MMI_ATTN_TRACE
extern "C" int mmi_debug_trace(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, tc::g_trace, sizeof(long long) * n);
}
#endif

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
  p.drop = make_drop(a->drop);
  MMI_CHECK_ARG(p.drop.thr8 < 256u, "attn_tc: dropout thr8 must be < 256");
  const bool drop_on = p.drop.thr8 != 0u;
  for (int i = 0; i < a->nblk; ++i) {
    const mmi_attn_block& s = a->blk[i];
    MMI_CHECK_ARG(s.q && s.k && s.v && s.mask_k && s.Lk > 0, "attn_tc: block %d has null pointer / Lk<=0", i);
    MMI_CHECK_ARG(s.ldq % 8 == 0 && s.ldk % 8 == 0 && s.ldv % 8 == 0, "attn_tc: leading dims must be multiples of 8 (TMA)");
    p.Lk[i] = s.Lk; p.mask_k[i] = s.mask_k;
    p.dq[i] = reinterpret_cast<__nv_bfloat16*>(s.dq); p.lddq[i] = s.lddq;
    p.dbq[i] = (kind == 1 || kind == 3 || kind == 4) ? s.dbq : nullptr;
  }
  if (a->nblk == 1) { p.Lk[1] = 0; p.mask_k[1] = p.mask_k[0]; }
  const int64_t q_rows = (int64_t)a->B * a->Lq;
  static size_t cfg_bytes[3] = {0, 0, 0};   // largest dynamic shared-memory size configured so far, per kernel
  const size_t bar_bytes = sizeof(Bars) + 1024 /*align*/;
  static const bool fwd64_env = []() { const char* e = getenv("MMI_ATTN_FWD64"); return e == nullptr || e[0] != '0'; }();
  const bool fwd64 = fwd64_env || drop_on;    // the 32-key forward kernel (A/B only) has no dropout instantiation
  if (kind == 0 || kind == 1) {
    const uint32_t kbox = (kind == 0 && fwd64) ? NT64 : NT;
    CUtensorMap mQ[2], mK[2], mV[2], mdO;
    for (int i = 0; i < 2; ++i) {
      const mmi_attn_block& s = a->blk[i < a->nblk ? i : 0];
      const int64_t k_rows = (int64_t)a->B * s.Lk;
      if (!map_rows(s.q, s.ldq, q_rows, width, QT, &mQ[i])) return MMI_ECUDA;
      if (!map_rows(s.k, s.ldk, k_rows, width, kbox, &mK[i])) return MMI_ECUDA;
      if (!map_rows(s.v, s.ldv, k_rows, width, kbox, &mV[i])) return MMI_ECUDA;
    }
    dim3 grid((a->Lq + QT - 1) / QT, a->H, a->B);
    if (kind == 0 && fwd64) {
      static size_t cfg64 = 0;
      const size_t T = (a->blk[0].Lk + NT64 - 1) / NT64 + (a->nblk > 1 ? (a->blk[1].Lk + NT64 - 1) / NT64 : 0);
      const size_t smem = 2 * TILE128 + 4 * TILE64 + PTILE64 + bar_bytes + 16 + T * 16;
      if (smem > cfg64) {
        int rc = set_smem(attn_fwd_tc64_kernel<false>, smem); if (rc) return rc;
        rc = set_smem(attn_fwd_tc64_kernel<true>, smem); if (rc) return rc;
        cfg64 = smem;
      }
      CUtensorMap mO;
      MMI_CHECK_ARG(a->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "attn_tc fwd: out must be 16-byte aligned with ldo a multiple of 8");
      if (!get_tensor_map_3d(a->out, width, a->Lq, a->B, a->ldo, DH, 32, CU_TENSOR_MAP_SWIZZLE_64B, &mO)) return MMI_ECUDA;
      if (drop_on) attn_fwd_tc64_kernel<true><<<grid, ATT_THREADS, smem, st>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], mO, p);
      else attn_fwd_tc64_kernel<false><<<grid, ATT_THREADS, smem, st>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], mO, p);
    } else if (kind == 0) {
      const size_t T = (a->blk[0].Lk + NT - 1) / NT + (a->nblk > 1 ? (a->blk[1].Lk + NT - 1) / NT : 0);
      const size_t smem = 2 * TILE128 + FWD_STAGES * 2 * TILE32 + 2 * TILE128 + bar_bytes + T * 8;
      if (smem > cfg_bytes[0]) { int rc = set_smem(attn_fwd_tc_kernel, smem); if (rc) return rc; cfg_bytes[0] = smem; }
      attn_fwd_tc_kernel<<<grid, ATT_THREADS, smem, st>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], p);
    } else {
      MMI_CHECK_ARG(a->dout && a->delta, "attn_tc bwd: null dout/delta");
      MMI_CHECK_ARG(a->lddo % 8 == 0, "attn_tc: lddo must be a multiple of 8");
      if (!map_rows(a->dout, a->lddo, q_rows, width, QT, &mdO)) return MMI_ECUDA;
      const size_t T = (a->blk[0].Lk + NT - 1) / NT + (a->nblk > 1 ? (a->blk[1].Lk + NT - 1) / NT : 0);
      const size_t smem = 3 * TILE128 + BWD_STAGES * 2 * TILE32 + 2 * TILE128 + bar_bytes + T * 8;
      if (smem > cfg_bytes[1]) {
        int rc = set_smem(attn_bwd_dq_tc_kernel<false>, smem); if (rc) return rc;
        rc = set_smem(attn_bwd_dq_tc_kernel<true>, smem); if (rc) return rc;
        cfg_bytes[1] = smem;
      }
      if (drop_on) attn_bwd_dq_tc_kernel<true><<<grid, ATT_THREADS, smem, st>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], mdO, p);
      else attn_bwd_dq_tc_kernel<false><<<grid, ATT_THREADS, smem, st>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], mdO, p);
    }
  } else if (kind == 4) {
    // one CTA per (b, h) owning every key of both blocks (attention_bwd_allkeys.cuh)
    MMI_CHECK_ARG(a->dout && a->out && a->lse, "attn_tc all-keys bwd: null dout / out / lse");
    MMI_CHECK_ARG(a->ldo % 8 == 0 && a->lddo % 8 == 0, "attn_tc all-keys bwd: ldo / lddo must be multiples of 8");
    int nt = 0;
    for (int i = 0; i < a->nblk; ++i) {
      nt += (a->blk[i].Lk + QT - 1) / QT;
      MMI_CHECK_ARG(a->blk[i].lddq % 8 == 0, "attn_tc all-keys bwd: lddq must be a multiple of 8");
    }
    if (nt > AK_MAXT) return 1;                 // more keys than one CTA's TMEM holds (5 x 128): the caller uses the other kernels
    CUtensorMap mQ[2], mK[2], mV[2], mdO;
    for (int i = 0; i < 2; ++i) {
      const mmi_attn_block& s = a->blk[i < a->nblk ? i : 0];
      const int64_t k_rows = (int64_t)a->B * s.Lk;
      if (!map_rows(s.q, s.ldq, q_rows, width, QN, &mQ[i])) return MMI_ECUDA;
      if (!map_rows(s.k, s.ldk, k_rows, width, QT, &mK[i])) return MMI_ECUDA;
      if (!map_rows(s.v, s.ldv, k_rows, width, QT, &mV[i])) return MMI_ECUDA;
    }
    if (!map_rows(a->dout, a->lddo, q_rows, width, QN, &mdO)) return MMI_ECUDA;
    const mmi_attn_block& s0 = a->blk[0];
    const mmi_attn_block& s1 = a->blk[a->nblk > 1 ? 1 : 0];
    // persistent: one CTA per SM walks the (b, h) items
    static int sm_count = 0;
    if (sm_count == 0) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0) sm_count = 148;
    }
    const int n_items = a->B * a->H;
    // a CTA count that is a multiple of H keeps one head per CTA (item w -> h = w % H), so the bias-gradient sums it holds in
    // registers leave once per CTA instead of once per item
    int ctas = sm_count >= a->H ? (sm_count / a->H) * a->H : sm_count;
    if (ctas > n_items) ctas = n_items;
    dim3 grid(ctas);
    // dK / dV leave by TMA through 3-D maps [B, Lk, H*dh] (a box never spills into the next batch item's rows)
    CUtensorMap mdK[2], mdV[2];
    int kv_store = 0;
    for (int i = 0; i < 2; ++i) {
      const mmi_attn_block& s = a->blk[i < a->nblk ? i : 0];
      MMI_CHECK_ARG(s.lddk % 8 == 0 && s.lddv % 8 == 0, "attn_tc all-keys bwd: lddk / lddv must be multiples of 8");
      MMI_CHECK_ARG(((reinterpret_cast<uintptr_t>(s.dk) | reinterpret_cast<uintptr_t>(s.dv)) & 15) == 0, "attn_tc all-keys bwd: dk / dv must be 16-byte aligned");
      // (a null gradient pointer: the map is built on k / v and never used)
      if (!get_tensor_map_3d(s.dk ? s.dk : s.k, width, s.Lk, a->B, s.dk ? s.lddk : s.ldk, DH, 32, CU_TENSOR_MAP_SWIZZLE_64B, &mdK[i])) return MMI_ECUDA;
      if (!get_tensor_map_3d(s.dv ? s.dv : s.v, width, s.Lk, a->B, s.dv ? s.lddv : s.ldv, DH, 32, CU_TENSOR_MAP_SWIZZLE_64B, &mdV[i])) return MMI_ECUDA;
      if (i < a->nblk) kv_store |= (s.dk ? 1 : 0) << (2 * i) | (s.dv ? 1 : 0) << (2 * i + 1);
    }
    const size_t smem = AK_MAXT * 2 * TILE128 + 2 * AK_QSTAGE + 2 * STILE + 16 * AK_OUT + 2 * sizeof(QVec64) + sizeof(AKBars) + 1024;
    static bool configured = false;
    if (!configured) {
      int rc = set_smem(attn_bwd_allkeys_tc_kernel<false>, smem); if (rc) return rc;
      rc = set_smem(attn_bwd_allkeys_tc_kernel<true>, smem); if (rc) return rc;
      configured = true;
    }
    if (drop_on)
      attn_bwd_allkeys_tc_kernel<true><<<grid, AK_THREADS, smem, st>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], mdO, mdK[0], mdK[1], mdV[0], mdV[1], p,
                                                                        kv_store, s0.dbk, s1.dbk, s0.dbv, s1.dbv, n_items);
    else
      attn_bwd_allkeys_tc_kernel<false><<<grid, AK_THREADS, smem, st>>>(mQ[0], mQ[1], mK[0], mK[1], mV[0], mV[1], mdO, mdK[0], mdK[1], mdV[0], mdV[1], p,
                                                                        kv_store, s0.dbk, s1.dbk, s0.dbv, s1.dbv, n_items);
  } else if (kind == 3) {
    MMI_CHECK_ARG(which >= 0 && which < a->nblk, "attn_tc fused bwd: bad block index %d", which);
    MMI_CHECK_ARG(a->dout && a->out && a->lse, "attn_tc fused bwd: null dout / out / lse");
    if (a->dq_acc[which] == nullptr || a->dq_count[which] == nullptr) return 1;     // caller did not provide the dQ accumulator
    MMI_CHECK_ARG((reinterpret_cast<uintptr_t>(a->dq_acc[which]) & 15) == 0, "attn_tc fused bwd: dq_acc must be 16-byte aligned");
    MMI_CHECK_ARG(a->ldo % 8 == 0 && a->lddo % 8 == 0, "attn_tc fused bwd: ldo / lddo must be multiples of 8");
    const mmi_attn_block& s = a->blk[which];
    MMI_CHECK_ARG(s.lddq % 4 == 0, "attn_tc fused bwd: lddq must be a multiple of 4");
    p.which = which;
    p.dk = reinterpret_cast<__nv_bfloat16*>(s.dk); p.lddk = s.lddk;
    p.dv = reinterpret_cast<__nv_bfloat16*>(s.dv); p.lddv = s.lddv;
    p.dbk = s.dbk; p.dbv = s.dbv;
    const int64_t k_rows = (int64_t)a->B * s.Lk;
    CUtensorMap mQ, mK, mV, mdO;
    if (!map_rows(s.q, s.ldq, q_rows, width, QN, &mQ)) return MMI_ECUDA;
    if (!map_rows(a->dout, a->lddo, q_rows, width, QN, &mdO)) return MMI_ECUDA;
    if (!map_rows(s.k, s.ldk, k_rows, width, QT, &mK)) return MMI_ECUDA;
    if (!map_rows(s.v, s.ldv, k_rows, width, QT, &mV)) return MMI_ECUDA;
    dim3 grid((s.Lk + QT - 1) / QT, a->H, a->B);
    const size_t smem = 2 * TILE128 + FB_STAGES * 2 * TILE64Q + 2 * STILE + FB_STAGES * sizeof(QVec64) + sizeof(FBars) + 1024;
    static bool configured = false;
    if (!configured) {
      int rc = set_smem(attn_bwd_fused_tc_kernel<false>, smem); if (rc) return rc;
      rc = set_smem(attn_bwd_fused_tc_kernel<true>, smem); if (rc) return rc;
      configured = true;
    }
    if (drop_on) attn_bwd_fused_tc_kernel<true><<<grid, FB_THREADS, smem, st>>>(mQ, mK, mV, mdO, p, a->dq_acc[which], a->dq_count[which]);
    else attn_bwd_fused_tc_kernel<false><<<grid, FB_THREADS, smem, st>>>(mQ, mK, mV, mdO, p, a->dq_acc[which], a->dq_count[which]);
  } else {
    MMI_CHECK_ARG(which >= 0 && which < a->nblk, "attn_tc dkv: bad block index %d", which);
    MMI_CHECK_ARG(a->dout && a->delta, "attn_tc bwd: null dout/delta");
    const mmi_attn_block& s = a->blk[which];
    p.which = which;
    p.dk = reinterpret_cast<__nv_bfloat16*>(s.dk); p.lddk = s.lddk;
    p.dv = reinterpret_cast<__nv_bfloat16*>(s.dv); p.lddv = s.lddv;
    p.dbk = s.dbk; p.dbv = s.dbv;
    const int64_t k_rows = (int64_t)a->B * s.Lk;
    CUtensorMap mQ, mK, mV, mdO;
    if (!map_rows(s.q, s.ldq, q_rows, width, NT, &mQ)) return MMI_ECUDA;
    if (!map_rows(a->dout, a->lddo, q_rows, width, NT, &mdO)) return MMI_ECUDA;
    if (!map_rows(s.k, s.ldk, k_rows, width, QT, &mK)) return MMI_ECUDA;
    if (!map_rows(s.v, s.ldv, k_rows, width, QT, &mV)) return MMI_ECUDA;
    dim3 grid((s.Lk + QT - 1) / QT, a->H, a->B);
    const size_t smem = 2 * TILE128 + 4 * 2 * TILE32 + 2 * TILE128 + BWD_STAGES * sizeof(QVec) + bar_bytes;
    if (smem > cfg_bytes[2]) {
      int rc = set_smem(attn_bwd_dkv_tc_kernel<false>, smem); if (rc) return rc;
      rc = set_smem(attn_bwd_dkv_tc_kernel<true>, smem); if (rc) return rc;
      cfg_bytes[2] = smem;
    }
    if (drop_on) attn_bwd_dkv_tc_kernel<true><<<grid, ATT_THREADS, smem, st>>>(mQ, mK, mV, mdO, p);
    else attn_bwd_dkv_tc_kernel<false><<<grid, ATT_THREADS, smem, st>>>(mQ, mK, mV, mdO, p);
  }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

}  // namespace mmi
