// MMI_IMPL_TC: bf16 GEMM on the 5th-gen tensor cores (sm_100a).
//
//   * operands staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a 4-stage
//     shared-memory ring, mbarrier full/empty handshake;
//   * tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction, issued by ONE
//     thread, fp32 accumulators in TMEM (2 x BN columns: the epilogue of tile i overlaps the
//     main loop of tile i+1);
//   * epilogue warps read TMEM with tcgen05.ld (32 lanes x 32 columns per warp) and apply
//     the fused epilogue of mmi_gemm (bias, GELU + saved pre-activation, GELU' multiply,
//     residual / position-embedding add, fp32 accumulate for weight gradients);
//   * TMA epilogue (bf16 outputs): thread = accumulator row; the residual / GELU' tile of the
//     NEXT output tile is prefetched by TMA into the warp's staging buffers while the tensor
//     core works, the result is written in place (SWIZZLE_128B, conflict-free 16 B accesses) and
//     leaves with one cp.async.bulk.tensor store per 32 x 64 block -- no per-row global
//     latency in the epilogue, so the kernel stays MMA-bound at K = 512;
//   * persistent CTAs (one per SM), tile order n-fastest so an activation row-block is
//     fetched from HBM once and re-read from L2; split-K (atomic fp32) for the
//     weight-gradient GEMMs whose K dimension is the token count.
//
// Layouts: NT uses K-major operands (A [M,K], B [N,K]); TN (weight gradient, dW = dY^T X)
// uses MN-major operands straight from the row-major activations -- no transposes in HBM.
// NN is not needed: the engine keeps a transposed bf16 shadow of each weight for dgrad.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "tc_common.cuh"
#include <string.h>

namespace mmi {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
// epilogue warps: warp 0 = TMA, warp 1 = MMA + TMEM alloc, then 8 (register epilogue: two per TMEM lane quarter, each half of the
// columns) or 16 (TMA epilogue: four per lane quarter, 32-column chunks -- with eight, two warps per scheduler could not hide
// the tcgen05.ld / MUFU / shared-memory latencies of the GELU and dropout epilogues: 40 % issue utilisation, 0.2-0.4 of peak)
__host__ __device__ constexpr int epi_warps(bool tma_epi) { return tma_epi ? 16 : 8; }
__host__ __device__ constexpr int num_threads(bool tma_epi) { return 64 + 32 * epi_warps(tma_epi); }
constexpr int EPI_WARPS = 8;     // register epilogue
constexpr int STG_COLS = 64;     // epilogue staging: 32 rows x 64 fp32 per warp (8 KB), XOR-swizzled
constexpr int STG_BYTES = 32 * STG_COLS * 4;
__host__ __device__ constexpr int stages_for(int bn) { return bn == 256 ? 3 : 4; }
struct TileInfo {
  int m_blk, n_blk, kb0, kb1;
  bool lead;
};

__device__ __forceinline__ TileInfo get_tile(int t, int m_tiles, int n_tiles, int num_kb, int split) {
  const int per = m_tiles * n_tiles;
  const int s = t / per, r = t % per;
  const int kb_per = (num_kb + split - 1) / split;
  TileInfo ti;
  ti.m_blk = r / n_tiles;
  ti.n_blk = r % n_tiles;
  ti.kb0 = s * kb_per;
  ti.kb1 = min(num_kb, ti.kb0 + kb_per);
  ti.lead = s == 0;
  return ti;
}

struct EpiMaps {
  CUtensorMap c, aux, c2;   // output, prefetched add / mul operand, second output (saved pre-activation)
};

constexpr int ECOLS = 32;                       // TMA epilogue: columns per chunk
constexpr int EBUF_BYTES = 32 * ECOLS * 2;      // 32 rows x 32 bf16 (one SWIZZLE_64B box per warp and chunk)

// DROP (TMA epilogue only): dropout of act(z) fused into the epilogue.  The no-dropout instantiation is byte-for-byte the
// kernel that existed before dropout was added; the generic register epilogue handles dropout at run time.
// TAUX: element type of the side operands (preact / mul_gelu_grad) of the register epilogue: bf16, or float for the
// split-fp32 mode (three bf16 terms per operand, six segment products: see GemmParams::split_terms).
template <int BN, bool MN_MAJOR, typename TOUT, bool TMA_EPI, bool DROP = false, typename TAUX = __nv_bfloat16>
__global__ void __launch_bounds__(num_threads(TMA_EPI), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ EpiMaps em, const GemmParams p, int m_tiles, int n_tiles, int num_kb, int split) {
  extern __shared__ uint8_t smem_raw[];
  // split-fp32 mode (TAUX = float): k-block range [kb0, kb1) of a tile is PER TERM; the loops walk the six (A term, B term)
  // pairs over it.  hi*hi goes to the tile's main accumulator, the five small products to a second one (the tensor core
  // truncates every addend to the accumulator's ulp: ~0.5 ulp per MMA step towards zero -- the small terms must not
  // spend steps on the main accumulator, and split-K slices keep the hi*hi chain short); the epilogue adds the two.
  constexpr bool SPLIT = sizeof(TAUX) == 4;
  constexpr uint32_t TMEM_COLS = SPLIT ? 4 * BN : 2 * BN;
  static_assert(!SPLIT || BN == 128, "split-fp32 mode: two accumulators per tile need BN = 128");
  constexpr int STAGES = stages_for(BN);
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // SWIZZLE_128B atoms must sit on 1024 B boundaries
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  constexpr int EW = epi_warps(TMA_EPI);
  uint64_t* aux_bars = tmem_empty + 2;        // [EW][2]  (TMA epilogue)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(aux_bars + 2 * EW);
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES + 256;   // 1024-aligned: STAGE_BYTES is a multiple of 1024

  const int warp = (int)uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;   // warp index in a uniform register
  const int total_tiles = m_tiles * n_tiles * split;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 32 * EW); }
    for (int b = 0; b < 2 * EW; ++b) mbar_init(&aux_bars[b], 1);
    if constexpr (TMA_EPI) { tma_prefetch_desc(&em.c); tma_prefetch_desc(&em.aux); tma_prefetch_desc(&em.c2); }
    fence_barrier_init();
  }
  if (warp == 1) {  // one warp owns TMEM alloc + dealloc; 2 accumulator buffers of BN fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    // warp-uniform loop: every lane waits on the mbarrier, one elected lane issues (operands stay in uniform
    // registers; a lane-0-only loop makes ptxas wrap every UTMALDG / UTCHMMA in a waterfall loop)
    uint32_t stage = 0, phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileInfo ti = get_tile(t, m_tiles, n_tiles, num_kb, split);
      const int len = ti.kb1 - ti.kb0, n_it = SPLIT ? 6 * len : len;
      for (int i2 = 0; i2 < n_it; ++i2) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        // split-fp32 mode: pair sg = i2 / len of A terms (hi, hi, mid, mid, hi, lo) x B terms (hi, mid, hi, mid, lo, hi);
        // term t of an operand is the block of seg_kb k-blocks starting at t * seg_kb
        int ka = ti.kb0 + i2, kbb = ka;
        if constexpr (SPLIT) {
          const int sg = i2 / len, r = i2 - sg * len;
          ka = (int)((0x201100u >> (4 * sg)) & 0xfu) * p.seg_kb + ti.kb0 + r;
          kbb = (int)((0x021010u >> (4 * sg)) & 0xfu) * p.seg_kb + ti.kb0 + r;
        }
        if (elect_one()) {
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if constexpr (!MN_MAJOR) {
            tma_load_2d(&tma_a, &full_bar[stage], sa, ka * BK, ti.m_blk * BM);    // box {64 k, 128 rows}
            tma_load_2d(&tma_b, &full_bar[stage], sb, kbb * BK, ti.n_blk * BN);   // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)                                    // boxes {64 mn, 64 k}
              tma_load_2d(&tma_a, &full_bar[stage], sa + j * (64 * BK * 2), ti.m_blk * BM + j * 64, ka * BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(&tma_b, &full_bar[stage], sb + j * (64 * BK * 2), ti.n_blk * BN + j * 64, kbb * BK);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform, elected lane issues)
    constexpr uint32_t idesc = make_idesc(BM, BN, MN_MAJOR, MN_MAJOR);
    const uint32_t tbase = uniform(tmem_base);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const TileInfo ti = get_tile(t, m_tiles, n_tiles, num_kb, split);
      const uint32_t buf = it & 1, use = it >> 1;
      mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);   // epilogue drained this accumulator
      tcgen05_fence_after();
      const uint32_t d_main = tbase + buf * (SPLIT ? 2 * BN : BN);
      const int len = ti.kb1 - ti.kb0, n_it = SPLIT ? 6 * len : len;
      for (int i2 = 0; i2 < n_it; ++i2) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        // (split mode: pair 0 = hi*hi -> main accumulator; pairs 1..5 -> the second accumulator, BN columns further)
        const bool small = SPLIT && i2 >= len;
        const uint32_t d_tmem = uniform(small ? d_main + BN : d_main);
        const uint32_t first = uniform((i2 == 0 || (SPLIT && i2 == len)) ? 1u : 0u);
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            uint64_t ad, bd;
            if constexpr (!MN_MAJOR) {
              ad = make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
              bd = make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
            } else {
              ad = make_smem_desc(sa + k * (UMMA_K * 128), 64 * BK * 2, 1024);
              bd = make_smem_desc(sb + k * (UMMA_K * 128), 64 * BK * 2, 1024);
            }
            umma_f16(d_tmem, ad, bd, idesc, (!first || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);             // smem slot free once these MMAs retire
          if (i2 == n_it - 1) umma_commit(&tmem_full[buf]);   // accumulator(s) complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if constexpr (TMA_EPI) {
    // ===================================================================== TMA epilogue (16 warps)
    // warp = 32 accumulator rows (its TMEM lane quarter) x every fourth 32-column chunk of the tile.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int cq = ew >> 2;
    constexpr int CPW = BN / 128;                      // chunks per warp per tile
    uint8_t* epi_base = smem + STAGES * STAGE_BYTES + 1024;            // 1024-aligned (SWIZZLE_128B); barriers sit below
    uint8_t* ebuf = epi_base + ew * (2 * EBUF_BYTES);
    float* sbias = reinterpret_cast<float*>(epi_base + EW * 2 * EBUF_BYTES) + ew * ECOLS;
    uint64_t* aux_bar = aux_bars + 2 * ew;
    // position-embedding add (fp32 table [add_mod, N], row = m mod add_mod): read straight from L1/L2 by the thread
    // that owns the accumulator row -- the table is ~1 MB and shared by every batch item, so it never leaves cache
    const bool pe_add = p.add != nullptr && p.add_mod < p.M;
    const bool aux_add = p.add != nullptr && !pe_add, aux_mul = p.mul_gelu_grad != nullptr;
    const bool has_aux = aux_add || aux_mul;
    const bool second = p.preact != nullptr;           // host: never together with an aux operand
    const bool do_gelu = p.act == MMI_ACT_GELU, do_relu = p.act == MMI_ACT_RELU;
    constexpr bool drop_on = DROP;
    // Linear -> dropout -> (+ residual): the survivors' scale rides in the bias add (x = acc * s + b * s), the mask is a
    // select on bits of the row's keep words (compile-time bit positions: ptxas turns a byte of them into one R2P)
    const float dscale = (drop_on && !do_gelu) ? p.drop.scale : 1.0f;
    auto issue_aux = [&](int t) {                      // lane 0 only
      const TileInfo ti = get_tile(t, m_tiles, n_tiles, num_kb, split);
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        mbar_expect_tx(&aux_bar[k], EBUF_BYTES);
        tma_load_2d(&em.aux, &aux_bar[k], ebuf + k * EBUF_BYTES, ti.n_blk * BN + (cq + 4 * k) * ECOLS, ti.m_blk * BM + q * 32);
      }
    };
    if (has_aux && (int)blockIdx.x < total_tiles && elect_one()) issue_aux(blockIdx.x);
    __syncwarp();
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const TileInfo ti = get_tile(t, m_tiles, n_tiles, num_kb, split);
      const uint32_t buf = it & 1, use = it >> 1;
      const int m0 = ti.m_blk * BM + q * 32;
      const uint32_t rowh = drop_on ? drop_rowhash(p.drop.key, (uint64_t)(m0 + lane)) : 0u;    // overlaps the wait below
      mbar_wait(&tmem_full[buf], use & 1);
      tcgen05_fence_after();
      if (!has_aux) {                                  // last tile's stores have finished reading the staging buffers
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
      }
#pragma unroll 1
      for (int k = 0; k < CPW; ++k) {
        const int n0 = ti.n_blk * BN + (cq + 4 * k) * ECOLS;
        uint32_t acc[ECOLS];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + (cq + 4 * k) * ECOLS, acc);
        tmem_ld_wait();
        if (k == CPW - 1) {                            // accumulator fully read: hand it back to the MMA warp
          tcgen05_fence_before();
          mbar_arrive(&tmem_empty[buf]);
        }
        if (second && k > 0) {                         // both staging buffers are reused within the tile
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
        {
          const int n = n0 + lane;
          float b = 0.f;
          if (p.bias != nullptr && n < p.N) b = p.bias[n];
          if constexpr (drop_on) b *= dscale;
          sbias[lane] = b;
        }
        __syncwarp();
        uint8_t* tile1 = ebuf + (second ? 0 : k * EBUF_BYTES);          // warp-uniform staging tiles
        uint8_t* tile2 = ebuf + EBUF_BYTES;
        uint8_t* row1 = tile1 + lane * (ECOLS * 2);
        uint8_t* row2 = tile2 + lane * (ECOLS * 2);
        if (has_aux) mbar_wait(&aux_bar[k], it & 1);
        const float* pe_row = pe_add ? reinterpret_cast<const float*>(p.add) + ((int64_t)(m0 + lane) % p.add_mod) * p.ld_add + n0 : nullptr;
        // dropout: the keep word of this row's 32 columns (dropout.cuh); bit c = column n0 + c survives
        uint32_t kw0 = 0xffffffffu;
        if constexpr (drop_on) kw0 = drop_keep_word(rowh, (uint32_t)(n0 >> 5), p.drop.thr8);
#pragma unroll
        for (int v = 0; v < ECOLS / 8; ++v) {
          const int phys = (v ^ ((lane >> 1) & 3)) << 4;                   // SWIZZLE_64B: 16-byte chunk ^ row bits 1..2
          const uint32_t kb8 = kw0 >> (8 * v);                             // bits 0..7: this vector's columns
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(sbias + 8 * v + j);
            if constexpr (drop_on) {
              x[j] = fmaf(__uint_as_float(acc[8 * v + j]), dscale, b.x);
              x[j + 1] = fmaf(__uint_as_float(acc[8 * v + j + 1]), dscale, b.y);
              x[j + 2] = fmaf(__uint_as_float(acc[8 * v + j + 2]), dscale, b.z);
              x[j + 3] = fmaf(__uint_as_float(acc[8 * v + j + 3]), dscale, b.w);
            } else {
              x[j] = __uint_as_float(acc[8 * v + j]) + b.x;
              x[j + 1] = __uint_as_float(acc[8 * v + j + 1]) + b.y;
              x[j + 2] = __uint_as_float(acc[8 * v + j + 2]) + b.z;
              x[j + 3] = __uint_as_float(acc[8 * v + j + 3]) + b.w;
            }
          }
          if (pe_add) {
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(pe_row + 8 * v + j));
              x[j] += a.x; x[j + 1] += a.y; x[j + 2] += a.z; x[j + 3] += a.w;
            }
          }
          if (do_relu) {                                 // relu(z) (times the survivors' scale, already folded into x)
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          if (drop_on && !do_gelu) {                     // dropout(x W^T + b) BEFORE the residual is added
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = ((kb8 >> j) & 1u) ? x[j] : 0.f;
          }
          if (has_aux) {
            const uint4 a = *reinterpret_cast<const uint4*>(row1 + phys);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a0 = __uint_as_float(aw[j] << 16), a1 = __uint_as_float(aw[j] & 0xffff0000u);
              if (aux_add) { x[2 * j] += a0; x[2 * j + 1] += a1; }
              else if (p.mul_is_grad == 2) { x[2 * j] = a0 > 0.f ? x[2 * j] * p.mul_scale : 0.f; x[2 * j + 1] = a1 > 0.f ? x[2 * j + 1] * p.mul_scale : 0.f; }
              else if (p.mul_is_grad) { x[2 * j] *= a0; x[2 * j + 1] *= a1; }
              else { float2 gg, dg; gelu_pair<false, true>(make_float2(a0, a1), gg, dg); x[2 * j] *= dg.x; x[2 * j + 1] *= dg.y; }
            }
          }
          if (second) {
            float z[8];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              z[j] = x[j]; z[j + 1] = x[j + 1];
              if (do_gelu || p.save_act_grad) {
                float2 gg, dg;
                gelu_pair<true, true>(make_float2(x[j], x[j + 1]), gg, dg);
                if (p.save_act_grad) { z[j] = dg.x; z[j + 1] = dg.y; }
                if (do_gelu) { x[j] = gg.x; x[j + 1] = gg.y; }
              }
              if (drop_on && do_gelu) {                  // dropout(gelu(z)); the saved gelu'(z) carries the same factor
                const float2 ds2 = splat2(p.drop.scale);
                const float2 xs = mul2(make_float2(x[j], x[j + 1]), ds2);
                x[j] = ((kb8 >> j) & 1u) ? xs.x : 0.f;
                x[j + 1] = ((kb8 >> (j + 1)) & 1u) ? xs.y : 0.f;
                if (p.save_act_grad) {
                  const float2 zs = mul2(make_float2(z[j], z[j + 1]), ds2);
                  z[j] = ((kb8 >> j) & 1u) ? zs.x : 0.f;
                  z[j + 1] = ((kb8 >> (j + 1)) & 1u) ? zs.y : 0.f;
                }
              }
            }
            *reinterpret_cast<uint4*>(row2 + phys) = make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]), pack_bf16x2(z[6], z[7]));
          } else if (do_gelu) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              float2 gg, dg;
              gelu_pair<true, false>(make_float2(x[j], x[j + 1]), gg, dg);
              if (drop_on) gg = mul2(gg, splat2(p.drop.scale));
              x[j] = (drop_on && !((kb8 >> j) & 1u)) ? 0.f : gg.x;
              x[j + 1] = (drop_on && !((kb8 >> (j + 1)) & 1u)) ? 0.f : gg.y;
            }
          }
          *reinterpret_cast<uint4*>(row1 + phys) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {                               // lane 0: owns this warp's bulk async-groups
          tma_store_2d(&em.c, tile1, n0, m0);
          if (second) tma_store_2d(&em.c2, tile2, n0, m0);
          bulk_commit();
        }
      }
      if (has_aux) {                                   // recycle the buffers: prefetch the next tile's operand
        if (elect_one()) {
          bulk_wait_read0();
          if (t + (int)gridDim.x < total_tiles) issue_aux(t + gridDim.x);
        }
        __syncwarp();
      }
    }
    if (lane == 0) bulk_wait0();
    __syncwarp();
  } else {
    // ===================================================================== epilogue (8 warps)
    // TMEM -> registers (thread = accumulator row) -> swizzled smem -> registers (lane = column pair
    // of one row) -> fused epilogue with fully coalesced global reads / writes.
    const int ew = warp - 2;
    const int q = warp & 3;                            // TMEM lane quarter this warp may read
    const int half = ew >> 2;                          // which half of the 64-column chunks
    float* stg = reinterpret_cast<float*>(stg_base + ew * STG_BYTES);
    constexpr int CHUNKS = BN / STG_COLS;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const TileInfo ti = get_tile(t, m_tiles, n_tiles, num_kb, split);
      const uint32_t buf = it & 1, use = it >> 1;
      mbar_wait(&tmem_full[buf], use & 1);
      tcgen05_fence_after();
      const int64_t m_base = static_cast<int64_t>(ti.m_blk) * BM + q * 32;
      const int64_t rem_rows = p.M - m_base;
      const int rows = rem_rows >= 32 ? 32 : (rem_rows > 0 ? (int)rem_rows : 0);   // warp-uniform
#pragma unroll 1
      for (int cc = half; cc < CHUNKS; cc += 2) {
        {
          uint32_t r0[32], r1[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * (SPLIT ? 2 * BN : BN) + cc * STG_COLS;
          tmem_ld_32x32(taddr, r0);
          tmem_ld_32x32(taddr + 32, r1);
          tmem_ld_wait();
          if constexpr (SPLIT) {                           // + the five small products (round-to-nearest adds)
            uint32_t s0[32];
            tmem_ld_32x32(taddr + BN, s0);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) r0[c] = __float_as_uint(__uint_as_float(r0[c]) + __uint_as_float(s0[c]));
            tmem_ld_32x32(taddr + BN + 32, s0);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) r1[c] = __float_as_uint(__uint_as_float(r1[c]) + __uint_as_float(s0[c]));
          }
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            *reinterpret_cast<uint4*>(stg + lane * STG_COLS + ((v ^ (lane & 7)) << 2)) = make_uint4(r0[4 * v], r0[4 * v + 1], r0[4 * v + 2], r0[4 * v + 3]);
            *reinterpret_cast<uint4*>(stg + lane * STG_COLS + (((v + 8) ^ (lane & 7)) << 2)) = make_uint4(r1[4 * v], r1[4 * v + 1], r1[4 * v + 2], r1[4 * v + 3]);
          }
        }
        __syncwarp();
        const int64_t n = static_cast<int64_t>(ti.n_blk) * BN + cc * STG_COLS + 2 * lane;
        if (n < p.N) {
          float b0 = 0.f, b1 = 0.f;
          if (ti.lead && p.bias != nullptr) { const float2 b = *reinterpret_cast<const float2*>(p.bias + n); b0 = b.x; b1 = b.y; }
          int64_t add_row = (p.add != nullptr && m_base >= p.add_mod) ? m_base % p.add_mod : m_base;   // once per 32-row block
#pragma unroll 4
          for (int rr = 0; rr < rows; ++rr) {
            const float2 x = *reinterpret_cast<const float2*>(stg + rr * STG_COLS + (((lane >> 1) ^ (rr & 7)) << 2) + ((lane & 1) << 1));
            gemm_epilogue2<TAUX, TOUT>(p, m_base + rr, n, x.x, x.y, b0, b1, ti.lead, add_row);
            if (++add_row == p.add_mod) add_row = 0;
          }
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
template <int BN, bool TMA_EPI>
constexpr size_t smem_bytes() {
  return stages_for(BN) * (BM * BK * 2 + BN * BK * 2) + 1024 /*align*/ +
         (TMA_EPI ? 1024 /*barriers*/ + epi_warps(true) * 2 * EBUF_BYTES + epi_warps(true) * ECOLS * 4 : 256 /*barriers*/ + EPI_WARPS * STG_BYTES);
}

template <int BN, bool MN_MAJOR, typename TOUT, bool TMA_EPI, bool DROP = false, typename TAUX = __nv_bfloat16>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const EpiMaps& em, const GemmParams& p, int m_tiles, int n_tiles,
                  int num_kb, int split, cudaStream_t st) {
  constexpr size_t smem = smem_bytes<BN, TMA_EPI>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, MN_MAJOR, TOUT, TMA_EPI, DROP, TAUX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return MMI_ECUDA; }
    configured = true;
  }
  const int total = m_tiles * n_tiles * split;
  const int grid = total < kNumSMs ? total : kNumSMs;
  gemm_tc_kernel<BN, MN_MAJOR, TOUT, TMA_EPI, DROP, TAUX><<<grid, num_threads(TMA_EPI), smem, st>>>(ta, tb, em, p, m_tiles, n_tiles, num_kb, split);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

static bool tma_ok(const void* ptr, int64_t ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0; }

// ------------------------------------------------------------------------------ fp32 operands: three-term bf16 split
// x = hi + mid + lo with hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid): 3 x 8 mantissa bits, exact for every
// fp32 x in bf16's exponent range.  src [R, C] fp32 (row stride ld) -> three bf16 copies in dst:
//   term_stride elements apart, each [Rp, ldd] with rows >= R and columns >= C (up to Cp) zero-filled.
__device__ __forceinline__ void split3(float x, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, int64_t ld, int64_t R, int64_t C, __nv_bfloat16* __restrict__ dst,
                                                     int64_t ldd, int64_t Rp, int64_t Cp, int64_t term_stride) {
  const int64_t cq = Cp / 4;                 // Cp is a multiple of 4
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Rp * cq; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cq, c = (i - r * cq) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < R) {
      if (c + 3 < C && (ld & 3) == 0) { const float4 t = *reinterpret_cast<const float4*>(src + r * ld + c); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
      else { for (int j = 0; j < 4; ++j) if (c + j < C) v[j] = src[r * ld + c + j]; }
    }
    __nv_bfloat16 h[4], m[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split3(v[j], h[j], m[j], l[j]);
    __nv_bfloat16* d = dst + r * ldd + c;
    *reinterpret_cast<uint2*>(d) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(d + term_stride) = *reinterpret_cast<uint2*>(m);
    *reinterpret_cast<uint2*>(d + 2 * term_stride) = *reinterpret_cast<uint2*>(l);
  }
}
// transposing variant (the [K, N] operand of an NN product): src [R, C] -> terms [C, 3 x Rp] (term t at column t * Rp),
// columns r >= R zero-filled up to Rp.  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) split3_transpose_kernel(const float* __restrict__ src, int64_t ld, int64_t R, int64_t C,
                                                               __nv_bfloat16* __restrict__ dst, int64_t ldd, int64_t Rp) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int64_t r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < C) ? src[r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t c = c0 + j, r = r0 + tx;      // output row = source column
    if (c < C && r < Rp) {
      __nv_bfloat16 h, m, l;
      split3(tile[tx][j], h, m, l);
      dst[c * ldd + r] = h; dst[c * ldd + Rp + r] = m; dst[c * ldd + 2 * Rp + r] = l;
    }
  }
}

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace tc

int64_t gemm_split_workspace(int layout, int64_t M, int64_t N, int64_t K) {
  const int64_t Kp = tc::round_up(K, tc::BK);
  if (layout == MMI_GEMM_TN) return 3 * Kp * (tc::round_up(M, 8) + tc::round_up(N, 8)) * 2 + 2048;
  return (M + N) * 3 * Kp * 2 + 2048;
}

bool tc_available() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    ok = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
        major == 10 && tc::encode_available())
      ok = 1;
    cudaGetLastError();
  }
  return ok == 1;
}

int gemm_tc(const GemmParams& p_in, cudaStream_t st) {
  using namespace tc;
  GemmParams p = p_in;
  if (p.in_dtype == MMI_F32) {
    // split-fp32 mode: three bf16 terms per operand in the caller's workspace, then the bf16 kernel over six segment pairs
    MMI_CHECK_ARG(p.out_dtype == MMI_F32, "gemm_tc: fp32 operands need an fp32 C");
    MMI_CHECK_ARG(p.split_ws != nullptr && p.split_ws_bytes >= gemm_split_workspace(p.layout, p.M, p.N, p.K),
                  "gemm_tc: fp32 operands need split_ws of mmi_gemm_split_workspace() bytes");
    const int64_t Kp = round_up(p.K, BK);
    __nv_bfloat16* wa = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(p.split_ws) + 1023) & ~static_cast<uintptr_t>(1023));
    const float* A = reinterpret_cast<const float*>(p.A);
    const float* B = reinterpret_cast<const float*>(p.B);
    auto blocks = [](int64_t n) { const int64_t b = (n + 255) / 256; return (unsigned)(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16); };
    __nv_bfloat16* wb;
    if (p.layout == MMI_GEMM_TN) {      // A [K, M], B [K, N]: the terms are stacked along the rows (k)
      const int64_t lda3 = round_up(p.M, 8), ldb3 = round_up(p.N, 8);
      wb = wa + 3 * Kp * lda3;
      split3_kernel<<<blocks(Kp * lda3 / 4), 256, 0, st>>>(A, p.lda, p.K, p.M, wa, lda3, Kp, lda3, Kp * lda3);
      split3_kernel<<<blocks(Kp * ldb3 / 4), 256, 0, st>>>(B, p.ldb, p.K, p.N, wb, ldb3, Kp, ldb3, Kp * ldb3);
      p.lda = lda3; p.ldb = ldb3;
    } else {                             // NT: A [M, K], B [N, K]; NN: B [K, N] is transposed on the way
      wb = wa + p.M * 3 * Kp;
      split3_kernel<<<blocks(p.M * Kp / 4), 256, 0, st>>>(A, p.lda, p.M, p.K, wa, 3 * Kp, p.M, Kp, Kp);
      if (p.layout == MMI_GEMM_NT) split3_kernel<<<blocks(p.N * Kp / 4), 256, 0, st>>>(B, p.ldb, p.N, p.K, wb, 3 * Kp, p.N, Kp, Kp);
      else {
        dim3 g((unsigned)((p.N + 31) / 32), (unsigned)(Kp / 32));
        split3_transpose_kernel<<<g, 256, 0, st>>>(B, p.ldb, p.K, p.N, wb, 3 * Kp, Kp);
      }
      p.lda = 3 * Kp; p.ldb = 3 * Kp;
      p.layout = MMI_GEMM_NT;
    }
    MMI_CHECK_LAUNCH();
    p.A = wa; p.B = wb;
    p.K = Kp;                            // one term; the kernel walks 6 x (Kp / 64) k-blocks
    p.split_terms = 6; p.seg_kb = (int)(Kp / BK);
  }
  MMI_CHECK_ARG(p.in_dtype == MMI_BF16 || p.split_terms, "gemm_tc: inputs must be bf16 (or fp32 with a split workspace)");
  MMI_CHECK_ARG(p.layout == MMI_GEMM_NT || p.layout == MMI_GEMM_TN, "gemm_tc: layouts NT and TN only (dgrad uses the transposed weight shadow)");
  MMI_CHECK_ARG((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.B) & 15) == 0, "gemm_tc: A/B must be 16-byte aligned");
  MMI_CHECK_ARG(p.lda % 8 == 0 && p.ldb % 8 == 0, "gemm_tc: lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
  const bool mn = p.layout == MMI_GEMM_TN;
  const int bn = (p.N % 256 == 0 && !p.split_terms) ? 256 : 128;   // split-fp32 mode: two accumulators per tile, BN = 128
  const int m_tiles = (int)((p.M + BM - 1) / BM), n_tiles = (int)((p.N + bn - 1) / bn);
  const int num_kb = (int)((p.K + BK - 1) / BK);     // (split-fp32 mode: per term)
  int split = p.split_k;
  if (p.split_terms && p.accumulate) {
    // weight gradients (K = tokens): slices of <= 8 k-blocks per term keep the hi*hi chain at 32 MMA steps; the slices meet
    // in C through round-to-nearest atomics
    split = (num_kb + 7) / 8;
    if (split > 4096) split = 4096;
  }
  if (split <= 0) {  // auto: fill the 148 SMs
    split = kNumSMs / (m_tiles * n_tiles);
    if (split < 1) split = 1;
  }
  if (split > num_kb) split = num_kb;
  {  // no empty split slices
    const int kb_per = (num_kb + split - 1) / split;
    split = (num_kb + kb_per - 1) / kb_per;
  }
  MMI_CHECK_ARG(split == 1 || p.accumulate, "gemm_tc: split-K needs accumulate=1");
  p.split_k = split;
  CUtensorMap ta, tb;
  if (!mn) {
    const uint64_t kext = (uint64_t)p.K * (p.split_terms ? 3 : 1);       // split mode: the three terms side by side along k
    if (!get_tensor_map(p.A, kext, (uint64_t)p.M, (uint64_t)p.lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B, &ta)) return MMI_ECUDA;
    if (!get_tensor_map(p.B, kext, (uint64_t)p.N, (uint64_t)p.ldb, BK, (uint32_t)bn, CU_TENSOR_MAP_SWIZZLE_128B, &tb)) return MMI_ECUDA;
  } else {
    const uint64_t kext = (uint64_t)p.K * (p.split_terms ? 3 : 1);
    if (!get_tensor_map(p.A, (uint64_t)p.M, kext, (uint64_t)p.lda, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B, &ta)) return MMI_ECUDA;
    if (!get_tensor_map(p.B, (uint64_t)p.N, kext, (uint64_t)p.ldb, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B, &tb)) return MMI_ECUDA;
  }
  const bool f32out = p.out_dtype == MMI_F32;
  MMI_CHECK_ARG(f32out || p.out_dtype == MMI_BF16, "gemm_tc: bad out dtype");
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  // TMA epilogue: bf16 output tiles leave through shared memory; at most one tile-shaped bf16 side operand
  bool tma_epi = !p.split_terms && !mn && !f32out && !p.accumulate && split == 1 && tma_ok(p.C, p.ldc) && !(p.add && p.mul_gelu_grad);
  const bool pe_add = p.add != nullptr && p.add_mod < p.M;   // position table: fp32 rows read directly, no TMA box
  if (tma_epi && p.add)
    tma_epi = pe_add ? (p.add_dtype == MMI_F32 && p.add_mod > 0 && p.ld_add % 4 == 0 && (reinterpret_cast<uintptr_t>(p.add) & 15) == 0 && !p.mul_gelu_grad)
                     : (p.add_dtype == MMI_BF16 && tma_ok(p.add, p.ld_add));
  if (tma_epi && p.mul_gelu_grad) tma_epi = tma_ok(p.mul_gelu_grad, p.ld_mul);
  if (tma_epi && p.preact) tma_epi = !p.add && !p.mul_gelu_grad && tma_ok(p.preact, p.ld_preact);
  if (tma_epi && pe_add && p.N % 64 != 0) tma_epi = false;  // the direct table reads cover whole 64-column chunks
  if (tma_epi) {
    const void* aux = pe_add ? nullptr : (p.add ? p.add : p.mul_gelu_grad);
    const int64_t ld_aux = p.add ? p.ld_add : p.ld_mul;
    if (!get_tensor_map(p.C, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldc, ECOLS, 32, CU_TENSOR_MAP_SWIZZLE_64B, &em.c)) return MMI_ECUDA;
    em.aux = em.c; em.c2 = em.c;
    if (aux && !get_tensor_map(aux, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)ld_aux, ECOLS, 32, CU_TENSOR_MAP_SWIZZLE_64B, &em.aux)) return MMI_ECUDA;
    if (p.preact && !get_tensor_map(p.preact, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ld_preact, ECOLS, 32, CU_TENSOR_MAP_SWIZZLE_64B, &em.c2)) return MMI_ECUDA;
    if (p.drop.thr8 != 0u) {
      if (bn == 256) return launch<256, false, __nv_bfloat16, true, true>(ta, tb, em, p, m_tiles, n_tiles, num_kb, split, st);
      return launch<128, false, __nv_bfloat16, true, true>(ta, tb, em, p, m_tiles, n_tiles, num_kb, split, st);
    }
    if (bn == 256) return launch<256, false, __nv_bfloat16, true>(ta, tb, em, p, m_tiles, n_tiles, num_kb, split, st);
    return launch<128, false, __nv_bfloat16, true>(ta, tb, em, p, m_tiles, n_tiles, num_kb, split, st);
  }
  if (p.split_terms)     // fp32 side operands, exact erf GELU, two accumulators per tile
    return mn ? launch<128, true, float, false, false, float>(ta, tb, em, p, m_tiles, n_tiles, num_kb, split, st)
              : launch<128, false, float, false, false, float>(ta, tb, em, p, m_tiles, n_tiles, num_kb, split, st);
#define MMI_TC_LAUNCH(BN_, MN_)                                                                                         \
  (f32out ? launch<BN_, MN_, float, false>(ta, tb, em, p, m_tiles, n_tiles, num_kb, split, st)                          \
          : launch<BN_, MN_, __nv_bfloat16, false>(ta, tb, em, p, m_tiles, n_tiles, num_kb, split, st))
  if (bn == 256) return mn ? MMI_TC_LAUNCH(256, true) : MMI_TC_LAUNCH(256, false);
  return mn ? MMI_TC_LAUNCH(128, true) : MMI_TC_LAUNCH(128, false);
#undef MMI_TC_LAUNCH
}

}  // namespace mmi
