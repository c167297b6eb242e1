// Included by attention_tc.cu inside namespace mmi::tc, after attention_bwd_fused.cuh (shares QVec64, dot8_bf16, desc_mn128).
//
// ====================================================================================== backward, one CTA per (b, h)
// The whole attention backward of one query side in ONE launch: the CTA owns every key of BOTH key blocks of its (b, h)
// (at most 5 tiles of 128 keys: 540 keys at the benchmark's shapes), so
//   * dK_j / dV_j of every key tile accumulate in TMEM over the whole query loop (5 x 64 columns),
//   * dQa / dQb of the current 64-query tile accumulate in TMEM over the key tiles of their block and are complete when the
//     key loop ends: they leave as bf16 rows straight from the accumulator -- no partial tiles, no atomics, no fp32
//     round trip through HBM (the per-key-block kernel of attention_bwd_fused.cuh needs all three and loses to them),
//   * Q, dO, lse and delta = rowsum(O * dO) of a query tile are fetched ONCE for all key tiles,
//   * S^T, dP^T and the exponential are computed once per score (5 MMAs + 1 ex2 instead of 7 + 2 in the dq / dk,dv pair).
// Per (query tile i, key tile j):  S^T = K_j Q_blk^T,  dP^T = V_j dO^T  (128 keys x 64 queries)  ->  softmax threads
// (thread = key row x 16-query chunk: 16 warps, per-query constants live in registers across the key tiles)  ->
// dV_j += P^T dO,  dK_j += dS^T Q_blk,  dQ_blk += dS K_j  (A = the dS^T staging tile read MN-major, M = 64).
// smem: K,V of all tiles 80 KB | Q ring [2][Qa | Qb | dO] 24 KB | P^T 16 KB | dS^T 16 KB | vectors, barriers   (~140 KB, 1 CTA / SM)
// TMEM: S^T @0 (64) | dP^T @64 (64) | dQa @128 | dQb @160 | dK_j @192+64j | dV_j @224+64j                        (512 columns)
// PERSISTENT: one CTA per SM walks the (b, h) items  blockIdx.x, + gridDim.x, ...  with every barrier phase running on
// global counters (t = key-tile steps, g = query tiles, n = items).  K/V tile j of the NEXT item is requested the moment
// the last product that reads tile j of the current item has retired (kv_empty[j]), so its load hides behind the remaining
// key tiles and the epilogue; barrier initialisation, the TMEM allocation and the launch tail are paid once per SM instead
// of once per (b, h) -- at the candidate side's shapes (one query tile per item) those were 3/4 of a CTA's life.
// score pairs (of the 8 a thread computes per tile) whose exponential runs as a degree-3 polynomial on the FMA pipe instead
// of MUFU.EX2: the 16 softmax warps hit the MUFU unit in lockstep (8 192 ex2 per tile = 512 cycles at 16 / clk / SM)
#ifndef AK_POLY
#define AK_POLY 0x00
#endif
__device__ __forceinline__ float2 ak_ex2(float2 x, int pair) { return ((AK_POLY >> pair) & 1) ? ex2_poly2(x) : ex2_mufu2(x); }
constexpr int AK_MAXT = 5;
constexpr int AK_THREADS = 64 + 16 * 32;
constexpr uint32_t AK_TMEM_COLS = 512;
constexpr uint32_t AK_QSTAGE = 3 * TILE64Q;      // Qa | Qb | dO
constexpr uint32_t AK_OUT = 32 * DH * 2;         // 2 KB: one warp's 32 key rows x 32 columns of dK or dV

struct AKBars {
  // p_ready[tile parity]: a warp whose lane quarter holds no key of tile t (nothing to compute) can finish tile t + 1 before
  // a slow warp has finished tile t; with ONE barrier its early arrival would complete tile t's phase (seen as garbage in
  // dK / dV of a partially filled key tile).  The issuer waits for p_ready of tile t before it can release S^T of tile
  // t + 2, so two alternating barriers are enough.
  uint64_t kv_full[AK_MAXT], kv_empty[AK_MAXT], q_full[2], q_empty[2], a_ready, s_free, p_ready[2], p_free, dq_ready, dq_free, done, acc_free;
  uint32_t tmem_slot, pad;
  uint32_t kmask[2][AK_MAXT * 4];   // [item parity][key tile][lane quarter]: bit l = key (tile, quarter, l) exists and is unmasked
};

template <bool DROP>
__global__ void __launch_bounds__(AK_THREADS, 1)
attn_bwd_allkeys_tc_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                           const __grid_constant__ CUtensorMap tmKa, const __grid_constant__ CUtensorMap tmKb,
                           const __grid_constant__ CUtensorMap tmVa, const __grid_constant__ CUtensorMap tmVb,
                           const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmdK0,
                           const __grid_constant__ CUtensorMap tmdK1, const __grid_constant__ CUtensorMap tmdV0,
                           const __grid_constant__ CUtensorMap tmdV1, const AttnTcParams p, int kv_store,
                           float* dbk0, float* dbk1, float* dbv0, float* dbv1, int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sKV = smem;                                    // [AK_MAXT][K 8 KB | V 8 KB]
  uint8_t* sQ = sKV + AK_MAXT * 2 * TILE128;              // [2][Qa | Qb | dO]
  uint8_t* sPT = sQ + 2 * AK_QSTAGE;                      // 104 KB from the base: 1024-aligned
  uint8_t* sdST = sPT + STILE;
  uint8_t* sOut = sdST + STILE;                           // [16 softmax warps][32 rows x 64 B, SWIZZLE_64B]: dK / dV slices on their way out
  QVec64* qv = reinterpret_cast<QVec64*>(sOut + 16 * AK_OUT);
  AKBars* bars = reinterpret_cast<AKBars*>(qv + 2);

  const int warp = (int)uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int nt0 = (p.Lk[0] + QT - 1) / QT, nt1 = p.nblk > 1 ? (p.Lk[1] + QT - 1) / QT : 0;
  const int NT = nt0 + nt1;                               // <= AK_MAXT (host)
  const int T = (p.Lq + QN - 1) / QN;

  // ---- producer state (warp 0): per-query scalars of the query tile about to be staged, two queries per lane
  float nl_n[2] = {0.f, 0.f}, nd_n[2] = {0.f, 0.f};
  uint32_t rh_n[2] = {0u, 0u};
  bool mq_n[2] = {false, false};
  auto fetch = [&](int b, int h, int i) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int qi = i * QN + u * 32 + lane;
      if (qi < p.Lq) {
        const int64_t tok = (int64_t)b * p.Lq + qi;
        const float lse = p.lse[((int64_t)b * p.H + h) * p.Lq + qi];
        mq_n[u] = p.mask_q[tok] != 0;
        const uint4* orow = reinterpret_cast<const uint4*>(p.out + tok * p.ldo + h * DH);
        const uint4* grow = reinterpret_cast<const uint4*>(p.dout + tok * p.lddo + h * DH);
        uint4 o[4], g[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) { o[d] = orow[d]; g[d] = grow[d]; }
        float delta = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) delta += dot8_bf16(o[d], g[d]);
        nl_n[u] = -lse * kLog2e;
        nd_n[u] = -delta * p.scale;
      } else { nl_n[u] = -INFINITY; nd_n[u] = 0.f; mq_n[u] = false; }
      rh_n[u] = DROP ? drop_rowhash(p.drop.key, (uint64_t)(((int64_t)b * p.H + h) * p.Lq + qi)) : 0u;
    }
  };
  // key-validity words of item (b, .), built by the producer warp one item ahead (coalesced byte loads + ballots): the softmax
  // threads used to fetch their own mask bytes at the start of an item -- five dependent global loads in front of its first tile
  auto fetch_mask = [&](int b, int slot) {
    uint8_t v[AK_MAXT * 4];                                // every byte load in flight before the first ballot needs one
#pragma unroll
    for (int j = 0; j < AK_MAXT; ++j) {
      const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j;
      const int Lk = blk ? p.Lk[1] : p.Lk[0];
      const uint8_t* mk = (blk ? p.mask_k[1] : p.mask_k[0]) + (int64_t)b * Lk;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int kj = kt * QT + q4 * 32 + lane;
        v[j * 4 + q4] = (j < NT && kj < Lk) ? __ldg(mk + kj) : (uint8_t)0;
      }
    }
#pragma unroll
    for (int x = 0; x < AK_MAXT * 4; ++x) {
      const uint32_t word = __ballot_sync(0xffffffffu, v[x] != 0);
      if (lane == 0) bars->kmask[slot][x] = word;
    }
  };
  if (warp == 0) {
    if (elect_one()) {
      for (int j = 0; j < AK_MAXT; ++j) { mbar_init(&bars->kv_full[j], 1); mbar_init(&bars->kv_empty[j], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&bars->q_full[s], 1); mbar_init(&bars->q_empty[s], 1); }
      mbar_init(&bars->a_ready, 1);
      mbar_init(&bars->s_free, 512);
      mbar_init(&bars->p_ready[0], 512);
      mbar_init(&bars->p_ready[1], 512);
      mbar_init(&bars->p_free, 1);
      mbar_init(&bars->dq_ready, 1);
      mbar_init(&bars->dq_free, 512);
      mbar_init(&bars->done, 1);
      mbar_init(&bars->acc_free, 512);
      fence_barrier_init();
    }
    __syncwarp();
    if ((int)blockIdx.x < n_items) { fetch(blockIdx.x / p.H, blockIdx.x % p.H, 0); fetch_mask(blockIdx.x / p.H, 0); }
  }
  if (warp == 2) TRACE(4090);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(AK_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  const uint32_t tdPT = tmem + QN, tdQ = tmem + 2 * QN, tdKV = tmem + 3 * QN;
  if (warp == 2) TRACE(4091);

  if (warp == 0) {
    // ================================================================= producer
    int g = 0;
    for (int w = blockIdx.x, n = 0; w < n_items; w += gridDim.x, ++n) {
    const int b = w / p.H, h = w % p.H;
    for (int i = 0; i < T; ++i, ++g) {
      const int st = g & 1;
      if (g >= 2) mbar_wait_bg(&bars->q_empty[st], ((g >> 1) & 1) ^ 1);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        qv[st].nlse2[u * 32 + lane] = nl_n[u];
        qv[st].nds[u * 32 + lane] = nd_n[u];
        if constexpr (DROP) qv[st].rh[u * 32 + lane] = rh_n[u];
        const uint32_t bits = __ballot_sync(0xffffffffu, mq_n[u]);
        if (lane == 0) qv[st].mq[u] = bits;
      }
      __syncwarp();
      if (elect_one()) {
        mbar_expect_tx(&bars->q_full[st], (p.nblk > 1 ? 3 : 2) * TILE64Q);
        uint8_t* dst = sQ + st * AK_QSTAGE;
        const int row = b * p.Lq + i * QN;
        tma_load_2d(&tmQa, &bars->q_full[st], dst, h * DH, row);
        if (p.nblk > 1) tma_load_2d(&tmQb, &bars->q_full[st], dst + TILE64Q, h * DH, row);
        tma_load_2d(&tmdO, &bars->q_full[st], dst + 2 * TILE64Q, h * DH, row);
      }
      __syncwarp();
      if (i == 0) {
        // K / V of this item, tile by tile as the previous item's last reader of each slot retires
        for (int j = 0; j < NT; ++j) {
          if (n >= 1) mbar_wait_bg(&bars->kv_empty[j], (n - 1) & 1);
          if (elect_one()) {
            const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j;
            const int krow = b * (blk ? p.Lk[1] : p.Lk[0]) + kt * QT;
            mbar_expect_tx(&bars->kv_full[j], 2 * TILE128);
            tma_load_2d(blk ? &tmKb : &tmKa, &bars->kv_full[j], sKV + j * 2 * TILE128, h * DH, krow);
            tma_load_2d(blk ? &tmVb : &tmVa, &bars->kv_full[j], sKV + j * 2 * TILE128 + TILE128, h * DH, krow);
          }
          __syncwarp();
        }
      }
      if (i + 1 < T) fetch(b, h, i + 1);
      else if (w + (int)gridDim.x < n_items) {
        fetch((w + gridDim.x) / p.H, (w + gridDim.x) % p.H, 0);
        fetch_mask((w + gridDim.x) / p.H, (n + 1) & 1);    // (its last readers were the softmax warps at the start of item n - 1)
      }
    }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    const uint32_t tS = uniform(tmem), tdPTu = uniform(tdPT), tdQu = uniform(tdQ), tdKVu = uniform(tdKV);
    const uint32_t aKV = smem_u32(sKV), aQs = smem_u32(sQ), aPT = smem_u32(sPT), adST = smem_u32(sdST);
    // (the issuer's operands must live in UNIFORM registers: anything data-dependent is routed through uniform(), and the
    // tile index is never divided -- a waterfall loop around every UTCHMMA costs ~80 cycles per MMA)
    auto issue_back = [&](int u, int iu, int ju, int gu, int nu) {
      const int su = gu & 1;
      const int blk = ju < nt0 ? 0 : 1, kt = blk ? ju - nt0 : ju;
      mbar_wait_bg(&bars->p_ready[u & 1], (u >> 1) & 1);
      if (ju == 0 && gu >= 1) mbar_wait_bg(&bars->dq_free, (gu - 1) & 1);      // last query tile's dQ has left the accumulators
      if (iu == 0 && ju == 0 && nu >= 1) mbar_wait_bg(&bars->acc_free, (nu - 1) & 1);   // last item's dK / dV have been read out
      tcgen05_fence_after();
      const uint32_t aQ = uniform(aQs + su * AK_QSTAGE + blk * TILE64Q), adO = uniform(aQs + su * AK_QSTAGE + 2 * TILE64Q);
      const uint32_t aK = uniform(aKV + ju * 2 * TILE128);
      const uint32_t tdK = uniform(tdKVu + ju * 2 * DH), tdV = tdK + DH, tdQb = uniform(tdQu + blk * DH);
      const uint32_t acc0 = uniform(iu > 0 ? 1u : 0u), accq = uniform(kt > 0 ? 1u : 0u);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdV, desc_k128(aPT, k), desc_mn64(adO, k), IDESC_O, k > 0 ? 1u : acc0);    // dV_j += P^T dO
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdK, desc_k128(adST, k), desc_mn64(aQ, k), IDESC_O, k > 0 ? 1u : acc0);    // dK_j += dS^T Q
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_f16(tdQb, desc_mn128(adST, k), desc_mn64(aK, k), IDESC_DQ, k > 0 ? 1u : accq);  // dQ_blk += dS K_j
        umma_commit(&bars->p_free);
        if (iu == T - 1) umma_commit(&bars->kv_empty[ju]);                      // the item's last reader of K_j / V_j
        if (ju == NT - 1) {
          umma_commit(&bars->q_empty[su]);
          umma_commit(&bars->dq_ready);
          if (iu == T - 1) umma_commit(&bars->done);
        }
      }
      __syncwarp();
    };
    int t = 0, g = 0, pi = 0, pj = 0, pg = 0, pn = 0;     // (pi, pj, pg, pn) = the (i, j, g, n) of step t - 1
    bool pending = false;                                 // the products of step t - 1 are still to be issued
    for (int w = blockIdx.x, n = 0; w < n_items; w += gridDim.x, ++n) {
    for (int i = 0; i < T; ++i, ++g) {
      const int st = g & 1;
      mbar_wait_bg(&bars->q_full[st], (g >> 1) & 1);
      for (int j = 0; j < NT; ++j, ++t) {
        const int blk = j < nt0 ? 0 : 1;
        // (a single key tile: the producer refills slot 0 only after the previous item's products have retired)
        if (NT == 1 && i == 0 && pending) { issue_back(t - 1, pi, pj, pg, pn); pending = false; }
        if (i == 0) mbar_wait_bg(&bars->kv_full[j], n & 1);
        if (t >= 1) mbar_wait_bg(&bars->s_free, (t - 1) & 1);
        tcgen05_fence_after();
        const uint32_t aQ = uniform(aQs + st * AK_QSTAGE + blk * TILE64Q), adO = uniform(aQs + st * AK_QSTAGE + 2 * TILE64Q);
        const uint32_t aK = uniform(aKV + j * 2 * TILE128), aV = aK + TILE128;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_f16(tS, desc_k64(aK, k), desc_k64(aQ, k), IDESC_S64T, k);        // S^T  = K_j Q^T
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_f16(tdPTu, desc_k64(aV, k), desc_k64(adO, k), IDESC_S64T, k);    // dP^T = V_j dO^T
          umma_commit(&bars->a_ready);
        }
        __syncwarp();
        if (t < 500) TRACE(t * 8 + 6);
        if (pending) issue_back(t - 1, pi, pj, pg, pn);
        if (t < 500) TRACE(t * 8 + 7);
        pi = i; pj = j; pg = g; pn = n; pending = true;
      }
    }
    }
    if (pending) issue_back(t - 1, pi, pj, pg, pn);
  } else {
    // ================================================================= softmax + dQ drain + epilogue (16 warps)
    const int qd = warp & 3, c16 = (warp - 2) >> 2, row = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t a_ready_a = smem_u32(&bars->a_ready), s_free_a = smem_u32(&bars->s_free), p_ready_a = smem_u32(&bars->p_ready[0]);
    const uint32_t p_free_a = smem_u32(&bars->p_free), q_full_a = smem_u32(&bars->q_full[0]), qv_a = smem_u32(qv);
    const uint32_t dq_ready_a = smem_u32(&bars->dq_ready), dq_free_a = smem_u32(&bars->dq_free);
    const uint32_t ptrow_a = smem_u32(sPT) + row * 128, dstrow_a = smem_u32(sdST) + row * 128, swz = row & 7;
    float dbq_run = 0.f;                                  // lane l < 16: running column sum of dQ column (c16 & 1) * 16 + l, block c16 >> 1
    // dK_j / dV_j leave by TMA: warp (qd, c16) takes rows qd*32 + lane of dK (c16 even) or dV (c16 odd) of the key tiles with
    // parity c16 >> 1, converts its 32 x 32 slice into a private SWIZZLE_64B staging tile and stores it with ONE
    // cp.async.bulk.tensor (3-D map: rows past Lk of this batch item are clipped).  Thread-per-row global stores put 32
    // different lines into every STG and made the LSU the bound of the epilogue (~6 000 cycles per item).
    // Bias-gradient column sums stay in registers over the tiles and items of one head (lanes l, l ^ 1 hold columns l >> 1
    // and 16 + (l >> 1); one pair per key block) and leave with one atomic per column when the head changes or the CTA ends.
    const int isv = c16 & 1;
    const uint32_t out_a = smem_u32(sOut) + (uint32_t)(warp - 2) * AK_OUT;
    const uint32_t outrow_a = out_a + lane * 64, oswz = (lane >> 1) & 3;
    float bs00 = 0.f, bs01 = 0.f, bs10 = 0.f, bs11 = 0.f;  // [key block][column half]
    int cur_h = -1;
    auto flush = [&]() {
      if (cur_h >= 0) {
        if ((c16 >> 1) < p.nblk && p.dbq[c16 >> 1] != nullptr && lane < 16)
          atomicAdd(p.dbq[c16 >> 1] + cur_h * DH + (c16 & 1) * 16 + lane, dbq_run);
        if ((lane & 1) == 0) {
          float* d0 = isv ? dbv0 : dbk0;
          float* d1 = isv ? dbv1 : dbk1;
          if (d0 != nullptr) { atomicAdd(d0 + cur_h * DH + (lane >> 1), bs00); atomicAdd(d0 + cur_h * DH + 16 + (lane >> 1), bs01); }
          if (nt1 > 0 && d1 != nullptr) { atomicAdd(d1 + cur_h * DH + (lane >> 1), bs10); atomicAdd(d1 + cur_h * DH + 16 + (lane >> 1), bs11); }
        }
      }
      dbq_run = 0.f; bs00 = 0.f; bs01 = 0.f; bs10 = 0.f; bs11 = 0.f;
    };
    auto drain_kv = [&](int b, int h, int j, uint32_t act_bits) {   // accumulators of key tile j are final (their last products retired)
      if ((c16 >> 1) != (j & 1)) return;                   // the other warp of this (lane quarter, accumulator) pair takes this tile
      if (!((act_bits >> j) & 1u)) return;                 // no key of the tile in this lane quarter (warp-uniform)
      const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j;
      const int Lk = blk ? p.Lk[1] : p.Lk[0];
      const bool k_in = kt * QT + row < Lk;
      const bool want_sum = (isv ? (blk ? dbv1 : dbv0) : (blk ? dbk1 : dbk0)) != nullptr;
      const bool want_store = (kv_store >> (blk * 2 + isv)) & 1;
      tcgen05_fence_after();
      if (lane == 0) bulk_wait_read0();                    // this warp's previous store has read the staging tile
      __syncwarp();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t rk[16];
        tmem_ld_32x32b_x16(tdKV + lane_addr + j * 2 * DH + isv * DH + half * 16, rk);
        tmem_ld_wait();
#pragma unroll
        for (uint32_t v = 0; v < 2; ++v)
          sts_u4(outrow_a + (((half * 2 + v) ^ oswz) << 4), pack_bf16x2(__uint_as_float(rk[8 * v]), __uint_as_float(rk[8 * v + 1])),
                 pack_bf16x2(__uint_as_float(rk[8 * v + 2]), __uint_as_float(rk[8 * v + 3])),
                 pack_bf16x2(__uint_as_float(rk[8 * v + 4]), __uint_as_float(rk[8 * v + 5])),
                 pack_bf16x2(__uint_as_float(rk[8 * v + 6]), __uint_as_float(rk[8 * v + 7])));
        if (want_sum) {
          float f[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) f[c] = k_in ? __uint_as_float(rk[c]) : 0.f;
#pragma unroll
          for (int o = 16, hv = 8; o >= 2; o >>= 1, hv >>= 1) {    // transposing butterfly: 16 columns x 32 rows -> column l >> 1
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int c = 0; c < hv; ++c) {
              const float send = up ? f[c] : f[c + hv];
              const float keep = up ? f[c + hv] : f[c];
              f[c] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          f[0] += __shfl_xor_sync(0xffffffffu, f[0], 1);
          if (blk) { if (half) bs11 += f[0]; else bs10 += f[0]; }
          else { if (half) bs01 += f[0]; else bs00 += f[0]; }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (want_store && elect_one()) {
        const CUtensorMap* m = isv ? (blk ? &tmdV1 : &tmdV0) : (blk ? &tmdK1 : &tmdK0);
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(m)), "r"(out_a), "r"(h * DH), "r"(kt * QT + qd * 32), "r"(b) : "memory");
        bulk_commit();
      }
      __syncwarp();
    };
    auto drain = [&](int b, int h, int i, int gq) {       // dQ of query tile i (global count gq): warp (qd, c16) owns rows qd*16 + lane,
      mbar_wait_a(dq_ready_a, gq & 1);                    // columns (c16 & 1) * 16 .. +16 of block c16 >> 1
      tcgen05_fence_after();
      const int bq = c16 >> 1, colh = (c16 & 1) * 16;
      uint32_t r[16];
      tmem_ld_32x32b_x16(tdQ + lane_addr + bq * DH + colh, r);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive_a(dq_free_a);
      if (bq < p.nblk) {
        const int q = i * QN + qd * 16 + lane;            // M = 64 accumulator: row m lives in lane (m & 15) of lane quarter m >> 4
        const bool ok = lane < 16 && q < p.Lq;
        if (ok && p.dq[bq] != nullptr) {
          uint4* dst = reinterpret_cast<uint4*>(p.dq[bq] + ((int64_t)b * p.Lq + q) * p.lddq[bq] + h * DH + colh);
          uint32_t w[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) w[c] = pack_bf16x2(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1]));
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        if (p.dbq[bq] != nullptr) {
          // column sums over the tile's rows -> bias gradient of the query projection: a transposing butterfly over the 16
          // row-holding lanes (15 shuffles) leaves column l's sum on lane l, which keeps ONE running register across the
          // query tiles (shared-memory float atomics from 16 warps in lockstep cost ~6 000 cycles per query tile here)
          float v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = ok ? __uint_as_float(r[c]) : 0.f;
#pragma unroll
          for (int o = 8; o >= 1; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int c = 0; c < o; ++c) {
              const float send = up ? v[c] : v[c + o];
              const float keep = up ? v[c + o] : v[c];
              v[c] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          dbq_run += v[0];
        }
      }
    };
    // the END of an item -- dQ of its last query tile, dK / dV of its last key tile -- is not waited for: it is picked up
    // behind the first key tile of the NEXT item, while the tensor pipe already works on that tile's products (which wait for
    // dq_free / acc_free, arrived here).  At one query tile per item this tail was 1/5 of the item.
    int pend_b = -1, pend_h = 0, pend_n = 0;
    uint32_t pend_act = 0u;
    auto finish_item = [&](int gq) {                      // gq = global count of the item's last query tile
      drain(pend_b, pend_h, T - 1, gq);
      mbar_wait(&bars->done, pend_n & 1);
      if (warp == 2 && pend_n < 20) TRACE(4000 + pend_n * 4 + 2);
      drain_kv(pend_b, pend_h, NT - 1, pend_act);
      tcgen05_fence_before();
      mbar_arrive(&bars->acc_free);                       // (every tcgen05.ld of this thread has completed)
      if (warp == 2 && pend_n < 20) TRACE(4000 + pend_n * 4 + 3);
      pend_b = -1;
    };
    int t = 0, g = 0;
    for (int w = blockIdx.x, n = 0; w < n_items; w += gridDim.x, ++n) {
    const int b = w / p.H, h = w % p.H;
    mbar_wait_a(q_full_a + (g & 1) * 8, (g >> 1) & 1);    // the item's first query tile: its arrival also publishes kmask[n & 1]
    uint32_t mk_bits = 0u;                                 // bit j: this thread's key of tile j exists and is unmasked
    uint32_t act_bits = 0u, allmk_bits = 0u;               // warp-uniform: tile j has a real key in this lane quarter / no masked key
    for (int j = 0; j < NT; ++j) {
      const int blk = j < nt0 ? 0 : 1, kt = blk ? j - nt0 : j;
      const uint32_t word = bars->kmask[n & 1][j * 4 + qd];
      mk_bits |= ((word >> lane) & 1u) << j;
      if (kt * QT + qd * 32 < (blk ? p.Lk[1] : p.Lk[0])) act_bits |= 1u << j;
      if (word == 0xffffffffu) allmk_bits |= 1u << j;
    }
    if (n == 0) cur_h = h;
    if (warp == 2 && n < 20) TRACE(4000 + n * 4);
    for (int i = 0; i < T; ++i, ++g) {
      const int st = g & 1;
      mbar_wait_a(q_full_a + st * 8, (g >> 1) & 1);       // acquire the producer's per-query vectors
      const uint32_t qva = qv_a + st * (uint32_t)sizeof(QVec64) + c16 * 16 * 4;    // this warp's 16 queries
      const uint32_t wq = (lds_u1(qv_a + st * (uint32_t)sizeof(QVec64) + QV_MQ + (c16 >> 1) * 4) >> ((c16 & 1) * 16)) & 0xffffu;
      // per-query constants -log2e*lse and -scale*delta: re-read from shared memory (broadcast LDS.64) next to their use -- kept
      // in registers across the key tiles they cost 32 registers and pushed the loop into local-memory spills at 96 regs/thread
      auto NL2 = [&](int c) { return lds_f2(qva + c * 4); };
      auto ND2 = [&](int c) { return lds_f2(qva + QV_NDS + c * 4); };
      uint32_t rh = 0u;
      if constexpr (DROP) rh = lds_u1(qva + QV_RH + (lane & 15) * 4);
      uint32_t kq2 = 0u;                                   // DROP: keep bits of this thread's key for the 16 queries, tiles j (low half) and j + 1 (high half)
      for (int j = 0; j < NT; ++j, ++t) {
        const bool act = (act_bits >> j) & 1u;
        const bool mk = (mk_bits >> j) & 1u;
        const bool fast = !DROP && ((allmk_bits >> j) & 1u) && wq == 0xffffu;
        uint32_t kq = 0u;
        if constexpr (DROP) {
          if ((j & 1) == 0) {
            // lanes 0-15 generate the keep words of (query lane, this warp's 32 keys of tile j), lanes 16-31 those of tile
            // j + 1; one 32 x 32 bit transpose hands every thread (= key) its bits for both tiles
            const int jj = min(j + (lane >> 4), NT - 1);
            const int blk2 = jj < nt0 ? 0 : 1, kt2 = blk2 ? jj - nt0 : jj;
            const uint32_t Wq = drop_keep_word(rh, attn_group(blk2, kt2 * 4 + qd), p.drop.thr8);
            kq2 = warp_bit_transpose32(Wq, lane);
          }
          kq = (j & 1) ? (kq2 >> 16) : (kq2 & 0xffffu);
        }
        if (warp == 2 && t < 500) TRACE(t * 8 + 0);
        mbar_wait_a(a_ready_a, t & 1);
        tcgen05_fence_after();
        if (warp == 2 && t < 500) TRACE(t * 8 + 1);
        uint32_t rs[16], rp[16];
        if (act) {
          tmem_ld_32x32b_x16(tmem + lane_addr + c16 * 16, rs);
          tmem_ld_32x32b_x16(tdPT + lane_addr + c16 * 16, rp);
          tmem_ld_wait();
        }
        tcgen05_fence_before();
        mbar_arrive_a(s_free_a);
        if (warp == 2 && t < 500) TRACE(t * 8 + 2);
        uint32_t pp[8], pd[8];
        if (!act) {
#pragma unroll
          for (int c = 0; c < 8; ++c) { pp[c] = 0u; pd[c] = 0u; }
        } else if constexpr (DROP) {
          const float ds_ = p.drop.scale;
          const float2 sl2 = splat2(p.scale_log2 * ds_), sc2 = splat2(p.scale * ds_), dsc2 = splat2(ds_);
          if (((allmk_bits >> j) & 1u) && wq == 0xffffu) {
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              const bool k0_ = (kq >> c) & 1u, k1_ = (kq >> (c + 1)) & 1u;
              const float s0 = k0_ ? __uint_as_float(rs[c]) : 0.f, s1 = k1_ ? __uint_as_float(rs[c + 1]) : 0.f;
              const float2 pr = ak_ex2(fma2(make_float2(s0, s1), sl2, NL2(c)), c >> 1);
              const float2 ds = mul2(pr, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, mul2(ND2(c), dsc2)));
              pp[c >> 1] = pack_bf16x2(pr.x, pr.y);
              pd[c >> 1] = pack_bf16x2(k0_ ? ds.x : 0.f, k1_ ? ds.y : 0.f);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              const bool k0_ = (kq >> c) & 1u, k1_ = (kq >> (c + 1)) & 1u;
              const bool v0 = mk && ((wq >> c) & 1u), v1 = mk && ((wq >> (c + 1)) & 1u);
              const float s0 = k0_ ? (v0 ? __uint_as_float(rs[c]) : -10000.0f) : 0.f, s1 = k1_ ? (v1 ? __uint_as_float(rs[c + 1]) : -10000.0f) : 0.f;
              const float2 pr = ex2_mufu2(fma2(make_float2(s0, s1), sl2, NL2(c)));
              const float2 ds = mul2(pr, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, mul2(ND2(c), dsc2)));
              pp[c >> 1] = pack_bf16x2(pr.x, pr.y);
              pd[c >> 1] = pack_bf16x2((k0_ && v0) ? ds.x : 0.f, (k1_ && v1) ? ds.y : 0.f);
            }
          }
        } else if (fast) {
          const float2 sl2 = splat2(p.scale_log2), sc2 = splat2(p.scale);
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float2 x = fma2(make_float2(__uint_as_float(rs[c]), __uint_as_float(rs[c + 1])), sl2, NL2(c));
            const float2 pr = ak_ex2(x, c >> 1);
            const float2 ds = mul2(pr, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, ND2(c)));
            pp[c >> 1] = pack_bf16x2(pr.x, pr.y);
            pd[c >> 1] = pack_bf16x2(ds.x, ds.y);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float pr[2], ds[2];
            const float2 nl2 = NL2(c), nd2 = ND2(c);
            const float nl[2] = {nl2.x, nl2.y}, nd[2] = {nd2.x, nd2.y};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int cc = c + u;
              const bool valid = mk && (((wq >> cc) & 1u) != 0);
              const float x = valid ? __uint_as_float(rs[cc]) * p.scale_log2 : p.fill_log2;
              pr[u] = ex2(x + nl[u]);                    // queries past Lq: nlse2 = -inf => 0
              ds[u] = valid ? pr[u] * fmaf(__uint_as_float(rp[cc]), p.scale, nd[u]) : 0.f;
            }
            pp[c >> 1] = pack_bf16x2(pr[0], pr[1]);
            pd[c >> 1] = pack_bf16x2(ds[0], ds[1]);
          }
        }
        if (warp == 2 && t < 500) TRACE(t * 8 + 3);
        if (t >= 1) mbar_wait_a(p_free_a, (t - 1) & 1);    // the products of the previous tile have consumed the staging tiles
        if (warp == 2 && t < 500) TRACE(t * 8 + 4);
#pragma unroll
        for (uint32_t v = 0; v < 2; ++v) {
          const uint32_t off = ((c16 * 2 + v) ^ swz) << 4;
          sts_u4(ptrow_a + off, pp[4 * v], pp[4 * v + 1], pp[4 * v + 2], pp[4 * v + 3]);
          sts_u4(dstrow_a + off, pd[4 * v], pd[4 * v + 1], pd[4 * v + 2], pd[4 * v + 3]);
        }
        fence_proxy_async_smem();
        mbar_arrive_a(p_ready_a + (t & 1) * 8);
        if (warp == 2 && t < 500) TRACE(t * 8 + 5);
        if (j == 0 && i >= 1) drain(b, h, i - 1, g - 1);   // the previous query tile's dQ: complete long ago, never waited for
        // last query tile: dK / dV of the previous key tile are final (p_free of its products was waited for above) and
        // leave while the tensor pipe works on this tile's products
        if (i == 0 && j == 0 && pend_b >= 0) {             // the previous item's tail, then (maybe) a new head
          finish_item(g - 1);
          if (h != cur_h) { flush(); cur_h = h; }
        }
        if (i == T - 1 && j >= 1) drain_kv(b, h, j - 1, act_bits);
      }
    }
    if (warp == 2 && n < 20) TRACE(4000 + n * 4 + 1);
    pend_b = b; pend_h = h; pend_n = n; pend_act = act_bits;
    }
    if (pend_b >= 0) finish_item(g - 1);
    flush();
    if (lane == 0) bulk_wait0();                          // the staging tiles stay valid until the last stores have read them
    __syncwarp();
  }
  if (warp == 2) TRACE(4094);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) TRACE(4092);
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AK_TMEM_COLS) : "memory");
  }
}
