// Included by attention_tc.cu inside namespace mmi::tc (it shares that file's helpers).
//
// ====================================================================================== backward, fused: dK, dV AND dQ
// One kernel per key block instead of dq + dk/dv: S^T, dP^T and the exponential are computed ONCE per score (5 MMAs and
// 1 ex2 per score instead of 7 and 2), and the query-side gradient leaves as partial tiles.
//   CTA = 128 keys of block `which` for one (b, h); loops over 64-query tiles; thread = (key row, 32-query half):
//   warp 0 producer (per-query lse / delta / mask / dropout row hash + TMA of Q, dO), warp 1 MMA + TMEM, warps 2-9 softmax
//   (TMEM lane quarter = warp & 3, query half = (warp - 2) >> 2).  Two CTAs per SM: 16 softmax warps, like the two-kernel path.
//     S^T  = K Q^T,  dP^T = V dO^T                       (M 128 keys x N 64 queries, K-major operands)
//     dV  += P^T dO, dK += dS^T Q                        (A = staged P^T / dS^T, K-major SW128; B MN-major)
//     dQ_i = dS_i K                                      (M 64 queries x N 32: A = the SAME dS^T staging tile read MN-major,
//                                                         B = the CTA's K tile read MN-major; fresh accumulator per tile)
//   dQ_i is a partial sum over this CTA's 128 keys: the softmax warps pull it out of TMEM (an M = 64 accumulator sits in
//   lanes 0-15 of every lane quarter) and add it to an fp32 accumulator in global memory with red.global.add.v4.f32
//   (measured: 6.4 TB/s of partial tiles while the target lines are L2-resident, tools/micro/red_bw.cu).  The LAST key-tile
//   CTA of a (b, h) -- an atomic counter decides -- converts that (b, h)'s accumulator columns to bf16, adds their column
//   sums to the query-projection bias gradient and clears what it read, so the accumulator never needs a separate pass.
//   delta = rowsum(O * dO) is recomputed by the producer warp from the rows it is about to stage (no delta tensor).
// smem: K 8 KB | V 8 KB | Q,dO ring [3][4 KB + 4 KB] | P^T 16 KB | dS^T 16 KB | per-query vectors | barriers   (~75 KB)
// TMEM: S^T @0 (64) | dP^T @64 (64) | dK @128 | dV @160 | dQ @192                                        (256 columns)
constexpr int QN = 64;
constexpr int FB_THREADS = 320;
constexpr int FB_STAGES = 3;
constexpr uint32_t FB_TMEM_COLS = 256;
constexpr uint32_t TILE64Q = QN * DH * 2;        // 4 KB : 64 rows x 64 B (SWIZZLE_64B)
constexpr uint32_t STILE = QT * QN * 2;          // 16 KB: 128 rows x 128 B (SWIZZLE_128B)
constexpr uint32_t IDESC_S64T = make_idesc(QT, QN, false, false);
constexpr uint32_t IDESC_DQ = make_idesc(QN, DH, true, true);
// MN-major SWIZZLE_128B operand: rows = K dimension (128 B each), 8-row groups 1024 B apart, 16 rows per MMA
__device__ __forceinline__ uint64_t desc_mn128(uint32_t addr, int kstep) { return make_smem_desc(addr + kstep * 2048, 8192, 1024, 2); }

struct QVec64 {
  float nlse2[QN];         // -lse * log2(e); -inf for queries past Lq (P = 0)
  float nds[QN];           // -delta * scale
  uint32_t rh[QN];         // DROP: dropout row hash of the query
  uint32_t mq[2];          // valid-query bits of the two halves
  uint32_t pad[2];
};
constexpr uint32_t QV_NDS = QN * 4, QV_RH = 2 * QN * 4, QV_MQ = 3 * QN * 4;
struct FBars {
  uint64_t once, kv_full[FB_STAGES], kv_empty[FB_STAGES], a_ready, s_free, p_ready, p_free, dq_ready, dq_free, done;
  uint32_t tmem_slot, is_last;
  float colsum[DH];
};
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& g) {
  const uint32_t x[4] = {a.x, a.y, a.z, a.w}, y[4] = {g.x, g.y, g.z, g.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    s += __uint_as_float(x[i] << 16) * __uint_as_float(y[i] << 16) + __uint_as_float(x[i] & 0xffff0000u) * __uint_as_float(y[i] & 0xffff0000u);
  return s;
}

template <bool DROP>
__global__ void __launch_bounds__(FB_THREADS, 2)
attn_bwd_fused_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const AttnTcParams p,
                         float* __restrict__ dq_acc, int* __restrict__ dq_count) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE128;
  uint8_t* sQdO = sV + TILE128;                           // [FB_STAGES][Q 4 KB | dO 4 KB]
  uint8_t* sPT = sQdO + FB_STAGES * 2 * TILE64Q;          // 40 KB from the base: 1024-aligned
  uint8_t* sdST = sPT + STILE;
  QVec64* qv = reinterpret_cast<QVec64*>(sdST + STILE);
  FBars* bars = reinterpret_cast<FBars*>(qv + FB_STAGES);

  const int warp = (int)uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * QT;
  const int blk = p.which;
  const int Lk = (blk ? p.Lk[1] : p.Lk[0]);
  const int T = (p.Lq + QN - 1) / QN;
  const int rows_valid = min(QT, Lk - k0);
  const int nact = (rows_valid + 31) >> 5;                // lane quarters with at least one real key

  // ---- producer state: the per-query scalars of the tile about to be staged, two queries per lane (lane, lane + 32)
  float nl_n[2] = {0.f, 0.f}, nd_n[2] = {0.f, 0.f};
  uint32_t rh_n[2] = {0u, 0u};
  bool mq_n[2] = {false, false};
  auto fetch = [&](int i) {                               // whole warp 0
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
#ifdef MMI_FB_NOFETCH
#pragma unroll
        for (int d = 0; d < 4; ++d) { o[d] = make_uint4(0, 0, 0, 0); g[d] = o[d]; }
        (void)orow; (void)grow;
#else
#pragma unroll
        for (int d = 0; d < 4; ++d) { o[d] = orow[d]; g[d] = grow[d]; }
#endif
        float delta = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) delta += dot8_bf16(o[d], g[d]);
        nl_n[u] = -lse * kLog2e;
        nd_n[u] = -delta * p.scale;
      } else { nl_n[u] = -INFINITY; nd_n[u] = 0.f; mq_n[u] = false; }
      rh_n[u] = DROP ? drop_rowhash(p.drop.key, (uint64_t)(((int64_t)b * p.H + h) * p.Lq + qi)) : 0u;
    }
  };
  if (warp == 0) {
    if (elect_one()) {
      mbar_init(&bars->once, 1);
      for (int s = 0; s < FB_STAGES; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
      mbar_init(&bars->a_ready, 1);
      mbar_init(&bars->s_free, 64 * nact);
      mbar_init(&bars->p_ready, 64 * nact);
      mbar_init(&bars->p_free, 1);
      mbar_init(&bars->dq_ready, 1);
      mbar_init(&bars->dq_free, 256);
      mbar_init(&bars->done, 1);
      fence_barrier_init();
      mbar_expect_tx(&bars->once, 2 * TILE128);
      tma_load_2d(&tmK, &bars->once, sK, h * DH, b * Lk + k0);
      tma_load_2d(&tmV, &bars->once, sV, h * DH, b * Lk + k0);
    }
    __syncwarp();
    fetch(0);
  }
  if (threadIdx.x < DH) bars->colsum[threadIdx.x] = 0.f;
  // dS^T rows of lane quarters without a single real key are never written by a softmax thread but ARE summed over by the
  // dQ product (its K dimension is the CTA's 128 keys): zero them once
  if (warp >= 2 && (warp & 3) >= nact) {
    const uint32_t row = (warp & 3) * 32 + lane, hf = (warp - 2) >> 2;
#pragma unroll
    for (uint32_t v = 0; v < 4; ++v) sts_u4(smem_u32(sdST) + row * 128 + (((hf * 4 + v) ^ (row & 7)) << 4), 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  const int kj_pre = k0 + (warp & 3) * 32 + lane;
  const uint8_t mk_pre = (warp >= 2 && kj_pre < Lk) ? (blk ? p.mask_k[1] : p.mask_k[0])[(int64_t)b * Lk + kj_pre] : (uint8_t)0;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(FB_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  const uint32_t tdPT = tmem + QN, tdK = tmem + 2 * QN, tdV = tdK + DH, tdQ = tdV + DH;

  if (warp == 0) {
    // ================================================================= producer
    for (int i = 0; i < T; ++i) {
      const int st = i % FB_STAGES;
      if (i >= FB_STAGES) mbar_wait_bg(&bars->kv_empty[st], ((i / FB_STAGES) & 1) ^ 1);
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
        mbar_expect_tx(&bars->kv_full[st], 2 * TILE64Q);
        uint8_t* dst = sQdO + st * 2 * TILE64Q;
        const int row = b * p.Lq + i * QN;
        tma_load_2d(&tmQ, &bars->kv_full[st], dst, h * DH, row);
        tma_load_2d(&tmdO, &bars->kv_full[st], dst + TILE64Q, h * DH, row);
      }
      __syncwarp();
      if (i + 1 < T) fetch(i + 1);                        // the next tile's scalars: their latency hides behind the consumers
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    const uint32_t tS = uniform(tmem), tdPTu = uniform(tdPT), tdKu = uniform(tdK), tdVu = uniform(tdV), tdQu = uniform(tdQ);
    mbar_wait(&bars->once, 0);
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aPT = smem_u32(sPT), adST = smem_u32(sdST);
    auto issue_back = [&](int u) {
      const int st = u % FB_STAGES;
      mbar_wait_bg(&bars->p_ready, u & 1);
#ifndef MMI_FB_NODRAIN
      if (u >= 1) mbar_wait_bg(&bars->dq_free, (u - 1) & 1);
#endif
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t aQ = smem_u32(sQdO + st * 2 * TILE64Q), adO = aQ + TILE64Q;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdVu, desc_k128(aPT, k), desc_mn64(adO, k), IDESC_O, (u > 0 || k > 0) ? 1u : 0u);    // dV += P^T dO
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdKu, desc_k128(adST, k), desc_mn64(aQ, k), IDESC_O, (u > 0 || k > 0) ? 1u : 0u);    // dK += dS^T Q
#pragma unroll
#ifndef MMI_FB_NODQMMA
        for (int k = 0; k < 8; ++k) umma_f16(tdQu, desc_mn128(adST, k), desc_mn64(aK, k), IDESC_DQ, k > 0 ? 1u : 0u);             // dQ_u = dS K
#endif
        umma_commit(&bars->p_free);
        umma_commit(&bars->kv_empty[st]);
        umma_commit(&bars->dq_ready);
      }
      __syncwarp();
    };
    for (int i = 0; i < T; ++i) {
      const int st = i % FB_STAGES;
      mbar_wait_bg(&bars->kv_full[st], (i / FB_STAGES) & 1);
      if (i >= 1) mbar_wait_bg(&bars->s_free, (i - 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t aQ = smem_u32(sQdO + st * 2 * TILE64Q), adO = aQ + TILE64Q;
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tS, desc_k64(aK, k), desc_k64(aQ, k), IDESC_S64T, k);        // S^T  = K Q^T
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(tdPTu, desc_k64(aV, k), desc_k64(adO, k), IDESC_S64T, k);    // dP^T = V dO^T
        umma_commit(&bars->a_ready);
      }
      __syncwarp();
      if (i >= 1) issue_back(i - 1);
    }
    issue_back(T - 1);
    if (elect_one()) umma_commit(&bars->done);
    __syncwarp();
  } else {
    // ================================================================= softmax + dQ drain (8 warps)
    const int qd = warp & 3, hf = (warp - 2) >> 2, row = qd * 32 + lane;
    const bool active = qd < nact;
    const int kj = k0 + row;
    const bool k_in = kj < Lk;
    const bool mk = k_in ? (mk_pre != 0) : false;
    const bool warp_all_mk = __all_sync(0xffffffffu, mk);
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t a_ready_a = smem_u32(&bars->a_ready), s_free_a = smem_u32(&bars->s_free), p_ready_a = smem_u32(&bars->p_ready);
    const uint32_t p_free_a = smem_u32(&bars->p_free), kv_full_a = smem_u32(&bars->kv_full[0]), qv_a = smem_u32(qv);
    const uint32_t dq_ready_a = smem_u32(&bars->dq_ready), dq_free_a = smem_u32(&bars->dq_free);
    const uint32_t ptrow_a = smem_u32(sPT) + row * 128, dstrow_a = smem_u32(sdST) + row * 128, swz = row & 7;
    float* acc_base = dq_acc + ((int64_t)b * p.Lq * p.H + h) * DH + hf * 16;      // + q * H * DH
    auto drain = [&](int u) {                             // every softmax warp: 16 rows x 16 columns of dQ_u
#ifdef MMI_FB_NODRAIN
      return;
#endif
      mbar_wait_a(dq_ready_a, u & 1);
      tcgen05_fence_after();
      uint32_t r[16];
      tmem_ld_32x32b_x16(tdQ + lane_addr + hf * 16, r);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive_a(dq_free_a);
      const int q = u * QN + qd * 16 + lane;              // M = 64 accumulator: row m lives in lane (m & 15) of lane quarter m >> 4
#ifndef MMI_FB_NORED
      if (lane < 16 && q < p.Lq) {
        float* dst = acc_base + (int64_t)q * p.H * DH;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          red_add_v4(dst + 4 * j, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      }
#endif
    };
    int st = 0, st_phase = 0;
    for (int i = 0; i < T; ++i) {
      if (active) {
        mbar_wait_a(kv_full_a + st * 8, st_phase);        // acquire the producer's per-query vectors
        const uint32_t qva = qv_a + st * (uint32_t)sizeof(QVec64) + hf * 32 * 4;   // this half's 32 queries
        const uint32_t wq = lds_u1(qv_a + st * (uint32_t)sizeof(QVec64) + QV_MQ + hf * 4);
        const bool fast = !DROP && warp_all_mk && wq == 0xffffffffu;
        uint32_t kq = 0u;                                  // DROP: bit c = this thread's key is kept for query c of the half
        if constexpr (DROP) {
          const uint32_t Wq = drop_keep_word(lds_u1(qva + QV_RH + lane * 4), attn_group(blk, (k0 >> 5) + qd), p.drop.thr8);
          kq = warp_bit_transpose32(Wq, lane);
        }
        mbar_wait_a(a_ready_a, i & 1);
        tcgen05_fence_after();
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int ck = 0; ck < 2; ++ck) {
          uint32_t rs[16], rp[16];
          tmem_ld_32x32b_x16(tmem + lane_addr + hf * 32 + ck * 16, rs);
          tmem_ld_32x32b_x16(tdPT + lane_addr + hf * 32 + ck * 16, rp);
          tmem_ld_wait();
          if (ck == 1) {
            tcgen05_fence_before();
            mbar_arrive_a(s_free_a);
          }
          if constexpr (DROP) {
            const float ds_ = p.drop.scale;
            const float2 sl2 = splat2(p.scale_log2 * ds_), sc2 = splat2(p.scale * ds_), dsc2 = splat2(ds_);
            const uint32_t kqh = kq >> (ck * 16), wqh = wq >> (ck * 16);
            if (warp_all_mk && wq == 0xffffffffu) {
#pragma unroll
              for (int c = 0; c < 16; c += 2) {
                const float2 nl = lds_f2(qva + (ck * 16 + c) * 4), nd = lds_f2(qva + QV_NDS + (ck * 16 + c) * 4);
                const bool k0_ = (kqh >> c) & 1u, k1_ = (kqh >> (c + 1)) & 1u;
                const float s0 = k0_ ? __uint_as_float(rs[c]) : 0.f, s1 = k1_ ? __uint_as_float(rs[c + 1]) : 0.f;
                const float2 pr = ex2_mufu2(fma2(make_float2(s0, s1), sl2, nl));
                const float2 ds = mul2(pr, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, mul2(nd, dsc2)));
                pp[ck * 8 + (c >> 1)] = pack_bf16x2(pr.x, pr.y);
                pd[ck * 8 + (c >> 1)] = pack_bf16x2(k0_ ? ds.x : 0.f, k1_ ? ds.y : 0.f);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 16; c += 2) {
                const float2 nl = lds_f2(qva + (ck * 16 + c) * 4), nd = lds_f2(qva + QV_NDS + (ck * 16 + c) * 4);
                const bool k0_ = (kqh >> c) & 1u, k1_ = (kqh >> (c + 1)) & 1u;
                const bool v0 = mk && ((wqh >> c) & 1u), v1 = mk && ((wqh >> (c + 1)) & 1u);
                const float s0 = k0_ ? (v0 ? __uint_as_float(rs[c]) : -10000.0f) : 0.f, s1 = k1_ ? (v1 ? __uint_as_float(rs[c + 1]) : -10000.0f) : 0.f;
                const float2 pr = ex2_mufu2(fma2(make_float2(s0, s1), sl2, nl));
                const float2 ds = mul2(pr, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, mul2(nd, dsc2)));
                pp[ck * 8 + (c >> 1)] = pack_bf16x2(pr.x, pr.y);
                pd[ck * 8 + (c >> 1)] = pack_bf16x2((k0_ && v0) ? ds.x : 0.f, (k1_ && v1) ? ds.y : 0.f);
              }
            }
          } else if (fast) {
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              const float4 nl = lds_f4(qva + (ck * 16 + c) * 4), nd = lds_f4(qva + QV_NDS + (ck * 16 + c) * 4);
              const float2 sl2 = splat2(p.scale_log2), sc2 = splat2(p.scale);
              const float2 xa = fma2(make_float2(__uint_as_float(rs[c]), __uint_as_float(rs[c + 1])), sl2, make_float2(nl.x, nl.y));
              const float2 xb = fma2(make_float2(__uint_as_float(rs[c + 2]), __uint_as_float(rs[c + 3])), sl2, make_float2(nl.z, nl.w));
              const float2 pa = ex2_mufu2(xa), pb = ex2_mufu2(xb);
              const float2 da = mul2(pa, fma2(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), sc2, make_float2(nd.x, nd.y)));
              const float2 db = mul2(pb, fma2(make_float2(__uint_as_float(rp[c + 2]), __uint_as_float(rp[c + 3])), sc2, make_float2(nd.z, nd.w)));
              pp[ck * 8 + (c >> 1)] = pack_bf16x2(pa.x, pa.y);
              pp[ck * 8 + (c >> 1) + 1] = pack_bf16x2(pb.x, pb.y);
              pd[ck * 8 + (c >> 1)] = pack_bf16x2(da.x, da.y);
              pd[ck * 8 + (c >> 1) + 1] = pack_bf16x2(db.x, db.y);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              const float4 nl = lds_f4(qva + (ck * 16 + c) * 4), nd = lds_f4(qva + QV_NDS + (ck * 16 + c) * 4);
              const float nlv[4] = {nl.x, nl.y, nl.z, nl.w}, ndv[4] = {nd.x, nd.y, nd.z, nd.w};
              float pr[4], ds[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int cc = c + t;
                const bool valid = mk && (((wq >> (ck * 16 + cc)) & 1u) != 0);
                const float x = valid ? __uint_as_float(rs[cc]) * p.scale_log2 : p.fill_log2;
                pr[t] = ex2(x + nlv[t]);                   // queries past Lq: nlse2 = -inf => 0
                ds[t] = valid ? pr[t] * fmaf(__uint_as_float(rp[cc]), p.scale, ndv[t]) : 0.f;
              }
              pp[ck * 8 + (c >> 1)] = pack_bf16x2(pr[0], pr[1]);
              pp[ck * 8 + (c >> 1) + 1] = pack_bf16x2(pr[2], pr[3]);
              pd[ck * 8 + (c >> 1)] = pack_bf16x2(ds[0], ds[1]);
              pd[ck * 8 + (c >> 1) + 1] = pack_bf16x2(ds[2], ds[3]);
            }
          }
        }
        if (i >= 1) mbar_wait_a(p_free_a, (i - 1) & 1);    // dK / dV / dQ products of tile i-1 have consumed the staging tiles
#pragma unroll
        for (uint32_t v = 0; v < 4; ++v) {
          const uint32_t off = ((hf * 4 + v) ^ swz) << 4;
          sts_u4(ptrow_a + off, pp[4 * v], pp[4 * v + 1], pp[4 * v + 2], pp[4 * v + 3]);
          sts_u4(dstrow_a + off, pd[4 * v], pd[4 * v + 1], pd[4 * v + 2], pd[4 * v + 3]);
        }
        fence_proxy_async_smem();
        mbar_arrive_a(p_ready_a);
      }
      if (++st == FB_STAGES) { st = 0; st_phase ^= 1; }
      if (i >= 1) drain(i - 1);
    }
    drain(T - 1);
    if (active) {
      // epilogue: the query-half-0 warp of a lane quarter writes dK, its half-1 twin dV
      mbar_wait(&bars->done, 0);
      tcgen05_fence_after();
      uint32_t rk[32];
      tmem_ld_32x32((hf ? tdV : tdK) + lane_addr, rk);
      tmem_ld_wait();
      __nv_bfloat16* dst = hf ? p.dv : p.dk;
      const int64_t ldd = hf ? p.lddv : p.lddk;
      float* db = hf ? p.dbv : p.dbk;
      if (k_in && dst != nullptr) store_row32_bf16(dst + ((int64_t)b * Lk + kj) * ldd + h * DH, rk, 1.0f);
      if (db != nullptr) add_bias_grad(db + h * DH, rk, k_in, lane);
    }
    __threadfence();                                      // this thread's red.global.add are ordered before the counter bump below
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(FB_TMEM_COLS) : "memory");
  }
  // ---- the last key-tile CTA of this (b, h) converts the accumulated dQ columns of the head and clears them
  int* cnt = dq_count + (int64_t)b * p.H + h;
  if (threadIdx.x == 0) {
    const int old = atomicAdd(cnt, 1);
    bars->is_last = (old == (int)gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
#ifdef MMI_FB_NOFINISH
  if (bars->is_last && threadIdx.x == 0) *cnt = 0;
  bars->is_last = 0;
#endif
  if (bars->is_last) {
    __threadfence();
    __nv_bfloat16* dqo = p.dq[blk];
    const int64_t ldq = p.lddq[blk];
    const int chunk = threadIdx.x & 7;                    // 4 of the head's 32 columns
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = threadIdx.x >> 3; q < p.Lq; q += FB_THREADS / 8) {
      float4* src = reinterpret_cast<float4*>(dq_acc + (((int64_t)b * p.Lq + q) * p.H + h) * DH) + chunk;
      const float4 v = __ldcg(src);
      __stcg(src, make_float4(0.f, 0.f, 0.f, 0.f));
      cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
      if (dqo != nullptr)
        *reinterpret_cast<uint2*>(dqo + ((int64_t)b * p.Lq + q) * ldq + h * DH + chunk * 4) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    if (p.dbq[blk] != nullptr) {
      atomicAdd(&bars->colsum[chunk * 4 + 0], cs.x);
      atomicAdd(&bars->colsum[chunk * 4 + 1], cs.y);
      atomicAdd(&bars->colsum[chunk * 4 + 2], cs.z);
      atomicAdd(&bars->colsum[chunk * 4 + 3], cs.w);
      __syncthreads();
      if (threadIdx.x < DH) atomicAdd(p.dbq[blk] + h * DH + threadIdx.x, bars->colsum[threadIdx.x]);
    }
    if (threadIdx.x == 0) *cnt = 0;
  }
}
