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
//   * persistent CTAs (one per SM), tile order n-fastest so an activation row-block is
//     fetched from HBM once and re-read from L2; split-K (atomic fp32) for the
//     weight-gradient GEMMs whose K dimension is the token count.
//
// Layouts: NT uses K-major operands (A [M,K], B [N,K]); TN (weight gradient, dW = dY^T X)
// uses MN-major operands straight from the row-major activations -- no transposes in HBM.
// NN is not needed: the engine keeps a transposed bf16 shadow of each weight for dgrad.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace mmi {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192; // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr uint32_t SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();  // a protocol bug must fail the launch, not hang the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B.
//   K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused
//   MN-major: 64-element (128 B) MN chunks `lbo_bytes` apart, 8-k-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);          // start address  [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;     // leading byte offset [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;     // stride byte offset  [32,46)
  d |= static_cast<uint64_t>(1) << 46;                             // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                             // SWIZZLE_128B
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

template <int BN, bool MN_MAJOR, typename TOUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p,
               int m_tiles, int n_tiles, int num_kb, int split) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // SWIZZLE_128B atoms must sit on 1024 B boundaries
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = m_tiles * n_tiles * split;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 1) {  // one warp owns TMEM alloc + dealloc; 2 accumulator buffers of BN fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileInfo ti = get_tile(t, m_tiles, n_tiles, num_kb, split);
        for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if constexpr (!MN_MAJOR) {
            tma_load_2d(&tma_a, &full_bar[stage], sa, kb * BK, ti.m_blk * BM);   // box {64 k, 128 rows}
            tma_load_2d(&tma_b, &full_bar[stage], sb, kb * BK, ti.n_blk * BN);   // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)                                    // boxes {64 mn, 64 k}
              tma_load_2d(&tma_a, &full_bar[stage], sa + j * (64 * BK * 2), ti.m_blk * BM + j * 64, kb * BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(&tma_b, &full_bar[stage], sb + j * (64 * BK * 2), ti.n_blk * BN + j * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, MN_MAJOR, MN_MAJOR);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const TileInfo ti = get_tile(t, m_tiles, n_tiles, num_kb, split);
        const uint32_t buf = it & 1, use = it >> 1;
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);   // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
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
            umma_f16(d_tmem, ad, bd, idesc, (kb > ti.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);             // smem slot free once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[buf]);                 // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (4 warps)
    const int q = warp & 3;                            // TMEM lane quarter this warp may read
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const TileInfo ti = get_tile(t, m_tiles, n_tiles, num_kb, split);
      const uint32_t buf = it & 1, use = it >> 1;
      mbar_wait(&tmem_full[buf], use & 1);
      tcgen05_fence_after();
      const int64_t m = static_cast<int64_t>(ti.m_blk) * BM + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + c * 32, r);
        tmem_ld_wait();
        if (m < p.M) {
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const int64_t n = static_cast<int64_t>(ti.n_blk) * BN + c * 32 + g4 * 4;
            if (n < p.N) {
              float v[4] = {__uint_as_float(r[g4 * 4]), __uint_as_float(r[g4 * 4 + 1]), __uint_as_float(r[g4 * 4 + 2]),
                            __uint_as_float(r[g4 * 4 + 3])};
              gemm_epilogue4<__nv_bfloat16, TOUT>(p, m, n, v, ti.lead);
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

struct MapKey {
  const void* ptr; uint64_t inner, outer, stride; uint32_t box_inner, box_outer;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && stride == o.stride && box_inner == o.box_inner && box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.stride); mix(k.box_inner); mix(k.box_outer);
    return h;
  }
};

// immutable per-(pointer, shape) descriptor cache; the only global state of the library
static bool get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t stride_elems, uint32_t box_inner,
                           uint32_t box_outer, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, inner, outer, stride_elems, box_inner, box_outer};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("gemm_tc: cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {stride_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu stride=%llu box=%ux%u", (int)r, ptr,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)stride_elems, box_inner, box_outer);
    return false;
  }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return true;
}

template <int BN, bool MN_MAJOR, typename TOUT>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int m_tiles, int n_tiles, int num_kb, int split,
                  cudaStream_t st) {
  constexpr size_t smem = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 /*align*/ + 256 /*barriers*/;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, MN_MAJOR, TOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return MMI_ECUDA; }
    configured = true;
  }
  const int total = m_tiles * n_tiles * split;
  const int grid = total < kNumSMs ? total : kNumSMs;
  gemm_tc_kernel<BN, MN_MAJOR, TOUT><<<grid, NUM_THREADS, smem, st>>>(ta, tb, p, m_tiles, n_tiles, num_kb, split);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

}  // namespace tc

bool tc_available() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    ok = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
        major == 10 && tc::get_encode() != nullptr)
      ok = 1;
    cudaGetLastError();
  }
  return ok == 1;
}

int gemm_tc(const GemmParams& p_in, cudaStream_t st) {
  using namespace tc;
  GemmParams p = p_in;
  MMI_CHECK_ARG(p.in_dtype == MMI_BF16, "gemm_tc: inputs must be bf16");
  MMI_CHECK_ARG(p.layout == MMI_GEMM_NT || p.layout == MMI_GEMM_TN, "gemm_tc: layouts NT and TN only (dgrad uses the transposed weight shadow)");
  MMI_CHECK_ARG((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.B) & 15) == 0, "gemm_tc: A/B must be 16-byte aligned");
  MMI_CHECK_ARG(p.lda % 8 == 0 && p.ldb % 8 == 0, "gemm_tc: lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
  const bool mn = p.layout == MMI_GEMM_TN;
  const int bn = (p.N % 256 == 0) ? 256 : 128;
  const int m_tiles = (int)((p.M + BM - 1) / BM), n_tiles = (int)((p.N + bn - 1) / bn);
  const int num_kb = (int)((p.K + BK - 1) / BK);
  int split = p.split_k;
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
    if (!get_tensor_map(p.A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)p.lda, BK, BM, &ta)) return MMI_ECUDA;
    if (!get_tensor_map(p.B, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)p.ldb, BK, (uint32_t)bn, &tb)) return MMI_ECUDA;
  } else {
    if (!get_tensor_map(p.A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.lda, 64, BK, &ta)) return MMI_ECUDA;
    if (!get_tensor_map(p.B, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.ldb, 64, BK, &tb)) return MMI_ECUDA;
  }
  const bool f32out = p.out_dtype == MMI_F32;
  MMI_CHECK_ARG(f32out || p.out_dtype == MMI_BF16, "gemm_tc: bad out dtype");
#define MMI_TC_LAUNCH(BN_, MN_)                                                                                         \
  (f32out ? launch<BN_, MN_, float>(ta, tb, p, m_tiles, n_tiles, num_kb, split, st)                                     \
          : launch<BN_, MN_, __nv_bfloat16>(ta, tb, p, m_tiles, n_tiles, num_kb, split, st))
  if (bn == 256) return mn ? MMI_TC_LAUNCH(256, true) : MMI_TC_LAUNCH(256, false);
  return mn ? MMI_TC_LAUNCH(128, true) : MMI_TC_LAUNCH(128, false);
#undef MMI_TC_LAUNCH
}

int attn_tc(int, const mmi_attn_args*, int, cudaStream_t) {
  set_error("attn_tc: not built yet");
  return MMI_ENOSUP;
}

}  // namespace mmi
