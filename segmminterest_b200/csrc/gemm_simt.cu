// Strict-parity GEMM: fp32 FFMA on CUDA cores, 128x128x16 CTA tile, 8x8 register tile.
// This is the MMI_IMPL_SIMT path of mmi_gemm (fp32 mode, 1e-4 parity bar); the bf16
// performance path is the tcgen05 kernel in gemm_tc.cu.
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace mmi {

constexpr int BM = 128, BN = 128, BK = 16, PADM = 4;

template <typename TIN, typename TOUT, int LAYOUT>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmParams p) {
  __shared__ __align__(16) float As[BK][BM + PADM];
  __shared__ __align__(16) float Bs[BK][BN + PADM];
  const TIN* __restrict__ A = reinterpret_cast<const TIN*>(p.A);
  const TIN* __restrict__ B = reinterpret_cast<const TIN*>(p.B);
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int64_t n0 = (int64_t)blockIdx.x * BN;
  // split-K range
  const int64_t kchunk = ((p.K + p.split_k - 1) / p.split_k + BK - 1) / BK * BK;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  const int64_t kend = min(p.K, kbeg + kchunk);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  constexpr bool A_KMAJOR = (LAYOUT == MMI_GEMM_NT || LAYOUT == MMI_GEMM_NN);  // A [M,K]
  constexpr bool B_KMAJOR = (LAYOUT == MMI_GEMM_NT);                           // B [N,K]

  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = t + 256 * i;
      if (A_KMAJOR) {
        const int row = f >> 2, kv = f & 3;
        const int64_t m = m0 + row, k = k0 + kv * 4;
        ra[i] = (m < p.M && k < kend) ? load4(A + m * p.lda + k) : make_float4(0, 0, 0, 0);
      } else {
        const int kk = f >> 5, mv = f & 31;
        const int64_t k = k0 + kk, m = m0 + mv * 4;
        ra[i] = (k < kend && m < p.M) ? load4(A + k * p.lda + m) : make_float4(0, 0, 0, 0);
      }
      if (B_KMAJOR) {
        const int row = f >> 2, kv = f & 3;
        const int64_t n = n0 + row, k = k0 + kv * 4;
        rb[i] = (n < p.N && k < kend) ? load4(B + n * p.ldb + k) : make_float4(0, 0, 0, 0);
      } else {
        const int kk = f >> 5, nv = f & 31;
        const int64_t k = k0 + kk, n = n0 + nv * 4;
        rb[i] = (k < kend && n < p.N) ? load4(B + k * p.ldb + n) : make_float4(0, 0, 0, 0);
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = t + 256 * i;
      if (A_KMAJOR) {
        const int row = f >> 2, kv = f & 3;
        As[kv * 4 + 0][row] = ra[i].x; As[kv * 4 + 1][row] = ra[i].y;
        As[kv * 4 + 2][row] = ra[i].z; As[kv * 4 + 3][row] = ra[i].w;
      } else {
        const int kk = f >> 5, mv = f & 31;
        *reinterpret_cast<float4*>(&As[kk][mv * 4]) = ra[i];
      }
      if (B_KMAJOR) {
        const int row = f >> 2, kv = f & 3;
        Bs[kv * 4 + 0][row] = rb[i].x; Bs[kv * 4 + 1][row] = rb[i].y;
        Bs[kv * 4 + 2][row] = rb[i].z; Bs[kv * 4 + 3][row] = rb[i].w;
      } else {
        const int kk = f >> 5, nv = f & 31;
        *reinterpret_cast<float4*>(&Bs[kk][nv * 4]) = rb[i];
      }
    }
  };

  if (kbeg < kend) {
    load_tiles(kbeg);
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
      store_tiles();
      __syncthreads();
      if (k0 + BK < kend) load_tiles(k0 + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  const bool lead = blockIdx.z == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t n = n0 + h * 64 + tx * 4;
      if (n >= p.N) continue;
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      gemm_epilogue4<TIN, TOUT>(p, m, n, v, lead);
    }
  }
}

template <typename TIN, typename TOUT>
static int launch_simt(const GemmParams& p, cudaStream_t st) {
  dim3 grid((unsigned)((p.N + BN - 1) / BN), (unsigned)((p.M + BM - 1) / BM), (unsigned)p.split_k);
  switch (p.layout) {
    case MMI_GEMM_NT: gemm_simt_kernel<TIN, TOUT, MMI_GEMM_NT><<<grid, 256, 0, st>>>(p); break;
    case MMI_GEMM_NN: gemm_simt_kernel<TIN, TOUT, MMI_GEMM_NN><<<grid, 256, 0, st>>>(p); break;
    case MMI_GEMM_TN: gemm_simt_kernel<TIN, TOUT, MMI_GEMM_TN><<<grid, 256, 0, st>>>(p); break;
    default: set_error("gemm: bad layout %d", p.layout); return MMI_EINVAL;
  }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

int gemm_simt(const GemmParams& p, cudaStream_t st) {
  if (p.in_dtype == MMI_F32 && p.out_dtype == MMI_F32) return launch_simt<float, float>(p, st);
  if (p.in_dtype == MMI_BF16 && p.out_dtype == MMI_BF16) return launch_simt<__nv_bfloat16, __nv_bfloat16>(p, st);
  if (p.in_dtype == MMI_BF16 && p.out_dtype == MMI_F32) return launch_simt<__nv_bfloat16, float>(p, st);
  set_error("gemm(simt): unsupported dtypes in=%d out=%d", p.in_dtype, p.out_dtype);
  return MMI_EINVAL;
}

}  // namespace mmi
