// Fused GEMM epilogue shared by the SIMT and tcgen05 GEMM kernels (see mmi_gemm in
// include/mmi_b200.h for the op order).
#pragma once
#include "common.cuh"

namespace mmi {

struct GemmParams {
  int layout, impl, in_dtype, out_dtype;
  int64_t M, N, K;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
  const float* bias;
  int act;
  void* preact; int64_t ld_preact;
  const void* mul_gelu_grad; int64_t ld_mul;
  const void* add; int64_t ld_add; int64_t add_mod; int add_dtype;
  int accumulate;
  int split_k;
};

// 4 consecutive columns n..n+3 of row m.  `lead` = this CTA owns the bias/add terms
// (split-K slice 0).
template <typename TIN, typename TOUT>
__device__ __forceinline__ void gemm_epilogue4(const GemmParams& p, int64_t m, int64_t n, float (&v)[4], bool lead) {
  if (lead && p.bias != nullptr) {
    const float4 b = *reinterpret_cast<const float4*>(p.bias + n);
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (p.preact != nullptr) store4(reinterpret_cast<TIN*>(p.preact) + m * p.ld_preact + n, make_float4(v[0], v[1], v[2], v[3]));
  if (p.act == MMI_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_f(v[j]);
  }
  if (p.mul_gelu_grad != nullptr) {
    const float4 z = load4(reinterpret_cast<const TIN*>(p.mul_gelu_grad) + m * p.ld_mul + n);
    v[0] *= gelu_grad_f(z.x); v[1] *= gelu_grad_f(z.y); v[2] *= gelu_grad_f(z.z); v[3] *= gelu_grad_f(z.w);
  }
  if (lead && p.add != nullptr) {
    const int64_t r = m % p.add_mod;
    float4 a;
    if (p.add_dtype == MMI_F32) a = load4(reinterpret_cast<const float*>(p.add) + r * p.ld_add + n);
    else a = load4(reinterpret_cast<const __nv_bfloat16*>(p.add) + r * p.ld_add + n);
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
  }
  TOUT* c = reinterpret_cast<TOUT*>(p.C) + m * p.ldc + n;
  if (p.accumulate) {
    if constexpr (sizeof(TOUT) == 4) {
      float* cf = reinterpret_cast<float*>(c);
      if (p.split_k > 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(cf + j, v[j]);
      } else {
        float4 o = *reinterpret_cast<float4*>(cf);
        o.x += v[0]; o.y += v[1]; o.z += v[2]; o.w += v[3];
        *reinterpret_cast<float4*>(cf) = o;
      }
    }
  } else {
    store4(c, make_float4(v[0], v[1], v[2], v[3]));
  }
}

int gemm_simt(const GemmParams& p, cudaStream_t st);
int gemm_tc(const GemmParams& p, cudaStream_t st);
bool tc_available();

}  // namespace mmi
