// Fused GEMM epilogue shared by the SIMT and tcgen05 GEMM kernels (see mmi_gemm in
// include/mmi_b200.h for the op order).
#pragma once
#include "common.cuh"
#include "dropout.cuh"

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
  int save_act_grad;
  int mul_is_grad;
  DropParams drop;     // dropout of act(z) before `add` (and of the saved gelu'(z)); thr8 = 0: off
  float mul_scale;     // mul_is_grad == 2: y *= (mul > 0 ? mul_scale : 0)
  void* split_ws = nullptr; int64_t split_ws_bytes = 0;   // fp32 operands on the tensor cores (three-term bf16 split)
  int split_terms = 0;  // tcgen05 kernel: 6 = the K loop walks six (A term, B term) segment pairs of seg_kb k-blocks each
  int seg_kb = 0;
};

// 4 consecutive columns n..n+3 of row m.  `lead` = this CTA owns the bias/add terms
// (split-K slice 0).
template <typename TIN, typename TOUT>
__device__ __forceinline__ void gemm_epilogue4(const GemmParams& p, int64_t m, int64_t n, float (&v)[4], bool lead) {
  if (lead && p.bias != nullptr) {
    const float4 b = *reinterpret_cast<const float4*>(p.bias + n);
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  float keep[4] = {1.f, 1.f, 1.f, 1.f};   // dropout factor of the four columns (0 or scale)
  if (p.drop.thr8 != 0u) {
    const uint32_t w = drop_keep_word(drop_rowhash(p.drop.key, (uint64_t)m), (uint32_t)(n >> 5), p.drop.thr8) >> (n & 31);
#pragma unroll
    for (int j = 0; j < 4; ++j) keep[j] = ((w >> j) & 1u) ? p.drop.scale : 0.f;
  }
  if (p.preact != nullptr) {
    float4 z = make_float4(v[0], v[1], v[2], v[3]);
    if (p.save_act_grad)   // the saved derivative carries the dropout factor: the backward epilogue multiplies by it
      z = make_float4(gelu_grad_f(v[0]) * keep[0], gelu_grad_f(v[1]) * keep[1], gelu_grad_f(v[2]) * keep[2], gelu_grad_f(v[3]) * keep[3]);
    store4(reinterpret_cast<TIN*>(p.preact) + m * p.ld_preact + n, z);
  }
  if (p.act == MMI_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_f(v[j]);
  } else if (p.act == MMI_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (p.drop.thr8 != 0u) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= keep[j];
  }
  if (p.mul_gelu_grad != nullptr) {
    const float4 z = load4(reinterpret_cast<const TIN*>(p.mul_gelu_grad) + m * p.ld_mul + n);
    if (p.mul_is_grad == 2) {
      v[0] = z.x > 0.f ? v[0] * p.mul_scale : 0.f; v[1] = z.y > 0.f ? v[1] * p.mul_scale : 0.f;
      v[2] = z.z > 0.f ? v[2] * p.mul_scale : 0.f; v[3] = z.w > 0.f ? v[3] * p.mul_scale : 0.f;
    } else if (p.mul_is_grad) { v[0] *= z.x; v[1] *= z.y; v[2] *= z.z; v[3] *= z.w; }
    else { v[0] *= gelu_grad_f(z.x); v[1] *= gelu_grad_f(z.y); v[2] *= gelu_grad_f(z.z); v[3] *= gelu_grad_f(z.w); }
  }
  if (lead && p.add != nullptr) {
    const int64_t r = (m < p.add_mod) ? m : m % p.add_mod;
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

// erf GELU and its derivative for the bf16 tensor-core epilogues, TWO elements at a time on the packed fp32x2 pipe with ONE
// MUFU op per element.  A K = 512 tile (128 x 256) gives the epilogue 4 096 tensor-pipe cycles for 32 768 elements: the
// previous form (Abramowitz-Stegun erf: ex2 + rcp + ~25 scalar FMA-pipe ops per element) needed 4 096 MUFU and ~7 000 issue
// cycles and ran the GELU GEMMs at 0.19 of the tensor roofline.  Here
//     Phi(x) ~ 1/2 (1 + tanh(u)),  u = x (a + b x^2 + c x^4)   on |x| <= 8 (x is clamped beyond: Phi = 0 / 1 to fp32 there)
// with a minimax fit of (a, b, c) against erf: |dPhi| <= 4.5e-5, |x dPhi| <= 4.1e-5, and the derivative of the approximant
//     gelu'(x) ~ Phi + 1/2 x (1 - tanh^2 u) u'(x)
// is within 9.5e-5 of Phi + x phi(x).  tanh.approx.f32 adds 2^-11 relative: everything stays below the 2^-9 of the bf16
// the results are rounded to.  The fp32 strict-parity path keeps erff (common.cuh gelu_f / gelu_grad_f).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kGa = 7.97548299e-01f, kGb = 3.69484188e-02f, kGc = -3.37469906e-04f;
// g = gelu(x), dg = gelu'(x) for the two elements of x
template <bool WANT_G, bool WANT_DG>
__device__ __forceinline__ void gelu_pair(float2 x, float2& g, float2& dg) {
  const float2 xc = make_float2(fminf(fmaxf(x.x, -8.0f), 8.0f), fminf(fmaxf(x.y, -8.0f), 8.0f));
  const float2 x2 = mul2(xc, xc);
  const float2 p = fma2(x2, fma2(x2, splat2(kGc), splat2(kGb)), splat2(kGa));
  const float2 u = mul2(xc, p);
  const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 h = fma2(t, splat2(0.5f), splat2(0.5f));                                   // Phi
  if (WANT_G) g = mul2(x, h);
  if (WANT_DG) {
    const float2 up = fma2(x2, fma2(x2, splat2(5.0f * kGc), splat2(3.0f * kGb)), splat2(kGa));   // u'(x)
    const float2 s = fma2(mul2(t, splat2(-1.0f)), t, splat2(1.0f));                              // 1 - t^2
    dg = fma2(mul2(mul2(xc, s), up), splat2(0.5f), h);
  }
}
__device__ __forceinline__ float gelu_fast(float x) { float2 g, d; gelu_pair<true, false>(make_float2(x, x), g, d); return g.x; }
__device__ __forceinline__ float gelu_grad_fast(float x) { float2 g, d; gelu_pair<false, true>(make_float2(x, x), g, d); return d.x; }

// 2 consecutive elements <-> float2 (coalesced row-wise epilogue of the tcgen05 kernel)
__device__ __forceinline__ float2 load2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load2(const __nv_bfloat16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
__device__ __forceinline__ void store2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void store2(__nv_bfloat16* p, float2 v) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y); }

// Columns n, n+1 of row m; consecutive lanes own consecutive column pairs of the SAME row, so
// every global access below is a fully coalesced 128 B (bf16) / 256 B (fp32) row segment.
// b0/b1: bias of the two columns (already zero when there is no bias or this is not the lead split).
// add_row: row of the `add` operand that belongs to output row m (m for a residual, m mod add_mod for the position
// table); the caller walks it incrementally -- a 64-bit modulo per element made the position-embedding GEMM 5x slower.
template <typename TIN, typename TOUT>
__device__ __forceinline__ void gemm_epilogue2(const GemmParams& p, int64_t m, int64_t n, float v0, float v1, float b0, float b1,
                                               bool lead, int64_t add_row) {
  v0 += b0; v1 += b1;
  float k0 = 1.f, k1 = 1.f;                  // dropout factors (generic register epilogue: one keep word per call)
  if (p.drop.thr8 != 0u) {
    const uint32_t w = drop_keep_word(drop_rowhash(p.drop.key, (uint64_t)m), (uint32_t)(n >> 5), p.drop.thr8) >> (n & 31);
    k0 = (w & 1u) ? p.drop.scale : 0.f;
    k1 = (w & 2u) ? p.drop.scale : 0.f;
  }
  // fp32 side operands = the strict-parity mode (split-bf16 products): exact erf GELU like the FFMA path
  constexpr bool kExact = sizeof(TIN) == 4;
  if (p.preact != nullptr) {
    float2 z = make_float2(v0, v1);
    if (p.save_act_grad) z = kExact ? make_float2(gelu_grad_f(v0) * k0, gelu_grad_f(v1) * k1) : make_float2(gelu_grad_fast(v0) * k0, gelu_grad_fast(v1) * k1);
    store2(reinterpret_cast<TIN*>(p.preact) + m * p.ld_preact + n, z);
  }
  if (p.act == MMI_ACT_GELU) {
    if constexpr (kExact) { v0 = gelu_f(v0); v1 = gelu_f(v1); }
    else { v0 = gelu_fast(v0); v1 = gelu_fast(v1); }
  }
  else if (p.act == MMI_ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
  v0 *= k0; v1 *= k1;
  if (p.mul_gelu_grad != nullptr) {
    const float2 z = load2(reinterpret_cast<const TIN*>(p.mul_gelu_grad) + m * p.ld_mul + n);
    if (p.mul_is_grad == 2) { v0 = z.x > 0.f ? v0 * p.mul_scale : 0.f; v1 = z.y > 0.f ? v1 * p.mul_scale : 0.f; }
    else if (p.mul_is_grad) { v0 *= z.x; v1 *= z.y; }
    else if constexpr (kExact) { v0 *= gelu_grad_f(z.x); v1 *= gelu_grad_f(z.y); }
    else { v0 *= gelu_grad_fast(z.x); v1 *= gelu_grad_fast(z.y); }
  }
  if (lead && p.add != nullptr) {
    float2 a;
    if (p.add_dtype == MMI_F32) a = load2(reinterpret_cast<const float*>(p.add) + add_row * p.ld_add + n);
    else a = load2(reinterpret_cast<const __nv_bfloat16*>(p.add) + add_row * p.ld_add + n);
    v0 += a.x; v1 += a.y;
  }
  TOUT* c = reinterpret_cast<TOUT*>(p.C) + m * p.ldc + n;
  if (p.accumulate) {
    if constexpr (sizeof(TOUT) == 4) {
      float* cf = reinterpret_cast<float*>(c);
      if (p.split_k > 1) {
        atomicAdd(reinterpret_cast<float2*>(cf), make_float2(v0, v1));  // red.global.add.v2.f32 (sm_90+)
      } else {
        float2 o = *reinterpret_cast<float2*>(cf);
        o.x += v0; o.y += v1;
        *reinterpret_cast<float2*>(cf) = o;
      }
    }
  } else {
    store2(c, make_float2(v0, v1));
  }
}

int gemm_simt(const GemmParams& p, cudaStream_t st);
int gemm_tc(const GemmParams& p, cudaStream_t st);
int64_t gemm_split_workspace(int layout, int64_t M, int64_t N, int64_t K);
bool tc_available();

}  // namespace mmi
