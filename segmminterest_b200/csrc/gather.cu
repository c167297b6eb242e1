// a-1..a-3: segment-embedding gather + zero-pad + mask + L1 normalise.
// Replaces utils/dataloader_SegMM.py:301-350 (+ _pad_feature_list :251-268) and
// main_for_seq_leave_earlystop_SegMM.py:272-273.  HBM-bound: one warp owns one output
// row, issues all of its 128-bit streaming loads up front (MLP = NV per lane), reduces
// |x| with shuffles and writes the normalised row with 128-bit (or 64-bit, bf16) stores.
#include "common.cuh"

namespace mmi {

template <typename T> struct Vec16;  // 16-byte vector of T
template <> struct Vec16<float> { static constexpr int n = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int n = 8; };

template <typename TT> __device__ __forceinline__ void unpack(const uint4& r, float* f);
template <> __device__ __forceinline__ void unpack<float>(const uint4& r, float* f) {
  f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
}
template <> __device__ __forceinline__ void unpack<__nv_bfloat16>(const uint4& r, float* f) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <typename TO, int N> __device__ __forceinline__ void store_row_vec(TO* dst, const float* f);
template <> __device__ __forceinline__ void store_row_vec<float, 4>(float* dst, const float* f) {
  stg_stream(reinterpret_cast<uint4*>(dst), make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
}
template <> __device__ __forceinline__ void store_row_vec<float, 8>(float* dst, const float* f) {
  store_row_vec<float, 4>(dst, f);
  store_row_vec<float, 4>(dst + 4, f + 4);
}
template <> __device__ __forceinline__ void store_row_vec<__nv_bfloat16, 4>(__nv_bfloat16* dst, const float* f) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
  uint2 r = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  *reinterpret_cast<uint2*>(dst) = r;
}
template <> __device__ __forceinline__ void store_row_vec<__nv_bfloat16, 8>(__nv_bfloat16* dst, const float* f) {
  __nv_bfloat162 p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  stg_stream(reinterpret_cast<uint4*>(dst), make_uint4(*reinterpret_cast<uint32_t*>(&p[0]), *reinterpret_cast<uint32_t*>(&p[1]),
                                                        *reinterpret_cast<uint32_t*>(&p[2]), *reinterpret_cast<uint32_t*>(&p[3])));
}

// NV = 16-byte vectors per lane (row = up to 32*NV vectors)
template <typename TT, typename TO, int NV>
__global__ void __launch_bounds__(256) gather_l1norm_kernel(const TT* __restrict__ table, int64_t n_rows, int din,
                                                            const int32_t* __restrict__ idx, int64_t n_tokens,
                                                            TO* __restrict__ out, uint8_t* __restrict__ mask, int normalise) {
  constexpr int E = Vec16<TT>::n;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nvec = din / E;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_tokens; row += warps_total) {
    const int32_t r = __ldg(idx + row);
    const bool valid = r >= 0 && (int64_t)r < n_rows;
    TO* orow = out + row * (int64_t)din;
    if (lane == 0 && mask != nullptr) mask[row] = valid ? 1 : 0;
    uint4 v[NV];
    if (valid) {
      const uint4* src = reinterpret_cast<const uint4*>(table + (int64_t)r * din);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        v[i] = (c < nvec) ? ldg_stream(src + c) : make_uint4(0, 0, 0, 0);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = make_uint4(0, 0, 0, 0);
    }
    float denom = 1.0f;
    if (normalise) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float f[E];
        unpack<TT>(v[i], f);
#pragma unroll
        for (int j = 0; j < E; ++j) s += fabsf(f[j]);
      }
      s = warp_sum(s);
      denom = s + 1e-6f;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float f[E];
        unpack<TT>(v[i], f);
        if (normalise) {
#pragma unroll
          for (int j = 0; j < E; ++j) f[j] = __fdiv_rn(f[j], denom);  // same IEEE divide as the reference's x / (norm + 1e-6)
        }
        store_row_vec<TO, E>(orow + c * E, f);
      }
    }
  }
}

template <typename TT, typename TO, int NV>
static int launch_gather(const void* table, int64_t n_rows, int din, const int32_t* idx, int64_t n_tokens, void* out,
                         uint8_t* mask, int normalise, cudaStream_t st) {
  const int warps_per_cta = 8;
  int64_t ctas = (n_tokens + warps_per_cta - 1) / warps_per_cta;
  const int64_t cap = (int64_t)kNumSMs * 8;  // 8 resident CTAs of 256 threads per SM
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  gather_l1norm_kernel<TT, TO, NV><<<(unsigned)ctas, warps_per_cta * 32, 0, st>>>(
      reinterpret_cast<const TT*>(table), n_rows, din, idx, n_tokens, reinterpret_cast<TO*>(out), mask, normalise);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

template <typename TT, typename TO>
static int dispatch_nv(const void* table, int64_t n_rows, int din, const int32_t* idx, int64_t n_tokens, void* out,
                       uint8_t* mask, int normalise, cudaStream_t st) {
  const int nvec = din / Vec16<TT>::n;
  const int nv = (nvec + 31) / 32;
  switch (nv) {
    case 1: return launch_gather<TT, TO, 1>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
    case 2: return launch_gather<TT, TO, 2>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
    case 3: return launch_gather<TT, TO, 3>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
    case 4: return launch_gather<TT, TO, 4>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
    case 5: return launch_gather<TT, TO, 5>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
    case 6: return launch_gather<TT, TO, 6>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
    case 7: case 8: return launch_gather<TT, TO, 8>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
    default:
      set_error("gather: din=%d too wide (max %d)", din, 8 * 32 * Vec16<TT>::n);
      return MMI_ENOSUP;
  }
}

}  // namespace mmi

extern "C" int mmi_gather_l1norm_fwd(const void* table, int table_dtype, int64_t n_rows, int din, const int32_t* idx,
                                     int64_t n_tokens, void* out, int out_dtype, uint8_t* mask, int normalise,
                                     mmi_stream_t stream) {
  using namespace mmi;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(table && idx && out, "gather: null pointer");
  MMI_CHECK_ARG(n_tokens >= 0 && n_rows > 0 && din > 0, "gather: bad sizes");
  if (n_tokens == 0) return MMI_OK;
  const int e = table_dtype == MMI_F32 ? 4 : 8;
  MMI_CHECK_ARG(din % e == 0, "gather: din=%d must be a multiple of %d for 128-bit loads", din, e);
  MMI_CHECK_ARG((reinterpret_cast<uintptr_t>(table) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "gather: table/out must be 16-byte aligned");
  if (table_dtype == MMI_F32 && out_dtype == MMI_F32) return dispatch_nv<float, float>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
  if (table_dtype == MMI_F32 && out_dtype == MMI_BF16) return dispatch_nv<float, __nv_bfloat16>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
  if (table_dtype == MMI_BF16 && out_dtype == MMI_BF16) return dispatch_nv<__nv_bfloat16, __nv_bfloat16>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
  if (table_dtype == MMI_BF16 && out_dtype == MMI_F32) return dispatch_nv<__nv_bfloat16, float>(table, n_rows, din, idx, n_tokens, out, mask, normalise, st);
  set_error("gather: unsupported dtype combination %d -> %d", table_dtype, out_dtype);
  return MMI_EINVAL;
}
