// the reference's default multi-modal configuration.
//   * ID-embedding inputs (models/encoder.py:352-362, 426-435, 478-488): the video id, repeated over the 40 segments,
//     is embedded into d/2 columns; the other d/2 come from Linear(1 -> d/2) of the segment position; the user id is
//     ONE token embedded into d columns.  Position embedding is added here so that the output is the LayerNorm input.
//   * InteractionAggregation (models/decoder_leave_focal.py:392-423): w_x.x + w_y.y + sum_h x_h^T W_h y_h.  The two
//     bilinear products X_h W_h are tensor-core GEMMs (mmi_gemm); what is left is a row-wise dot product.
// All of it is a few MB per step: simple warp-per-row / thread-per-column kernels.
#include "common.cuh"

namespace mmi {

// out[b, l, c] = (c < tw ? table[id_b, c] : pos(b, l) * frame_w[c - tw] + frame_b[c - tw]) + pe[l, c]
// pos(b, l) = l, or frame_pos[b, l] when given (the 'noPos' ablation feeds a random permutation of 0..L-1 per row, encoder.py:428-429)
template <typename TO>
__global__ void __launch_bounds__(256) id_embed_fwd_kernel(const float* __restrict__ table, int64_t n_rows, int tw, const int64_t* __restrict__ ids,
                                                           int B, int L, int d, const float* __restrict__ frame_w,
                                                           const float* __restrict__ frame_b, const float* __restrict__ pe,
                                                           const float* __restrict__ frame_pos, TO* __restrict__ out) {
  const int64_t total = (int64_t)B * L * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d);
    const int l = (int)((i / d) % L);
    const int b = (int)(i / ((int64_t)d * L));
    float v;
    if (c < tw) {
      int64_t r = ids[b];
      r = r < 0 ? 0 : (r >= n_rows ? n_rows - 1 : r);       // torch would raise; the C ABI clamps (caller validates)
      v = table[r * tw + c];
    } else {
      v = (frame_pos ? frame_pos[(int64_t)b * L + l] : (float)l) * frame_w[c - tw] + frame_b[c - tw];
    }
    if (pe != nullptr) v += pe[(int64_t)l * d + c];
    out[i] = from_f32<TO>(v);
  }
}

// dtable[id_b, c] += sum_l de[b, l, c]  (c < tw);  dframe_w[c'] += sum_{b,l} l * de[b, l, tw + c'];  dframe_b[c'] += sum de
// grid = B CTAs, thread per column; duplicates of an id across the batch meet in atomicAdd.
template <typename TI>
__global__ void __launch_bounds__(256) id_embed_bwd_kernel(const TI* __restrict__ de, const int64_t* __restrict__ ids, int64_t n_rows, int tw,
                                                           int B, int L, int d, float* __restrict__ dtable, float* __restrict__ dframe_w,
                                                           float* __restrict__ dframe_b, const float* __restrict__ frame_pos) {
  const int b = blockIdx.x;
  int64_t r = ids[b];
  r = r < 0 ? 0 : (r >= n_rows ? n_rows - 1 : r);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f, sl = 0.f;
    for (int l = 0; l < L; ++l) {
      const float g = to_f32(de[((int64_t)b * L + l) * d + c]);
      s += g;
      sl += (frame_pos ? frame_pos[(int64_t)b * L + l] : (float)l) * g;
    }
    if (c < tw) atomicAdd(dtable + r * tw + c, s);
    else {
      if (dframe_w) atomicAdd(dframe_w + (c - tw), sl);
      if (dframe_b) atomicAdd(dframe_b + (c - tw), s);
    }
  }
}

// Row-sparse form of the table gradient (data parallelism): rows[b, c] = sum_l de[b, l, c] for c < tw -- ONE row
// per interaction, no atomics on the table -- so that ranks exchange (ids, rows) instead of all-reducing a dense
// [n_rows, tw] table (361 MB for the reference's 352 495 x 256 video table).  The frame projection's gradients are dense
// parameters and are accumulated as in id_embed_bwd_kernel.
template <typename TI>
__global__ void __launch_bounds__(256) id_rows_bwd_kernel(const TI* __restrict__ de, int tw, int B, int L, int d, float* __restrict__ rows,
                                                          float* __restrict__ dframe_w, float* __restrict__ dframe_b,
                                                          const float* __restrict__ frame_pos) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f, sl = 0.f;
    for (int l = 0; l < L; ++l) {
      const float g = to_f32(de[((int64_t)b * L + l) * d + c]);
      s += g;
      sl += (frame_pos ? frame_pos[(int64_t)b * L + l] : (float)l) * g;
    }
    if (c < tw) rows[(int64_t)b * tw + c] = s;
    else {
      if (dframe_w) atomicAdd(dframe_w + (c - tw), sl);
      if (dframe_b) atomicAdd(dframe_b + (c - tw), s);
    }
  }
}
// dtable[ids[i], :] += rows[i, :] for i < n (duplicates of an id meet in atomicAdd); ids are clamped like id_embed_fwd
__global__ void __launch_bounds__(256) scatter_rows_add_kernel(const int64_t* __restrict__ ids, const float* __restrict__ rows, int64_t n, int tw,
                                                               int64_t n_rows, float* __restrict__ dtable) {
  const int64_t total = n * tw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r_ = i / tw;
    const int c = (int)(i - r_ * tw);
    int64_t r = ids[r_];
    r = r < 0 ? 0 : (r >= n_rows ? n_rows - 1 : r);
    atomicAdd(dtable + r * tw + c, rows[i]);
  }
}

// out[r] = sum_c T[r,c] * Y[r,c] (+ add1[r]) (+ add2[r]);  warp per row, C % 4 == 0
template <typename T>
__global__ void __launch_bounds__(256) rowdot_fwd_kernel(const T* __restrict__ t, int64_t ldt, const T* __restrict__ y, int64_t ldy, int64_t R,
                                                         int C, const float* __restrict__ add1, const float* __restrict__ add2,
                                                         float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < R; row += wstride) {
    float s = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 a = load4(t + row * ldt + c), b = load4(y + row * ldy + c);
      s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    s = warp_sum(s);
    if (lane == 0) out[row] = s + (add1 ? add1[row] : 0.f) + (add2 ? add2[row] : 0.f);
  }
}

// dT[r,c] = g[r] * Y[r,c];   dY[r,c] = g[r] * T[r,c] (+ dy_add[r,c])
template <typename T>
__global__ void __launch_bounds__(256) rowdot_bwd_kernel(const float* __restrict__ g, const float* __restrict__ gscale, const T* __restrict__ t,
                                                         int64_t ldt, const T* __restrict__ y, int64_t ldy, int64_t R, int C,
                                                         const T* __restrict__ dy_add, T* __restrict__ dt, T* __restrict__ dy) {
  const int lane = threadIdx.x & 31;
  const float gs = gscale ? gscale[0] : 1.0f;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < R; row += wstride) {
    const float gr = g[row] * gs;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 a = load4(t + row * ldt + c), b = load4(y + row * ldy + c);
      store4(dt + row * (int64_t)C + c, make_float4(gr * b.x, gr * b.y, gr * b.z, gr * b.w));
      float4 o = make_float4(gr * a.x, gr * a.y, gr * a.z, gr * a.w);
      if (dy_add != nullptr) {
        const float4 e = load4(dy_add + row * (int64_t)C + c);
        o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
      }
      store4(dy + row * (int64_t)C + c, o);
    }
  }
}

// AdaptiveAvgPool1d(P) along the token axis (the CrossMLP ablation, models/encoder.py:395,504-506): window j of a length-L
// axis is [floor(j L / P), ceil((j + 1) L / P)) -- torch's definition; windows overlap when L is not a multiple of P.
__device__ __forceinline__ int pool_start(int j, int L, int P) { return (int)(((int64_t)j * L) / P); }
__device__ __forceinline__ int pool_end(int j, int L, int P) { return (int)((((int64_t)j + 1) * L + P - 1) / P); }

template <typename T>
__global__ void __launch_bounds__(256) adaptive_pool_fwd_kernel(const T* __restrict__ x, int B, int L, int d, int P, T* __restrict__ y) {
  const int64_t total = (int64_t)B * P * (d / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (d / 4)) * 4;
    const int j = (int)((i / (d / 4)) % P);
    const int b = (int)(i / ((int64_t)(d / 4) * P));
    const int s = pool_start(j, L, P), e = pool_end(j, L, P);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = s; t < e; ++t) {
      const float4 v = load4(x + ((int64_t)b * L + t) * d + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float inv = 1.0f / (float)(e - s);
    store4(y + ((int64_t)b * P + j) * d + c, make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) adaptive_pool_bwd_kernel(const T* __restrict__ dy, int B, int L, int d, int P, T* __restrict__ dx) {
  const int64_t total = (int64_t)B * L * (d / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (d / 4)) * 4;
    const int t = (int)((i / (d / 4)) % L);
    const int b = (int)(i / ((int64_t)(d / 4) * L));
    // windows that contain t: j with floor(j L / P) <= t < ceil((j + 1) L / P); start from a lower bound and scan
    int j = (int)(((int64_t)t * P) / L);
    while (j > 0 && pool_end(j - 1, L, P) > t) --j;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (; j < P && pool_start(j, L, P) <= t; ++j) {
      const int e = pool_end(j, L, P), s0 = pool_start(j, L, P);
      if (t >= e) continue;
      const float inv = 1.0f / (float)(e - s0);
      const float4 g = load4(dy + ((int64_t)b * P + j) * d + c);
      acc.x += g.x * inv; acc.y += g.y * inv; acc.z += g.z * inv; acc.w += g.w * inv;
    }
    store4(dx + ((int64_t)b * L + t) * d + c, acc);
  }
}

}  // namespace mmi

using namespace mmi;

static int pool_launch(bool fwd, const void* in, int dtype, int B, int L, int d, int P, void* out, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(in && out, "adaptive_pool: null pointer");
  MMI_CHECK_ARG(B > 0 && L > 0 && P > 0 && d > 0 && d % 4 == 0, "adaptive_pool: need B, L, out_len > 0 and d a positive multiple of 4");
  const int64_t total = (int64_t)B * (fwd ? P : L) * (d / 4);
  int64_t grid = (total + 255) / 256;
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  if (dtype == MMI_F32) {
    if (fwd) adaptive_pool_fwd_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float*)in, B, L, d, P, (float*)out);
    else adaptive_pool_bwd_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float*)in, B, L, d, P, (float*)out);
  } else if (dtype == MMI_BF16) {
    if (fwd) adaptive_pool_fwd_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>((const __nv_bfloat16*)in, B, L, d, P, (__nv_bfloat16*)out);
    else adaptive_pool_bwd_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>((const __nv_bfloat16*)in, B, L, d, P, (__nv_bfloat16*)out);
  } else { set_error("adaptive_pool: bad dtype %d", dtype); return MMI_EINVAL; }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}
extern "C" int mmi_adaptive_pool_fwd(const void* x, int dtype, int B, int L, int d, int out_len, void* y, mmi_stream_t stream) {
  return pool_launch(true, x, dtype, B, L, d, out_len, y, stream);
}
extern "C" int mmi_adaptive_pool_bwd(const void* dy, int dtype, int B, int L, int d, int out_len, void* dx, mmi_stream_t stream) {
  return pool_launch(false, dy, dtype, B, L, d, out_len, dx, stream);
}

extern "C" int mmi_id_embed_fwd(const float* table, int64_t n_rows, int tw, const int64_t* ids, int B, int L, int d,
                                const float* frame_w, const float* frame_b, const float* pe, const float* frame_pos, void* out,
                                int out_dtype, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(table && ids && out, "id_embed_fwd: null pointer");
  MMI_CHECK_ARG(B > 0 && L > 0 && d > 0 && tw > 0 && tw <= d && n_rows > 0, "id_embed_fwd: bad sizes");
  MMI_CHECK_ARG(tw == d || (frame_w && frame_b), "id_embed_fwd: columns [tw, d) need the frame projection");
  const int64_t total = (int64_t)B * L * d;
  int64_t grid = (total + 255) / 256;
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (out_dtype == MMI_F32) id_embed_fwd_kernel<float><<<(unsigned)grid, 256, 0, st>>>(table, n_rows, tw, ids, B, L, d, frame_w, frame_b, pe, frame_pos, (float*)out);
  else if (out_dtype == MMI_BF16) id_embed_fwd_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>(table, n_rows, tw, ids, B, L, d, frame_w, frame_b, pe, frame_pos, (__nv_bfloat16*)out);
  else { set_error("id_embed_fwd: bad dtype %d", out_dtype); return MMI_EINVAL; }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_id_embed_bwd(const void* de, int dtype, const int64_t* ids, int64_t n_rows, int tw, int B, int L, int d,
                                float* dtable, float* dframe_w, float* dframe_b, const float* frame_pos, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(de && ids && dtable, "id_embed_bwd: null pointer");
  MMI_CHECK_ARG(B > 0 && L > 0 && d > 0 && tw > 0 && tw <= d && n_rows > 0, "id_embed_bwd: bad sizes");
  if (dtype == MMI_F32) id_embed_bwd_kernel<float><<<B, 256, 0, st>>>((const float*)de, ids, n_rows, tw, B, L, d, dtable, dframe_w, dframe_b, frame_pos);
  else if (dtype == MMI_BF16) id_embed_bwd_kernel<__nv_bfloat16><<<B, 256, 0, st>>>((const __nv_bfloat16*)de, ids, n_rows, tw, B, L, d, dtable, dframe_w, dframe_b, frame_pos);
  else { set_error("id_embed_bwd: bad dtype %d", dtype); return MMI_EINVAL; }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_id_rows_bwd(const void* de, int dtype, int tw, int B, int L, int d, float* rows, float* dframe_w, float* dframe_b,
                               const float* frame_pos, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(de && rows, "id_rows_bwd: null pointer");
  MMI_CHECK_ARG(B > 0 && L > 0 && d > 0 && tw > 0 && tw <= d, "id_rows_bwd: bad sizes");
  if (dtype == MMI_F32) id_rows_bwd_kernel<float><<<B, 256, 0, st>>>((const float*)de, tw, B, L, d, rows, dframe_w, dframe_b, frame_pos);
  else if (dtype == MMI_BF16) id_rows_bwd_kernel<__nv_bfloat16><<<B, 256, 0, st>>>((const __nv_bfloat16*)de, tw, B, L, d, rows, dframe_w, dframe_b, frame_pos);
  else { set_error("id_rows_bwd: bad dtype %d", dtype); return MMI_EINVAL; }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_scatter_rows_add(const int64_t* ids, const float* rows, int64_t n, int tw, int64_t n_rows, float* dtable, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(ids && rows && dtable, "scatter_rows_add: null pointer");
  MMI_CHECK_ARG(n >= 0 && tw > 0 && n_rows > 0, "scatter_rows_add: bad sizes");
  if (n == 0) return MMI_OK;
  int64_t grid = (n * tw + 255) / 256;
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  scatter_rows_add_kernel<<<(unsigned)grid, 256, 0, st>>>(ids, rows, n, tw, n_rows, dtable);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_rowdot_fwd(const void* t, int64_t ldt, const void* y, int64_t ldy, int dtype, int64_t R, int C, const float* add1,
                              const float* add2, float* out, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(t && y && out, "rowdot_fwd: null pointer");
  MMI_CHECK_ARG(C > 0 && C % 4 == 0 && ldt % 4 == 0 && ldy % 4 == 0, "rowdot_fwd: C, ldt, ldy must be multiples of 4");
  if (R == 0) return MMI_OK;
  int64_t grid = (R + 7) / 8;
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (dtype == MMI_F32) rowdot_fwd_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float*)t, ldt, (const float*)y, ldy, R, C, add1, add2, out);
  else if (dtype == MMI_BF16) rowdot_fwd_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>((const __nv_bfloat16*)t, ldt, (const __nv_bfloat16*)y, ldy, R, C, add1, add2, out);
  else { set_error("rowdot_fwd: bad dtype %d", dtype); return MMI_EINVAL; }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

extern "C" int mmi_rowdot_bwd(const float* g, const float* gscale, const void* t, int64_t ldt, const void* y, int64_t ldy, int dtype,
                              int64_t R, int C, const void* dy_add, void* dt, void* dy, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(g && t && y && dt && dy, "rowdot_bwd: null pointer");
  MMI_CHECK_ARG(C > 0 && C % 4 == 0 && ldt % 4 == 0 && ldy % 4 == 0, "rowdot_bwd: C, ldt, ldy must be multiples of 4");
  if (R == 0) return MMI_OK;
  int64_t grid = (R + 7) / 8;
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (dtype == MMI_F32) rowdot_bwd_kernel<float><<<(unsigned)grid, 256, 0, st>>>(g, gscale, (const float*)t, ldt, (const float*)y, ldy, R, C, (const float*)dy_add, (float*)dt, (float*)dy);
  else if (dtype == MMI_BF16) rowdot_bwd_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>(g, gscale, (const __nv_bfloat16*)t, ldt, (const __nv_bfloat16*)y, ldy, R, C, (const __nv_bfloat16*)dy_add, (__nv_bfloat16*)dt, (__nv_bfloat16*)dy);
  else { set_error("rowdot_bwd: bad dtype %d", dtype); return MMI_EINVAL; }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}
