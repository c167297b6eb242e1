// a-5/a-6 strict-parity attention (MMI_IMPL_SIMT): one thread owns one query row (fwd, dq)
// or one key row (dk/dv); K/V (or Q/dO) tiles are staged in shared memory and read as
// warp-wide broadcasts; the two key blocks of the reference's joint softmax
// (models/encoder.py:138-161) are streamed through one online softmax, so the
// [B,16,Lq,Lk] logits the reference materialises in HBM four times never exist.
//
// Mask semantics follow models/encoder.py:65-71,146 exactly: a logit whose query OR key is
// padding is SET to -10000 (before the 1/sqrt(dh) scale), so a padded query row gets a
// uniform softmax over all keys (pads included) and gradients do not flow through the
// overwritten logits.
#include "common.cuh"
#include "dropout.cuh"

namespace mmi {

struct AttnBlk {
  const void* q; int64_t ldq;
  const void* k; int64_t ldk;
  const void* v; int64_t ldv;
  const uint8_t* mask_k; int Lk;
  void* dq; int64_t lddq;
  void* dk; int64_t lddk;
  void* dv; int64_t lddv;
};
struct AttnParams {
  int B, H, Lq, nblk;
  const uint8_t* mask_q;
  AttnBlk blk[2];
  void* out; int64_t ldo;
  float* lse;
  const void* dout; int64_t lddo;
  float* delta;
  float scale;
  DropParams drop;   // logits dropout (models/encoder.py:145-150): after the -10000 fill, before the scale
};
// keep-word group of key k of key block blk (include/mmi_b200.h)
__device__ __forceinline__ uint32_t attn_group(int blk, int k) { return (static_cast<uint32_t>(blk) << 20) + (static_cast<uint32_t>(k) >> 5); }

constexpr int kQThreads = 128;
constexpr int kKT = 64;   // keys per smem tile (fwd / dq)
constexpr int kQT = 32;   // queries per smem tile (dkv)
constexpr float kMaskFill = -10000.0f;

template <typename T, int DH>
__device__ __forceinline__ void load_row(const T* p, float (&r)[DH]) {
#pragma unroll
  for (int d = 0; d < DH; d += 4) {
    const float4 v = load4(p + d);
    r[d] = v.x; r[d + 1] = v.y; r[d + 2] = v.z; r[d + 3] = v.w;
  }
}
template <typename T, int DH>
__device__ __forceinline__ void store_row(T* p, const float (&r)[DH], float s) {
#pragma unroll
  for (int d = 0; d < DH; d += 4) store4(p + d, make_float4(r[d] * s, r[d + 1] * s, r[d + 2] * s, r[d + 3] * s));
}

// stage `n_rows` rows [row0, row0+n) x DH of a [B*L, ld] tensor (head offset applied) into smem as fp32
template <typename T, int DH, int ROWS, int THREADS>
__device__ __forceinline__ void stage_tile(float (*dst)[DH], const T* base, int64_t ld, int row0, int L, int tid) {
  constexpr int VPR = DH / 4;
  for (int f = tid; f < ROWS * VPR; f += THREADS) {
    const int r = f / VPR, c = (f % VPR) * 4;
    float4 v = make_float4(0, 0, 0, 0);
    if (row0 + r < L) v = load4(base + (int64_t)(row0 + r) * ld + c);
    *reinterpret_cast<float4*>(&dst[r][c]) = v;
  }
}

template <int DH>
__device__ __forceinline__ float dot_smem(const float (&q)[DH], const float* k) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < DH; d += 4) {
    const float4 kv = *reinterpret_cast<const float4*>(k + d);
    s = fmaf(q[d], kv.x, s); s = fmaf(q[d + 1], kv.y, s); s = fmaf(q[d + 2], kv.z, s); s = fmaf(q[d + 3], kv.w, s);
  }
  return s;
}

// ------------------------------------------------------------------------------ forward
template <typename T, int DH>
__global__ void __launch_bounds__(kQThreads) attn_fwd_simt_kernel(AttnParams a) {
  __shared__ __align__(16) float Ks[kKT][DH];
  __shared__ __align__(16) float Vs[kKT][DH];
  __shared__ uint8_t Mk[kKT];
  const int b = blockIdx.z, h = blockIdx.y, tid = threadIdx.x;
  const int qi = blockIdx.x * kQThreads + tid;
  const bool active = qi < a.Lq;
  const bool mq = active ? (a.mask_q[(int64_t)b * a.Lq + qi] != 0) : false;
  float m = -INFINITY, l = 0.f, o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
  const bool drop_on = a.drop.thr8 != 0u;
  const uint32_t rowh = drop_on ? drop_rowhash(a.drop.key, (uint64_t)(((int64_t)b * a.H + h) * a.Lq + qi)) : 0u;

  for (int bi = 0; bi < a.nblk; ++bi) {
    const AttnBlk& kb = a.blk[bi];
    float q[DH];
    if (active) load_row<T, DH>(reinterpret_cast<const T*>(kb.q) + ((int64_t)b * a.Lq + qi) * kb.ldq + h * DH, q);
    else {
#pragma unroll
      for (int d = 0; d < DH; ++d) q[d] = 0.f;
    }
    const T* kbase = reinterpret_cast<const T*>(kb.k) + (int64_t)b * kb.Lk * kb.ldk + h * DH;
    const T* vbase = reinterpret_cast<const T*>(kb.v) + (int64_t)b * kb.Lk * kb.ldv + h * DH;
    for (int k0 = 0; k0 < kb.Lk; k0 += kKT) {
      __syncthreads();
      stage_tile<T, DH, kKT, kQThreads>(Ks, kbase, kb.ldk, k0, kb.Lk, tid);
      stage_tile<T, DH, kKT, kQThreads>(Vs, vbase, kb.ldv, k0, kb.Lk, tid);
      if (tid < kKT) Mk[tid] = (k0 + tid < kb.Lk) ? kb.mask_k[(int64_t)b * kb.Lk + k0 + tid] : 0;
      __syncthreads();
      const int nk = min(kKT, kb.Lk - k0);
      uint32_t kw = 0xffffffffu;
      for (int j0 = 0; j0 < nk; j0 += 8) {
        float s[8];
        float mx = -INFINITY;
        if (drop_on && (j0 & 31) == 0) kw = drop_keep_word(rowh, attn_group(bi, k0 + j0), a.drop.thr8);   // k0 is a multiple of 64
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = j0 + jj;
          if (j < nk) {
            const float dt = dot_smem<DH>(q, &Ks[j][0]);
            float raw = (mq && Mk[j]) ? dt : kMaskFill;
            if (drop_on) raw = ((kw >> (j & 31)) & 1u) ? raw * a.drop.scale : 0.f;    // a dropped logit is 0, masked or not
            s[jj] = raw * a.scale;
          } else {
            s[jj] = -INFINITY;
          }
          mx = fmaxf(mx, s[jj]);
        }
        if (mx > m) {
          const float alpha = expf(m - mx);  // m = -inf -> 0
          l *= alpha;
#pragma unroll
          for (int d = 0; d < DH; ++d) o[d] *= alpha;
          m = mx;
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = j0 + jj;
          if (j < nk) {
            const float p = expf(s[jj] - m);
            l += p;
#pragma unroll
            for (int d = 0; d < DH; d += 4) {
              const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d]);
              o[d] = fmaf(p, vv.x, o[d]); o[d + 1] = fmaf(p, vv.y, o[d + 1]);
              o[d + 2] = fmaf(p, vv.z, o[d + 2]); o[d + 3] = fmaf(p, vv.w, o[d + 3]);
            }
          }
        }
      }
    }
  }
  if (active) {
    const float inv = 1.0f / l;
    store_row<T, DH>(reinterpret_cast<T*>(a.out) + ((int64_t)b * a.Lq + qi) * a.ldo + h * DH, o, inv);
    if (a.lse) a.lse[((int64_t)b * a.H + h) * a.Lq + qi] = m + logf(l);
  }
}

// ------------------------------------------------------------------------------ backward: dQ (+ delta)
template <typename T, int DH>
__global__ void __launch_bounds__(kQThreads) attn_bwd_dq_simt_kernel(AttnParams a) {
  __shared__ __align__(16) float Ks[kKT][DH];
  __shared__ __align__(16) float Vs[kKT][DH];
  __shared__ uint8_t Mk[kKT];
  const int b = blockIdx.z, h = blockIdx.y, tid = threadIdx.x;
  const int qi = blockIdx.x * kQThreads + tid;
  const bool active = qi < a.Lq;
  const bool mq = active ? (a.mask_q[(int64_t)b * a.Lq + qi] != 0) : false;
  float dO[DH];
  float delta = 0.f, lse = 0.f;
  if (active) {
    load_row<T, DH>(reinterpret_cast<const T*>(a.dout) + ((int64_t)b * a.Lq + qi) * a.lddo + h * DH, dO);
    float O[DH];
    load_row<T, DH>(reinterpret_cast<const T*>(a.out) + ((int64_t)b * a.Lq + qi) * a.ldo + h * DH, O);
#pragma unroll
    for (int d = 0; d < DH; ++d) delta = fmaf(dO[d], O[d], delta);
    lse = a.lse[((int64_t)b * a.H + h) * a.Lq + qi];
    a.delta[((int64_t)b * a.H + h) * a.Lq + qi] = delta;
  } else {
#pragma unroll
    for (int d = 0; d < DH; ++d) dO[d] = 0.f;
  }
  const bool drop_on = a.drop.thr8 != 0u;
  const uint32_t rowh = drop_on ? drop_rowhash(a.drop.key, (uint64_t)(((int64_t)b * a.H + h) * a.Lq + qi)) : 0u;
  const float gscale = drop_on ? a.scale * a.drop.scale : a.scale;      // d logit / d (q.k) for a surviving valid logit
  for (int bi = 0; bi < a.nblk; ++bi) {
    const AttnBlk& kb = a.blk[bi];
    float q[DH], dq[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { q[d] = 0.f; dq[d] = 0.f; }
    if (active) load_row<T, DH>(reinterpret_cast<const T*>(kb.q) + ((int64_t)b * a.Lq + qi) * kb.ldq + h * DH, q);
    const T* kbase = reinterpret_cast<const T*>(kb.k) + (int64_t)b * kb.Lk * kb.ldk + h * DH;
    const T* vbase = reinterpret_cast<const T*>(kb.v) + (int64_t)b * kb.Lk * kb.ldv + h * DH;
    for (int k0 = 0; k0 < kb.Lk; k0 += kKT) {
      __syncthreads();
      stage_tile<T, DH, kKT, kQThreads>(Ks, kbase, kb.ldk, k0, kb.Lk, tid);
      stage_tile<T, DH, kKT, kQThreads>(Vs, vbase, kb.ldv, k0, kb.Lk, tid);
      if (tid < kKT) Mk[tid] = (k0 + tid < kb.Lk) ? kb.mask_k[(int64_t)b * kb.Lk + k0 + tid] : 0;
      __syncthreads();
      const int nk = min(kKT, kb.Lk - k0);
      if (mq) {  // an overwritten (masked) or dropped logit passes no gradient to q/k
        uint32_t kw = 0xffffffffu;
        for (int j = 0; j < nk; ++j) {
          if (drop_on && (j & 31) == 0) kw = drop_keep_word(rowh, attn_group(bi, k0 + j), a.drop.thr8);
          if (!Mk[j] || !((kw >> (j & 31)) & 1u)) continue;
          const float s = dot_smem<DH>(q, &Ks[j][0]) * gscale;
          const float p = expf(s - lse);
          const float dp = dot_smem<DH>(dO, &Vs[j][0]);
          const float ds = p * (dp - delta) * gscale;
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            const float4 kv = *reinterpret_cast<const float4*>(&Ks[j][d]);
            dq[d] = fmaf(ds, kv.x, dq[d]); dq[d + 1] = fmaf(ds, kv.y, dq[d + 1]);
            dq[d + 2] = fmaf(ds, kv.z, dq[d + 2]); dq[d + 3] = fmaf(ds, kv.w, dq[d + 3]);
          }
        }
      }
    }
    if (active && kb.dq) store_row<T, DH>(reinterpret_cast<T*>(kb.dq) + ((int64_t)b * a.Lq + qi) * kb.lddq + h * DH, dq, 1.0f);
  }
}

// ------------------------------------------------------------------------------ backward: dK, dV of one key block
template <typename T, int DH>
__global__ void __launch_bounds__(kQThreads) attn_bwd_dkv_simt_kernel(AttnParams a, int which) {
  __shared__ __align__(16) float Qs[kQT][DH];
  __shared__ __align__(16) float dOs[kQT][DH];
  __shared__ float Ls[kQT], Ds[kQT];
  __shared__ uint8_t Mq[kQT];
  __shared__ uint32_t Rh[kQT];     // dropout row hashes of the tile's queries
  const bool drop_on = a.drop.thr8 != 0u;
  const AttnBlk& kb = a.blk[which];
  const int b = blockIdx.z, h = blockIdx.y, tid = threadIdx.x;
  const int kj = blockIdx.x * kQThreads + tid;
  const bool active = kj < kb.Lk;
  const bool mk = active ? (kb.mask_k[(int64_t)b * kb.Lk + kj] != 0) : false;
  float k[DH], v[DH], dk[DH], dv[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { k[d] = 0.f; v[d] = 0.f; dk[d] = 0.f; dv[d] = 0.f; }
  if (active) {
    load_row<T, DH>(reinterpret_cast<const T*>(kb.k) + ((int64_t)b * kb.Lk + kj) * kb.ldk + h * DH, k);
    load_row<T, DH>(reinterpret_cast<const T*>(kb.v) + ((int64_t)b * kb.Lk + kj) * kb.ldv + h * DH, v);
  }
  const T* qbase = reinterpret_cast<const T*>(kb.q) + (int64_t)b * a.Lq * kb.ldq + h * DH;
  const T* dobase = reinterpret_cast<const T*>(a.dout) + (int64_t)b * a.Lq * a.lddo + h * DH;
  const float* lse = a.lse + ((int64_t)b * a.H + h) * a.Lq;
  const float* del = a.delta + ((int64_t)b * a.H + h) * a.Lq;
  for (int q0 = 0; q0 < a.Lq; q0 += kQT) {
    __syncthreads();
    stage_tile<T, DH, kQT, kQThreads>(Qs, qbase, kb.ldq, q0, a.Lq, tid);
    stage_tile<T, DH, kQT, kQThreads>(dOs, dobase, a.lddo, q0, a.Lq, tid);
    if (tid < kQT) {
      const bool in = q0 + tid < a.Lq;
      Ls[tid] = in ? lse[q0 + tid] : 0.f;
      Ds[tid] = in ? del[q0 + tid] : 0.f;
      Mq[tid] = in ? a.mask_q[(int64_t)b * a.Lq + q0 + tid] : 0;
      Rh[tid] = drop_on ? drop_rowhash(a.drop.key, (uint64_t)(((int64_t)b * a.H + h) * a.Lq + q0 + tid)) : 0u;
    }
    __syncthreads();
    const int nq = min(kQT, a.Lq - q0);
    for (int i = 0; i < nq; ++i) {
      bool valid = mk && Mq[i];
      const float dt = dot_smem<DH>(k, &Qs[i][0]);
      float raw = valid ? dt : kMaskFill;
      float gscale = a.scale;
      if (drop_on) {                 // same keep bit the forward used for (query q0 + i, key kj of block `which`)
        const bool keep = (drop_keep_word(Rh[i], attn_group(which, kj), a.drop.thr8) >> (kj & 31)) & 1u;
        raw = keep ? raw * a.drop.scale : 0.f;
        valid = valid && keep;
        gscale *= a.drop.scale;
      }
      const float s = raw * a.scale;
      const float p = expf(s - Ls[i]);
      const float dp = dot_smem<DH>(v, &dOs[i][0]);
      const float ds = valid ? p * (dp - Ds[i]) * gscale : 0.f;
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        const float4 dov = *reinterpret_cast<const float4*>(&dOs[i][d]);
        const float4 qv = *reinterpret_cast<const float4*>(&Qs[i][d]);
        dv[d] = fmaf(p, dov.x, dv[d]); dv[d + 1] = fmaf(p, dov.y, dv[d + 1]);
        dv[d + 2] = fmaf(p, dov.z, dv[d + 2]); dv[d + 3] = fmaf(p, dov.w, dv[d + 3]);
        dk[d] = fmaf(ds, qv.x, dk[d]); dk[d + 1] = fmaf(ds, qv.y, dk[d + 1]);
        dk[d + 2] = fmaf(ds, qv.z, dk[d + 2]); dk[d + 3] = fmaf(ds, qv.w, dk[d + 3]);
      }
    }
  }
  if (active) {
    if (kb.dk) store_row<T, DH>(reinterpret_cast<T*>(kb.dk) + ((int64_t)b * kb.Lk + kj) * kb.lddk + h * DH, dk, 1.0f);
    if (kb.dv) store_row<T, DH>(reinterpret_cast<T*>(kb.dv) + ((int64_t)b * kb.Lk + kj) * kb.lddv + h * DH, dv, 1.0f);
  }
}

static int to_params(const mmi_attn_args* a, AttnParams& p, bool bwd) {
  MMI_CHECK_ARG(a != nullptr, "attn: null args");
  MMI_CHECK_ARG(a->nblk >= 1 && a->nblk <= 2, "attn: nblk must be 1 or 2");
  MMI_CHECK_ARG(a->B > 0 && a->H > 0 && a->Lq > 0, "attn: bad sizes");
  MMI_CHECK_ARG(a->mask_q && a->out && a->lse, "attn: null pointer");
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.nblk = a->nblk;
  p.mask_q = a->mask_q; p.out = a->out; p.ldo = a->ldo; p.lse = a->lse;
  p.dout = a->dout; p.lddo = a->lddo; p.delta = a->delta;
  p.scale = 1.0f / sqrtf((float)a->dh);
  p.drop = make_drop(a->drop);
  MMI_CHECK_ARG(p.drop.thr8 < 256u, "attn: dropout thr8 must be < 256");
  const int al = a->dtype == MMI_F32 ? 4 : 4;  // 4-element vector accesses
  for (int i = 0; i < a->nblk; ++i) {
    const mmi_attn_block& s = a->blk[i];
    MMI_CHECK_ARG(s.q && s.k && s.v && s.mask_k && s.Lk > 0, "attn: block %d has null pointer / Lk<=0", i);
    MMI_CHECK_ARG(s.ldq % al == 0 && s.ldk % al == 0 && s.ldv % al == 0, "attn: leading dims must be multiples of 4");
    p.blk[i] = AttnBlk{s.q, s.ldq, s.k, s.ldk, s.v, s.ldv, s.mask_k, s.Lk, s.dq, s.lddq, s.dk, s.lddk, s.dv, s.lddv};
  }
  if (bwd) MMI_CHECK_ARG(a->dout && a->delta, "attn bwd: null dout/delta");
  return MMI_OK;
}

template <typename T>
static int launch_attn(int kind, const AttnParams& p, int dh, int which, cudaStream_t st) {
  if (kind == 2) {
    dim3 grid((p.blk[which].Lk + kQThreads - 1) / kQThreads, p.H, p.B);
    if (dh == 32) attn_bwd_dkv_simt_kernel<T, 32><<<grid, kQThreads, 0, st>>>(p, which);
    else if (dh == 16) attn_bwd_dkv_simt_kernel<T, 16><<<grid, kQThreads, 0, st>>>(p, which);
    else { set_error("attn: head dim %d not supported (16 or 32)", dh); return MMI_ENOSUP; }
  } else {
    dim3 grid((p.Lq + kQThreads - 1) / kQThreads, p.H, p.B);
    if (kind == 0) {
      if (dh == 32) attn_fwd_simt_kernel<T, 32><<<grid, kQThreads, 0, st>>>(p);
      else if (dh == 16) attn_fwd_simt_kernel<T, 16><<<grid, kQThreads, 0, st>>>(p);
      else { set_error("attn: head dim %d not supported (16 or 32)", dh); return MMI_ENOSUP; }
    } else {
      if (dh == 32) attn_bwd_dq_simt_kernel<T, 32><<<grid, kQThreads, 0, st>>>(p);
      else if (dh == 16) attn_bwd_dq_simt_kernel<T, 16><<<grid, kQThreads, 0, st>>>(p);
      else { set_error("attn: head dim %d not supported (16 or 32)", dh); return MMI_ENOSUP; }
    }
  }
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}

int attn_simt(int kind, const mmi_attn_args* a, int which, cudaStream_t st) {
  AttnParams p;
  int rc = to_params(a, p, kind != 0);
  if (rc != MMI_OK) return rc;
  if (kind == 2) MMI_CHECK_ARG(which >= 0 && which < a->nblk, "attn dkv: bad block index %d", which);
  if (a->dtype == MMI_F32) return launch_attn<float>(kind, p, a->dh, which, st);
  if (a->dtype == MMI_BF16) return launch_attn<__nv_bfloat16>(kind, p, a->dh, which, st);
  set_error("attn: bad dtype %d", a->dtype);
  return MMI_EINVAL;
}

}  // namespace mmi
