// a-15 / 8f-4: validation metrics of models/my_evaluation.py on the device, so that the driver's `valid_model`
// (main_for_seq_leave_earlystop_SegMM.py:132-186,396-432) needs ONE small D2H copy per batch instead of a Python
// loop with an `.item()` sync per row and an sklearn call on host copies.
//
//   interests = sigmoid(logits) * exposure_prob                       main...SegMM.py:402-403
//   survival  = exp(cumsum(log interests))                            my_evaluation.py:273-274 (test_type 'new')
//             = interests themselves                                  my_evaluation.py:270-271 (test_type 'old', input_kind 2)
//   ProbAUC   = roc_auc_score(label, survival) over positions with gt != -2, label = (gt == -1 ? 0 : gt)   :73-80
//   per row   : LeaveMSE prediction = sum of survival over valid positions   :82-85
//               view_length = #(gt == 1), duration = #(gt != -2)
//               LeaveCTR = 1 - interest[view-1], LeaveCTR_view = 1 - survival[view-1]  (python index: -1 wraps)  :87-90
//               JaccardSim (length_aware) = (sum_{t<view} (1 - |gt_t - survival_t|) + (duration - view)) / duration  :37-57
//
// The AUC is the exact Mann-Whitney statistic (ties count 1/2), accumulated in integers: the same value
// sklearn's trapezoidal roc_auc_score returns, independent of summation order.
#include "common.cuh"

namespace mmi {

constexpr int kRowCols = 6;   // per-row output columns: pred_view, view_length, duration, LeaveCTR, LeaveCTR_view, JaccardSim

// one warp per row (L <= 64); writes scores / labels for the AUC pass and the per-row metrics
__global__ void __launch_bounds__(256) eval_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ gt, int B, int L,
                                                        const float* __restrict__ ep, int input_kind, float* __restrict__ score, int8_t* __restrict__ label,
                                                        float* __restrict__ rows, unsigned long long* __restrict__ counters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (blockIdx.x == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0ull;   // the pair kernel runs after this one on the stream
  if (row >= B) return;
  float itr[2], logp[2];
  long long g[2];
  bool in[2], valid[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int l = lane + 32 * h;
    in[h] = l < L;
    const float x = in[h] ? logits[(size_t)row * L + l] : 0.f;
    g[h] = in[h] ? gt[(size_t)row * L + l] : -2;
    valid[h] = in[h] && g[h] != -2;
    itr[h] = input_kind == 0 ? (1.0f / (1.0f + expf(-x))) * (in[h] ? ep[l] : 1.f) : (in[h] ? x : 1.f);   // 1, 2: interests given
    logp[h] = in[h] ? logf(itr[h]) : 0.f;
  }
  float sc0 = logp[0], sc1 = logp[1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float a = __shfl_up_sync(0xffffffffu, sc0, o), b = __shfl_up_sync(0xffffffffu, sc1, o);
    if (lane >= o) { sc0 += a; sc1 += b; }
  }
  sc1 += __shfl_sync(0xffffffffu, sc0, 31);
  const float surv[2] = {input_kind == 2 ? itr[0] : expf(sc0), input_kind == 2 ? itr[1] : expf(sc1)};   // 2: test_type 'old'
  const int n_view = __popc(__ballot_sync(0xffffffffu, in[0] && g[0] == 1)) + __popc(__ballot_sync(0xffffffffu, in[1] && g[1] == 1));
  const int n_valid = __popc(__ballot_sync(0xffffffffu, valid[0])) + __popc(__ballot_sync(0xffffffffu, valid[1]));
  float pred = 0.f, jac = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int l = lane + 32 * h;
    if (in[h]) {
      score[(size_t)row * L + l] = surv[h];
      label[(size_t)row * L + l] = valid[h] ? (g[h] == -1 ? 0 : (int8_t)(g[h] != 0)) : -1;
      if (valid[h]) pred += surv[h];
      if (l < n_view) jac += 1.f - fabsf((float)g[h] - surv[h]);
    }
  }
  pred = warp_sum(pred);
  jac = warp_sum(jac);
  const int pos = (n_view - 1 + L) % L;                      // interest[view_length - 1]: index -1 is the last position
  const float i_at = pos < 32 ? __shfl_sync(0xffffffffu, itr[0], pos) : __shfl_sync(0xffffffffu, itr[1], pos - 32);
  const float s_at = pos < 32 ? __shfl_sync(0xffffffffu, surv[0], pos) : __shfl_sync(0xffffffffu, surv[1], pos - 32);
  if (lane == 0) {
    float* r = rows + (size_t)row * kRowCols;
    r[0] = pred; r[1] = (float)n_view; r[2] = (float)n_valid;
    r[3] = 1.f - i_at; r[4] = 1.f - s_at;
    r[5] = (jac + (float)(n_valid - n_view)) / (float)n_valid;
  }
}

// counters: 0 = 2 * #(neg < pos) + #(neg == pos) over all (pos, neg) pairs, 1 = n_pos, 2 = n_neg
__global__ void __launch_bounds__(256) auc_pairs_kernel(const float* __restrict__ score, const int8_t* __restrict__ label, int n,
                                                        unsigned long long* __restrict__ counters) {
  __shared__ float s_sc[1024];
  __shared__ int8_t s_lb[1024];
  __shared__ unsigned long long red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool is_pos = i < n && label[i] == 1;
  const bool is_neg = i < n && label[i] == 0;
  const float si = i < n ? score[i] : 0.f;
  unsigned int cnt = 0;                                       // <= 2 * n per thread: n < 2^30 checked by the host
  const bool block_has_pos = __syncthreads_or(is_pos);
  if (block_has_pos) {
    for (int j0 = 0; j0 < n; j0 += 1024) {
      __syncthreads();
      for (int t = threadIdx.x; t < 1024; t += blockDim.x) {
        const int j = j0 + t;
        s_sc[t] = j < n ? score[j] : 0.f;
        s_lb[t] = j < n ? label[j] : (int8_t)-1;
      }
      __syncthreads();
      if (is_pos) {
#pragma unroll 8
        for (int t = 0; t < 1024; ++t) {
          const unsigned int neg = s_lb[t] == 0;
          cnt += neg * (2u * (s_sc[t] < si) + (s_sc[t] == si));
        }
      }
    }
  }
  unsigned long long c = cnt;
  unsigned long long np = __popc(__ballot_sync(0xffffffffu, is_pos)), nn = __popc(__ballot_sync(0xffffffffu, is_neg));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    if (tot) atomicAdd(&counters[0], tot);
  }
  if (lane == 0) {
    if (np) atomicAdd(&counters[1], np);
    if (nn) atomicAdd(&counters[2], nn);
  }
}

__global__ void auc_final_kernel(const unsigned long long* __restrict__ counters, float* __restrict__ out) {
  const double np = (double)counters[1], nn = (double)counters[2];
  out[0] = (float)(0.5 * (double)counters[0] / (np * nn));   // one class missing: 0/0 = NaN (sklearn raises ValueError)
  out[1] = (float)np;
  out[2] = (float)nn;
  out[3] = 0.f;
}

}  // namespace mmi

using namespace mmi;

// workspace layout (bytes): counters 4 x u64 | scores B*L f32 | labels B*L i8
extern "C" int64_t mmi_eval_metrics_workspace(int B, int L) {
  const int64_t n = (int64_t)B * L;
  return 32 + n * 4 + ((n + 15) / 16) * 16;
}

extern "C" int mmi_eval_metrics(const float* logits, const int64_t* gt, int B, int L, const float* exposure_prob, int input_kind,
                                void* workspace, float* rows, float* out, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(logits && gt && workspace && rows && out, "eval_metrics: null pointer");
  MMI_CHECK_ARG(input_kind == 1 || input_kind == 2 || (input_kind == 0 && exposure_prob),
                "eval_metrics: input_kind 0 (logits, needs exposure_prob), 1 (interests) or 2 (interests that already are survival probabilities)");
  MMI_CHECK_ARG(B > 0 && L > 0 && L <= 64, "eval_metrics: need B > 0 and 0 < L <= 64 (got B %d, L %d)", B, L);
  MMI_CHECK_ARG((int64_t)B * L < (1ll << 30), "eval_metrics: B * L must stay below 2^30");
  MMI_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "eval_metrics: workspace must be 8-byte aligned");
  const int n = B * L;
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(workspace);
  float* score = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 32);
  int8_t* label = reinterpret_cast<int8_t*>(score + n);
  eval_rows_kernel<<<(B + 7) / 8, 256, 0, st>>>(logits, gt, B, L, exposure_prob, input_kind, score, label, rows, counters);
  MMI_CHECK_LAUNCH();
  auc_pairs_kernel<<<(n + 255) / 256, 256, 0, st>>>(score, label, n, counters);
  MMI_CHECK_LAUNCH();
  auc_final_kernel<<<1, 1, 0, st>>>(counters, out);
  MMI_CHECK_LAUNCH();
  return MMI_OK;
}
