// C-ABI glue: error reporting, argument validation and dispatch for mmi_gemm / mmi_attn_*.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace mmi {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int attn_simt(int kind, const mmi_attn_args* a, int which, cudaStream_t st);
int attn_tc(int kind, const mmi_attn_args* a, int which, cudaStream_t st);

}  // namespace mmi

using namespace mmi;

extern "C" int mmi_version(void) { return 100; }
extern "C" const char* mmi_last_error(void) { return g_err; }
extern "C" int mmi_has_tc(void) { return tc_available() ? 1 : 0; }

extern "C" int64_t mmi_gemm_split_workspace(int layout, int64_t M, int64_t N, int64_t K) { return gemm_split_workspace(layout, M, N, K); }

extern "C" int mmi_gemm(const mmi_gemm_args* a, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(a != nullptr, "gemm: null args");
  MMI_CHECK_ARG(a->A && a->B && a->C, "gemm: null operand");
  MMI_CHECK_ARG(a->M >= 0 && a->N > 0 && a->K > 0, "gemm: bad sizes M=%lld N=%lld K=%lld", (long long)a->M, (long long)a->N, (long long)a->K);
  MMI_CHECK_ARG(a->layout >= MMI_GEMM_NT && a->layout <= MMI_GEMM_TN, "gemm: bad layout %d", a->layout);
  MMI_CHECK_ARG(a->N % 4 == 0 && a->ldc % 4 == 0, "gemm: N and ldc must be multiples of 4");
  MMI_CHECK_ARG(!(a->accumulate && a->out_dtype != MMI_F32), "gemm: accumulate needs an fp32 C");
  MMI_CHECK_ARG(a->split_k >= 0 && (a->split_k == 1 || a->accumulate), "gemm: split_k != 1 needs accumulate=1 (0 = auto)");
  MMI_CHECK_ARG(!(a->add && a->add_mod <= 0), "gemm: add needs add_mod > 0");
  if (a->M == 0) return MMI_OK;
  GemmParams p;
  p.layout = a->layout; p.impl = a->impl; p.in_dtype = a->in_dtype; p.out_dtype = a->out_dtype;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.A = a->A; p.lda = a->lda; p.B = a->B; p.ldb = a->ldb; p.C = a->C; p.ldc = a->ldc;
  p.bias = a->bias; p.act = a->act; p.preact = a->preact; p.ld_preact = a->ld_preact;
  p.mul_gelu_grad = a->mul_gelu_grad; p.ld_mul = a->ld_mul;
  p.add = a->add; p.ld_add = a->ld_add; p.add_mod = a->add_mod; p.add_dtype = a->add_dtype;
  p.accumulate = a->accumulate; p.split_k = a->split_k;
  p.save_act_grad = a->save_act_grad; p.mul_is_grad = a->mul_is_grad;
  p.drop = make_drop(a->drop);
  p.mul_scale = a->mul_scale;
  p.split_ws = a->split_ws; p.split_ws_bytes = a->split_ws_bytes;
  MMI_CHECK_ARG(a->act >= MMI_ACT_NONE && a->act <= MMI_ACT_RELU, "gemm: bad activation %d", a->act);
  MMI_CHECK_ARG(!(a->act == MMI_ACT_RELU && a->preact), "gemm: ReLU saves nothing (its backward reads the layer's output)");
  MMI_CHECK_ARG(p.drop.thr8 < 256u, "gemm: dropout thr8 must be < 256");
  MMI_CHECK_ARG(!(p.drop.thr8 && (a->accumulate || a->split_k > 1)), "gemm: dropout cannot be combined with accumulate / split-K");
  MMI_CHECK_ARG(!(p.drop.thr8 && a->mul_gelu_grad && a->act != MMI_ACT_NONE), "gemm: dropout with both act and mul_gelu_grad is undefined");
  if (a->impl != MMI_IMPL_TC && p.split_k == 0) p.split_k = 1;
  if (a->impl == MMI_IMPL_TC) {
    if (!tc_available()) { set_error("gemm: tcgen05 path requested but not available on this device/build"); return MMI_ENOSUP; }
    return gemm_tc(p, st);
  }
  // SIMT path: 4-element vector loads along the contiguous dimension of each operand
  if (a->layout == MMI_GEMM_NT) MMI_CHECK_ARG(a->K % 4 == 0 && a->lda % 4 == 0 && a->ldb % 4 == 0, "gemm NT: K, lda, ldb must be multiples of 4");
  if (a->layout == MMI_GEMM_NN) MMI_CHECK_ARG(a->K % 4 == 0 && a->lda % 4 == 0 && a->ldb % 4 == 0, "gemm NN: K, lda, ldb must be multiples of 4");
  if (a->layout == MMI_GEMM_TN) MMI_CHECK_ARG(a->M % 4 == 0 && a->lda % 4 == 0 && a->ldb % 4 == 0, "gemm TN: M, lda, ldb must be multiples of 4");
  return gemm_simt(p, st);
}

static int attn_dispatch(int kind, const mmi_attn_args* a, int which, mmi_stream_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MMI_CHECK_ARG(a != nullptr, "attn: null args");
  if (a->impl == MMI_IMPL_TC) return attn_tc(kind, a, which, st);
  for (int i = 0; i < a->nblk && i < 2; ++i)
    MMI_CHECK_ARG(kind == 0 || (!a->blk[i].dbq && !a->blk[i].dbk && !a->blk[i].dbv), "attn: fused bias-gradient sums (dbq/dbk/dbv) need MMI_IMPL_TC");
  return attn_simt(kind, a, which, st);
}

extern "C" int mmi_attn_fwd(const mmi_attn_args* a, mmi_stream_t stream) { return attn_dispatch(0, a, 0, stream); }
extern "C" int mmi_attn_bwd_dq(const mmi_attn_args* a, mmi_stream_t stream) { return attn_dispatch(1, a, 0, stream); }
extern "C" int mmi_attn_bwd_dkv(const mmi_attn_args* a, int which, mmi_stream_t stream) { return attn_dispatch(2, a, which, stream); }
extern "C" int mmi_attn_bwd_all(const mmi_attn_args* a, mmi_stream_t stream) {
  MMI_CHECK_ARG(a != nullptr, "attn: null args");
  if (a->impl != MMI_IMPL_TC) return 1;          // the strict-parity FFMA path keeps its two kernels
  return attn_tc(4, a, 0, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mmi_attn_bwd_fused(const mmi_attn_args* a, int which, mmi_stream_t stream) {
  MMI_CHECK_ARG(a != nullptr, "attn: null args");
  if (a->impl != MMI_IMPL_TC) return 1;          // the strict-parity FFMA path keeps its two kernels
  return attn_tc(3, a, which, reinterpret_cast<cudaStream_t>(stream));
}
