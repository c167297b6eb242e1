// placeholder: replaced by the tcgen05 GEMM
#include "common.cuh"
#include "gemm_epilogue.cuh"
namespace mmi {
bool tc_available() { return false; }
int gemm_tc(const GemmParams&, cudaStream_t) { set_error("gemm_tc: not built"); return MMI_ENOSUP; }
int attn_tc(int, const mmi_attn_args*, int, cudaStream_t) { set_error("attn_tc: not built"); return MMI_ENOSUP; }
}
