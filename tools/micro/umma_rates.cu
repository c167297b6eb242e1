// Microbenchmark: cycles per tcgen05.mma for the operand shapes / layouts of the attention backward (one CTA, one issuing
// thread, R back-to-back MMAs into one accumulator, then commit + wait).  Answers: which of the five products of a
// (128 keys x 64 queries) tile is expensive -- the tensor pipe (floor 16 / 32 cycles at N = 32 / 64) or the shared-memory
// operand fetch -- and what an A operand read from TMEM (".ts" form) would cost instead.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I segmminterest_b200/csrc tools/micro/umma_rates.cu -o tools/micro/umma_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace mmi::tc;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t d_k64(uint32_t a, int k) { return make_smem_desc(a + k * 32, 16, 512, 4); }
__device__ __forceinline__ uint64_t d_k128(uint32_t a, int k) { return make_smem_desc(a + k * 32, 16, 1024, 2); }
__device__ __forceinline__ uint64_t d_mn64(uint32_t a, int k) { return make_smem_desc(a + k * 1024, 512, 512, 4); }
__device__ __forceinline__ uint64_t d_mn128(uint32_t a, int k) { return make_smem_desc(a + k * 2048, 8192, 1024, 2); }

constexpr int NCFG = 12;
__global__ void __launch_bounds__(128, 1) k(long long* out, int R, int noise_warps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  const uint32_t a0 = smem_u32(smem), b0 = a0 + 48 * 1024;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 0) {
    uint32_t ph = 0;
    const uint32_t tmu = uniform(tm), a0u = uniform(a0), b0u = uniform(b0);
    for (int cfg = 0; cfg < NCFG; ++cfg) {
      long long t0 = clock64(), t1 = 0, t2 = 0;
      // warp-uniform loop, one elected lane issues: the operands stay in uniform registers (a lane-0-only loop makes
      // ptxas wrap every UTCHMMA in a waterfall loop and measures that instead)
#define LOOP(EXPR) for (int r = 0; r < R; r += 8) { if (elect_one()) { _Pragma("unroll") for (int q = 0; q < 8; ++q) { const uint32_t acc = (r + q) > 0; EXPR; } } __syncwarp(); }
      switch (cfg) {
        case 0: LOOP(umma_f16(tmu, d_k64(a0u, q & 1), d_k64(b0u, q & 1), make_idesc(128, 64, false, false), acc)); break;
        case 1: LOOP(umma_f16(tmu, d_k128(a0u, q & 3), d_mn64(b0u, q & 3), make_idesc(128, 32, false, true), acc)); break;
        case 2: LOOP(umma_f16(tmu, d_mn128(a0u, q & 7), d_mn64(b0u, q & 7), make_idesc(64, 32, true, true), acc)); break;
        case 3: LOOP(umma_ts(tmu, tmu + 256 + (q & 3) * 8, d_mn64(b0u, q & 3), make_idesc(128, 32, false, true), acc)); break;
        case 4: LOOP(umma_f16(tmu, make_smem_desc(a0u + (q & 3) * 32, 16, 1024), make_smem_desc(b0u + (q & 3) * 32, 16, 1024), make_idesc(128, 256, false, false), acc)); break;
        case 5: LOOP(umma_f16(tmu, d_k64(a0u, q & 1), d_k64(b0u, q & 1), make_idesc(128, 32, false, false), acc)); break;
        case 6: LOOP(umma_f16(tmu, d_mn128(a0u, q & 7), d_mn64(b0u, q & 7), make_idesc(128, 32, true, true), acc)); break;
        case 7: LOOP(umma_f16(tmu, d_k128(a0u, q & 3), d_k64(b0u, q & 1), make_idesc(128, 32, false, false), acc)); break;
        case 8: LOOP(umma_ts(tmu, tmu + 256 + (q & 3) * 8, d_k64(b0u, q & 1), make_idesc(128, 32, false, false), acc)); break;
        case 9: LOOP(umma_f16(tmu, d_k64(a0u, q & 1), d_k64(b0u, q & 1), make_idesc(128, 128, false, false), acc)); break;
        case 10: LOOP(umma_f16(tmu, d_k64(a0u, q & 1), d_k64(b0u, q & 1), make_idesc(64, 64, false, false), acc)); break;
        case 11: LOOP(umma_ts(tmu, tmu + 256 + (q & 3) * 8, d_mn64(b0u, q & 3), make_idesc(128, 64, false, true), acc)); break;
      }
      t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, ph);
      ph ^= 1;
      tcgen05_fence_after();
      t2 = clock64();
      if (lane == 0) { out[cfg * 2] = t1 - t0; out[cfg * 2 + 1] = t2 - t0; }
      __syncwarp();
    }
    if (lane == 0) stop = 1;
  } else if (warp - 1 < noise_warps) {
    // shared-memory noise: 16-byte stores + loads like the softmax warps' staging traffic
    uint32_t x = 0;
    uint32_t addr = a0 + 80 * 1024 + (threadIdx.x - 32) * 16;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(addr + ((i * 1536) & 8191)), "r"(x) : "memory");
      }
      x++;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
  }
}

int main() {
  const char* names[NCFG] = {"S^T/dP^T  M128 N64  A,B K-major SW64 (SS)", "dV/dK     M128 N32  A K-major SW128, B MN-major SW64 (SS)",
                             "dQ        M64  N32  A MN-major SW128, B MN-major SW64 (SS)", "dV/dK     M128 N32  A in TMEM, B MN-major SW64 (TS)",
                             "GEMM      M128 N256 A,B K-major SW128 (SS)", "          M128 N32  A,B K-major SW64 (SS)",
                             "          M128 N32  A MN-major SW128, B MN-major SW64 (SS)", "          M128 N32  A K-major SW128, B K-major SW64 (SS)",
                             "          M128 N32  A in TMEM, B K-major SW64 (TS)", "          M128 N128 A,B K-major SW64 (SS)",
                             "          M64  N64  A,B K-major SW64 (SS)", "          M128 N64  A in TMEM, B MN-major SW64 (TS)"};
  long long* d;
  cudaMalloc(&d, NCFG * 2 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int noise = 0; noise <= 3; noise += 3) {
    const int R = 256;
    k<<<1, 128, 100 * 1024>>>(d, R, noise);
    k<<<1, 128, 100 * 1024>>>(d, R, noise);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[NCFG * 2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("# %d back-to-back tcgen05.mma (K = 16 each), one CTA, %d warps storing to shared memory meanwhile\n", R, noise);
    printf("%-66s %10s %12s\n", "product", "issue clk", "complete clk  (per MMA)");
    for (int c = 0; c < NCFG; ++c) printf("%-66s %10.1f %12.1f\n", names[c], (double)h[2 * c] / R, (double)h[2 * c + 1] / R);
  }
  return 0;
}
