// Microbenchmark: per-SM issue rates of the instructions the attention softmax loops are made of, to decide which
// formulation the softmax threads should use (scalar vs packed f32x2 arithmetic, f32 vs f16x2 / bf16x2 exponentials,
// 2- vs 3-input max, cvt packs).  Each kernel runs NW warps per SM, every thread ILP independent chains, and reports
// lane-operations per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ILP = 8;
constexpr int ITERS = 4096;

enum Op { EX2_F32, EX2_F16X2, EX2_BF16X2, FFMA, FFMA2, FADD, FADD2, FMNMX, FMNMX3, CVT_BF16X2, CVT_F16X2, MIX_FWD_SCALAR, MIX_FWD_PACKED, MIX_FWD_F16X2,
          TANH_F32, HFMA2_BF16, MIX_FWD_POLY, N_OPS };
const char* kNames[N_OPS] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2", "fma.rn.f32", "fma.rn.f32x2", "add.f32", "add.f32x2",
                             "max.f32 (2-in)", "max.f32 (3-in)", "cvt.rn.bf16x2.f32", "cvt.rn.f16x2.f32", "fwd mix scalar (ffma+ex2+fadd+.5cvt+fmnmx)",
                             "fwd mix packed (.5ffma2+ex2+.5fadd2+.5cvt+.5fmnmx3)", "fwd mix f16x2 exp (.5ffma2+.5cvt+.5ex2h2+.5hadd2+.5fmnmx3)",
                             "tanh.approx.f32", "fma.rn.bf16x2", "fwd mix poly/mufu 1:1"};
// lane-level useful results per instruction group, for the report
const double kOpsPerIter[N_OPS] = {1, 2, 2, 1, 2, 1, 2, 1, 2, 2, 2, 1, 2, 2, 1, 2, 2};

template <int OP>
__global__ void __launch_bounds__(1024) rate_kernel(float* out, long long* cycles, float seed) {
  float a[ILP], b[ILP];
  uint32_t h[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = seed + 0.001f * (threadIdx.x + i); b[i] = seed * 0.5f + i; h[i] = 0x3c003c00u + i + threadIdx.x; }
  const float c0 = seed * 0.999f, c1 = seed * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if constexpr (OP == EX2_F32) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if constexpr (OP == TANH_F32) {
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      } else if constexpr (OP == EX2_F16X2) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      } else if constexpr (OP == EX2_BF16X2) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      } else if constexpr (OP == HFMA2_BF16) {
        asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(__float_as_uint(c0)));
      } else if constexpr (OP == FFMA) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1));
      } else if constexpr (OP == FFMA2) {
        asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2}; mov.b64 z, {%3, %3}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}"
                     : "+f"(a[i]), "+f"(b[i]) : "f"(c0), "f"(c1));
      } else if constexpr (OP == FADD) {
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
      } else if constexpr (OP == FADD2) {
        asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2}; add.f32x2 x, x, y; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(b[i]) : "f"(c1));
      } else if constexpr (OP == FMNMX) {
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
      } else if constexpr (OP == FMNMX3) {
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c0));
      } else if constexpr (OP == CVT_BF16X2) {
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
        a[i] = __uint_as_float(h[i]);
      } else if constexpr (OP == CVT_F16X2) {
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
        a[i] = __uint_as_float(h[i]);
      } else if constexpr (OP == MIX_FWD_SCALAR) {
        // one score: x = s*scale - m ; e = ex2(x) ; l += e ; max ; half a pack
        float x, e;
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x) : "f"(a[i]), "f"(c0), "f"(c1));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(e));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e));
        if (i & 1) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(e), "f"(x)); }
      } else if constexpr (OP == MIX_FWD_PACKED) {
        // two scores per group
        float x0, x1, e0, e1;
        asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %4}; mov.b64 z, {%5, %5}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}"
                     : "=f"(x0), "=f"(x1) : "f"(a[i]), "f"(b[i]), "f"(c0), "f"(c1));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
        asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; add.f32x2 x, x, y; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(b[i]) : "f"(e0), "f"(e1));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(e0), "f"(e1));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(e0), "f"(e1));
      } else if constexpr (OP == MIX_FWD_F16X2) {
        float x0, x1;
        uint32_t xh, eh;
        asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %4}; mov.b64 z, {%5, %5}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}"
                     : "=f"(x0), "=f"(x1) : "f"(a[i]), "f"(b[i]), "f"(c0), "f"(c1));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(xh) : "f"(x1), "f"(x0));
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(eh) : "r"(xh));
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(eh));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(x0), "f"(x1));
      } else if constexpr (OP == MIX_FWD_POLY) {
        // two scores: one through MUFU, one through a degree-3 polynomial on the FMA pipe (Cody-Waite: 2^x = 2^floor * p(frac))
        float x0, x1, e0;
        asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %4}; mov.b64 z, {%5, %5}; fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}"
                     : "=f"(x0), "=f"(x1) : "f"(a[i]), "f"(b[i]), "f"(c0), "f"(c1));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
        float fl, fr, pl;
        asm volatile("add.rm.f32 %0, %1, 12582912.0;" : "=f"(fl) : "f"(x1));          // round-to-floor via magic constant: integer part in low mantissa bits
        float flf;
        asm volatile("sub.f32 %0, %1, 12582912.0;" : "=f"(flf) : "f"(fl));
        asm volatile("sub.f32 %0, %1, %2;" : "=f"(fr) : "f"(x1), "f"(flf));
        asm volatile("fma.rn.f32 %0, %1, 0.0555041, 0.2402265;" : "=f"(pl) : "f"(fr));
        asm volatile("fma.rn.f32 %0, %0, %1, 0.6931472;" : "+f"(pl) : "f"(fr));
        asm volatile("fma.rn.f32 %0, %0, %1, 1.0;" : "+f"(pl) : "f"(fr));
        uint32_t bits;
        asm volatile("shl.b32 %0, %1, 23;" : "=r"(bits) : "r"(__float_as_uint(fl)));
        asm volatile("add.s32 %0, %0, %1;" : "+r"(bits) : "r"(__float_as_uint(pl)));
        const float e1 = __uint_as_float(bits);
        asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; add.f32x2 x, x, y; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(b[i]) : "f"(e0), "f"(e1));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(e0), "f"(e1));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] + b[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, float* out, long long* cyc_d) {
  for (int nw : {4, 8, 16, 32}) {
    rate_kernel<OP><<<sms, nw * 32>>>(out, cyc_d, 0.5f);
    cudaDeviceSynchronize();
    rate_kernel<OP><<<sms, nw * 32>>>(out, cyc_d, 0.5f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-60s ERROR %s\n", kNames[OP], cudaGetErrorString(e)); return; }
    long long cyc[1024];
    cudaMemcpy(cyc, cyc_d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)cyc[i];
    avg /= sms;
    const double groups = (double)ITERS * ILP * nw * 32;
    printf("%-60s warps/SM %2d : %7.2f results/clk/SM  (%6.2f instr-groups/clk/SM)\n", kNames[OP], nw, groups * kOpsPerIter[OP] / avg, groups / avg);
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs\n", prop.name, sms);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  run<EX2_F32>(sms, out, cyc);
  run<EX2_F16X2>(sms, out, cyc);
  run<EX2_BF16X2>(sms, out, cyc);
  run<TANH_F32>(sms, out, cyc);
  run<FFMA>(sms, out, cyc);
  run<FFMA2>(sms, out, cyc);
  run<HFMA2_BF16>(sms, out, cyc);
  run<FADD>(sms, out, cyc);
  run<FADD2>(sms, out, cyc);
  run<FMNMX>(sms, out, cyc);
  run<FMNMX3>(sms, out, cyc);
  run<CVT_BF16X2>(sms, out, cyc);
  run<CVT_F16X2>(sms, out, cyc);
  run<MIX_FWD_SCALAR>(sms, out, cyc);
  run<MIX_FWD_PACKED>(sms, out, cyc);
  run<MIX_FWD_F16X2>(sms, out, cyc);
  run<MIX_FWD_POLY>(sms, out, cyc);
  return 0;
}
