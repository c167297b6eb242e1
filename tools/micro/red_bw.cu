// Microbenchmark: how fast can partial dQ tiles (fp32) be ACCUMULATED into global memory from many CTAs?
// Decides how a key-tile-owner attention backward hands dQ over:
//   mode 0: red.global.add.f32          (scalar, one lane = one float)
//   mode 1: red.global.add.v4.f32       (16 B per lane, sm_90+)
//   mode 2: cp.reduce.async.bulk.global.shared::cta.add.f32  (bulk reduce of an 8 KB smem tile, one thread issues)
// Every CTA adds 8 KB tiles ([64 rows x 32 floats], row pitch `pitch` floats) at pseudo-random row offsets of a large
// buffer, `conc` CTAs hitting each tile (the number of key tiles that share a query tile).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bw red_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(128) k(float* buf, long long n_tiles, int pitch, int iters, int conc) {
  __shared__ __align__(128) float tile[64 * 32];
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) tile[i] = 1.0f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int it = 0; it < iters; ++it) {
    // `conc` consecutive CTAs share a tile
    unsigned long long h = ((unsigned long long)(blockIdx.x / conc) * 0x9E3779B97F4A7C15ull + (unsigned long long)it * 0xC2B2AE3D27D4EB4Full);
    const long long t = (long long)((h >> 20) % (unsigned long long)n_tiles);
    float* dst = buf + t * 64 * (long long)pitch;
    if (MODE == 0) {
      for (int r = warp; r < 64; r += 4) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (long long)r * pitch + lane), "f"(tile[r * 32 + lane]) : "memory");
    } else if (MODE == 1) {
      // 8 lanes cover one 128 B row; a warp covers 4 rows per instruction
      for (int r = warp * 4 + (lane >> 3); r < 64; r += 16) {
        const float4 v = *reinterpret_cast<const float4*>(&tile[r * 32 + (lane & 7) * 4]);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + (long long)r * pitch + (lane & 7) * 4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
    } else {
      // one bulk reduce per row (128 B) would be 64 instructions; rows are contiguous only when pitch == 32
      if (threadIdx.x == 0) {
        if (pitch == 32) {
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(tile)), "r"(64 * 32 * 4) : "memory");
        } else {
          for (int r = 0; r < 64; ++r)
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst + (long long)r * pitch), "r"(smem_u32(tile + r * 32)), "r"(128) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
  }
  if (MODE == 2 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
void run(const char* name, float* buf, long long n_tiles, int pitch, int conc) {
  const int iters = 256, grid = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 128>>>(buf, n_tiles, pitch, 8, conc);
  cudaEventRecord(e0);
  k<MODE><<<grid, 128>>>(buf, n_tiles, pitch, iters, conc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)grid * iters * 8192.0;
  printf("%-34s pitch=%4d conc=%d : %8.3f ms  %8.1f GB/s of partial tiles  (%s)\n", name, pitch, conc, ms, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const long long bytes = 1ll << 30;   // 1 GiB accumulation buffer (larger than L2)
  float* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  for (int pitch : {32, 512}) {
    const long long n_tiles = bytes / 4 / 64 / pitch;
    for (int conc : {1, 4}) {
      run<0>("red.global.add.f32", buf, n_tiles, pitch, conc);
      run<1>("red.global.add.v4.f32", buf, n_tiles, pitch, conc);
      run<2>("cp.reduce.async.bulk .add.f32", buf, n_tiles, pitch, conc);
    }
  }
  // L2-resident target (64 MiB)
  {
    const long long n_tiles = (64ll << 20) / 4 / 64 / 32;
    run<1>("red.v4 (64 MiB target, L2)", buf, n_tiles, 32, 4);
    run<2>("bulk reduce (64 MiB target, L2)", buf, n_tiles, 32, 4);
  }
  cudaFree(buf);
  return 0;
}
