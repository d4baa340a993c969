// Micro-benchmark (development tool): tcgen05.ld throughput with 1..8 warps reading concurrently
// (32x32b.x16 and .x4 shapes), to size the stem epilogue and the converters of the fused kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I enhance-cb-whisper_b200/csrc \
//        tools/tmem_probe.cu enhance-cb-whisper_b200/csrc/kws_abi.cu -o tools/tmem_probe
#include <vector>
#include "kws_common.cuh"
using namespace kws;

__device__ __forceinline__ void ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// mode 0: x16 loads, wait after each; 1: x16 loads, wait after 4; 2: x4 loads, wait after 12; 3: x32, wait after 2
__global__ void __launch_bounds__(256, 1) probe(int n_warps, int mode, int iters, long long* out) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < n_warps) {
    for (int i = 0; i < iters; ++i) {
      if (mode == 0) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) { tmem_ld16(t + ((i * 64 + j * 16) & 511), v); tmem_ld_wait(); acc += v[0] ^ v[15]; }
      } else if (mode == 1) {
        uint32_t v[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j) tmem_ld16(t + ((i * 64 + j * 16) & 511), v[j]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += v[j][0] ^ v[j][15];
      } else if (mode == 2) {
        uint32_t v[16][4];
#pragma unroll
        for (int j = 0; j < 16; ++j) ld4(t + ((i * 64 + j * 4) & 511), v[j]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += v[j][0] ^ v[j][3];
      } else {
        uint32_t v[2][32];
#pragma unroll
        for (int j = 0; j < 2; ++j) ld32(t + ((i * 64 + j * 32) & 511), v[j]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 2; ++j) acc += v[j][0] ^ v[j][31];
      }
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345u) out[1] = acc;
  if (threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_slot, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int iters = 2000;
  for (int mode = 0; mode < 4; ++mode)
    for (int nw : {1, 2, 4, 8}) {
      for (int rep = 0; rep < 2; ++rep) {
        probe<<<1, 256>>>(nw, mode, iters, d);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed\n"); return 1; }
      }
      long long c;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      // each iteration reads 64 columns x 32 lanes x 4 B = 8 KB per warp
      printf("mode %d warps %d: %.1f cyc per 64-col read per warp, %.1f B/clk total\n", mode, nw, (double)c / iters,
             8192.0 * nw * iters / (double)c);
    }
  return 0;
}
