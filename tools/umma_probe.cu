// Micro-benchmark (development tool, not part of the library): issue rate of tcgen05.mma for the
// operand shapes/layouts the stem and similarity kernels use, with both operands in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I enhance-cb-whisper_b200/csrc \
//        tools/umma_probe.cu enhance-cb-whisper_b200/csrc/kws_abi.cu -o tools/umma_probe
// Prints cycles per MMA (one CTA per SM, all SMs busy) for a list of (M, N, layout) cases.
#include <vector>

#include "kws_common.cuh"

using namespace kws;

struct Case {
  int N;          // MMA N
  int swizzle;    // 0: no-swizzle K-major (stem layout), 1: 128B swizzle
  int a_span;     // distinct A start offsets cycled through (bytes step 16) -> models tap shifts
  int n_mma;      // MMAs per timed batch
  int batches;
  int ldtm;       // epilogue warps concurrently reading TMEM (0/1)
};

__global__ void __launch_bounds__(192, 1) probe_kernel(Case c, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: A region 64 KB, B region 64 KB (contents irrelevant for timing; zero them)
  for (int i = threadIdx.x; i < (128 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    stop = 0;
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_f16(128, c.N, 0);
    const uint32_t sa = smem_u32(base), sb = smem_u32(base + 64 * 1024);
    uint64_t adesc, bdesc;
    if (c.swizzle) {
      adesc = make_smem_desc(sa, 16, 1024, LAYOUT_SW128);
      bdesc = make_smem_desc(sb, 16, 1024, LAYOUT_SW128);
    } else {
      adesc = make_smem_desc(sa, 2112, 128, LAYOUT_NONE);
      bdesc = make_smem_desc(sb, c.N * 16, 128, LAYOUT_NONE);
    }
    long long total = 0;
    uint32_t ph = 0;
    for (int b = 0; b < c.batches; ++b) {
      const long long t0 = clock64();
      for (int i = 0; i < c.n_mma; ++i) {
        const uint64_t aoff = (uint64_t)(c.swizzle ? (i & 3) * 2 : (i % c.a_span));
        const uint64_t boff = (uint64_t)(c.swizzle ? (i & 3) * 2 : (i % 7) * ((c.N * 32) >> 4));
        umma_f16(tmem + (i & 1) * 256, adesc + aoff, bdesc + boff, idesc, i > 1);
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph, 1);
      ph ^= 1;
      total += clock64() - t0;
    }
    out_cycles[blockIdx.x] = total;
    stop = 1;
  } else if (warp >= 2 && c.ldtm) {
    // four warps hammer TMEM reads (as an epilogue would) while the MMAs run
    const int q = warp & 3;
    uint32_t v[16];
    uint32_t acc = 0;
    while (!*(volatile int*)&stop) {
#pragma unroll 1
      for (int ch = 0; ch < 16; ++ch) {
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + ch * 16, v);
        tmem_ld_wait();
        acc += v[0] + v[15];
      }
    }
    if (acc == 0x12345678u) out_cycles[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = 1024 + 128 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * sms);
  std::vector<long long> h(sms);
  const Case cases[] = {
      {64, 0, 5, 490, 20, 0},  {64, 0, 5, 490, 20, 1},  {64, 1, 1, 490, 20, 0},  {128, 0, 5, 490, 20, 0},
      {128, 1, 1, 490, 20, 0}, {256, 1, 1, 490, 20, 0}, {160, 1, 1, 490, 20, 0}, {160, 1, 1, 490, 20, 1},
      {32, 1, 1, 490, 20, 0},  {16, 1, 1, 490, 20, 0},
  };
  for (const Case& c : cases) {
    for (int rep = 0; rep < 2; ++rep) {
      probe_kernel<<<sms, 192, smem>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("launch failed: %s\n", cudaGetErrorString(e));
        return 1;
      }
    }
    cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mx = 0, sum = 0;
    for (int i = 0; i < sms; ++i) {
      sum += (double)h[i];
      if ((double)h[i] > mx) mx = (double)h[i];
    }
    const double per = sum / sms / ((double)c.n_mma * c.batches);
    const double floor_cyc = 128.0 * c.N / 256.0;
    printf("M=128 N=%3d %s a_span=%d ldtm=%d : %.1f cyc/MMA (mean over SMs; max SM %.1f), floor %.0f, smem bytes/MMA %d -> %.0f B/clk\n",
           c.N, c.swizzle ? "SW128 " : "NOSWZ ", c.a_span, c.ldtm, per, mx / ((double)c.n_mma * c.batches), floor_cyc,
           (128 + c.N) * 32, (128 + c.N) * 32 / per);
  }
  return 0;
}
