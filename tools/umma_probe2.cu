// Micro-benchmark v2 (development tool): tcgen05.mma issue rate vs operand layout, N and the number of
// independent accumulators.  Timing only (operands are zeros).
#include <vector>
#include "kws_common.cuh"
using namespace kws;

struct Case {
  int N;
  int a_layout, b_layout;  // 0 none, 6 = 32B, 4 = 64B, 2 = 128B swizzle
  int nacc;                // independent accumulators cycled through
  int a_lbo, a_sbo, b_lbo, b_sbo;
  int a_step;              // start-address step (16B units) between consecutive MMAs, cycled mod 4
  int n_mma, batches;
};

__global__ void __launch_bounds__(192, 1) probe_kernel(Case c, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_f16(128, c.N, 0);
    const uint32_t sa = smem_u32(base), sb = smem_u32(base + 64 * 1024);
    const uint64_t adesc = make_smem_desc(sa, c.a_lbo, c.a_sbo, c.a_layout);
    const uint64_t bdesc = make_smem_desc(sb, c.b_lbo, c.b_sbo, c.b_layout);
    long long total = 0;
    uint32_t ph = 0;
    for (int b = 0; b < c.batches; ++b) {
      const long long t0 = clock64();
      // 7 MMAs per group with compile-time offsets (like the stem's dj loop); groups cycle accumulators
      for (int g = 0, acc = 0; g < c.n_mma / 7; ++g) {
        const uint32_t d = tmem + acc * c.N;
        const uint64_t a = adesc + (uint64_t)((g & 1) * 64);
        const uint64_t bb = bdesc + (uint64_t)((g & 3) * 896);
#pragma unroll
        for (int j = 0; j < 7; ++j)
          umma_f16(d, a + (uint64_t)((j >> 1) * c.a_step + (j & 1) * 264), bb + (uint64_t)(j * 128), idesc,
                   (g >= c.nacc) | (j > 0));
        if (++acc == c.nacc) acc = 0;
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph, 1);
      ph ^= 1;
      total += clock64() - t0;
    }
    out_cycles[blockIdx.x] = total;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = 1024 + 160 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * sms);
  std::vector<long long> h(sms);
  std::vector<Case> cases;
  const int NM = 448, NB = 10;
  for (int n : {16, 32, 48}) cases.push_back({n, 2, 2, 2, 16, 1024, 16, 1024, 2, NM, NB});
  // accumulators sweep, SW128 operands (GEMM-like k-slices)
  for (int nacc : {1, 2, 4, 8}) cases.push_back({64, 2, 2, nacc, 16, 1024, 16, 1024, 2, NM, NB});
  for (int nacc : {1, 2, 4}) cases.push_back({128, 2, 2, nacc, 16, 1024, 16, 1024, 2, NM, NB});
  for (int nacc : {1, 2}) cases.push_back({256, 2, 2, nacc, 16, 1024, 16, 1024, 2, NM, NB});
  // A = 32B-swizzle rows (pixel-major 16 channels), B = 32B swizzle; shifts of one pixel (2 x 16B)
  for (int nacc : {1, 2, 4, 8}) cases.push_back({64, 6, 6, nacc, 16, 256, 16, 256, 2, NM, NB});
  for (int nacc : {2, 4}) cases.push_back({128, 6, 6, nacc, 16, 256, 16, 256, 2, NM, NB});
  cases.push_back({192, 6, 6, 2, 16, 256, 16, 256, 2, NM, NB});
  cases.push_back({256, 6, 6, 2, 16, 256, 16, 256, 2, NM, NB});
  // A = 64B swizzle
  for (int nacc : {2, 8}) cases.push_back({64, 4, 4, nacc, 16, 512, 16, 512, 2, NM, NB});
  // no swizzle: canonical contiguous chunks (LBO = 128 rows x 16 B) and the stem's 2112
  for (int nacc : {2, 8}) cases.push_back({64, 0, 0, nacc, 2048, 128, 1024, 128, 1, NM, NB});
  for (int nacc : {2, 8}) cases.push_back({64, 0, 0, nacc, 2112, 128, 1024, 128, 1, NM, NB});
  // no swizzle A with B swizzled and vice versa (which operand is the slow one?)
  cases.push_back({64, 0, 6, 8, 2112, 128, 16, 256, 1, NM, NB});
  cases.push_back({64, 6, 0, 8, 16, 256, 1024, 128, 2, NM, NB});
  for (const Case& c : cases) {
    for (int rep = 0; rep < 2; ++rep) {
      probe_kernel<<<sms, 192, smem>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < sms; ++i) sum += (double)h[i];
    const double per = sum / sms / ((double)c.n_mma * c.batches);
    printf("N=%3d A=%d B=%d nacc=%d lbo/sbo A %d/%d B %d/%d : %6.1f cyc/MMA  floor %3.0f  (%.0f B/clk smem)\n", c.N,
           c.a_layout, c.b_layout, c.nacc, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, per, 128.0 * c.N / 256.0,
           (128 + c.N) * 32 / per);
  }
  return 0;
}
