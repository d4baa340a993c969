// Probe (development tool): semantics and cost of tcgen05.shift.cta_group::1.down on sm_100a.
// Question: can the fused kernel's half-b shift (out[x] = D_a[x] + D_b[x+2], today 64 SHFL per thread and step + a
// mailbox) be done in tensor memory instead -- shift the half-a columns DOWN by two rows, so that row m holds out[m-2]?
// The PTX text says "shifts 32-byte elements down across all the rows, except the last, by one row; the lane of taddr
// must be aligned to 32".  Unknown: how many rows and columns one instruction covers, and what it costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I enhance-cb-whisper_b200/csrc \
//        tools/shift_probe.cu enhance-cb-whisper_b200/csrc/kws_abi.cu -o tools/shift_probe
#include <vector>
#include "kws_common.cuh"
using namespace kws;

__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// 128 threads (4 warps = 128 TMEM lanes).  Fill columns [0, 32) with value = lane * 1000 + column, issue n_shift shifts at
// (lane_base, col_base), read back.  out[lane * 32 + col]; timing of `reps` shifts + commit in out_t.
__global__ void __launch_bounds__(128, 1) probe(int lane_base, int col_base, int n_shift, int reps, uint32_t* out, long long* out_t) {
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t t_lane = tb + ((uint32_t)(warp * 32) << 16);
  for (int h = 0; h < 2; ++h) {
    uint32_t v[16];
    for (int e = 0; e < 16; ++e) v[e] = (uint32_t)((warp * 32 + lane) * 1000 + h * 16 + e);
    st16(t_lane + h * 16, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long dt = 0;
  if (threadIdx.x == 0) {
    const uint32_t ta = tb + ((uint32_t)lane_base << 16) + (uint32_t)col_base;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int s = 0; s < n_shift; ++s) asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(ta) : "memory");
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    dt = clock64() - t0;
    out_t[0] = dt;
  }
  __syncthreads();
  tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t v[16];
    tmem_ld16(t_lane + h * 16, v);
    tmem_ld_wait();
    for (int e = 0; e < 16; ++e) out[(warp * 32 + lane) * 32 + h * 16 + e] = v[e];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_slot, 512); }
}

int main() {
  uint32_t* d;
  long long* dt;
  cudaMalloc(&d, 128 * 32 * 4);
  cudaMalloc(&dt, 64);
  std::vector<uint32_t> h(128 * 32);
  auto run = [&](int lane_base, int col_base, int n_shift, int reps, bool show) {
    probe<<<1, 128>>>(lane_base, col_base, n_shift, reps, d, dt);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lane_base %d col_base %d: launch failed: %s\n", lane_base, col_base, cudaGetErrorString(e)); return false; }
    long long c;
    cudaMemcpy(&c, dt, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    printf("lane_base %3d col_base %2d shifts %d x %d: %lld cycles (%.1f per shift incl. commit)\n", lane_base, col_base, n_shift, reps, c,
           (double)c / (n_shift * reps));
    if (show) {
      // which (lane, col) changed, and where did the value come from?
      int changed = 0, cmin = 99, cmax = -1, lmin = 999, lmax = -1;
      for (int l = 0; l < 128; ++l)
        for (int c2 = 0; c2 < 32; ++c2) {
          const uint32_t v = h[l * 32 + c2], exp = (uint32_t)(l * 1000 + c2);
          if (v != exp) { ++changed; cmin = c2 < cmin ? c2 : cmin; cmax = c2 > cmax ? c2 : cmax; lmin = l < lmin ? l : lmin; lmax = l > lmax ? l : lmax; }
        }
      printf("   changed entries %d: lanes [%d, %d], columns [%d, %d]\n", changed, lmin, lmax, cmin, cmax);
      for (int l : {0, 1, 2, 3, 30, 31, 32, 33, 34, 62, 63, 64, 65, 66, 95, 96, 97, 126, 127}) {
        const uint32_t v = h[l * 32 + (col_base & 31)];
        printf("   lane %3d col %2d: %6u  (= lane %u col %u)\n", l, col_base & 31, v, v / 1000, v % 1000);
      }
    }
    return true;
  };
  run(0, 0, 1, 1, true);    // one shift of the whole matrix at column 0
  run(0, 8, 2, 1, true);    // two shifts at column 8
  run(32, 16, 1, 1, true);  // lane base 32
  run(0, 0, 16, 100, false);  // cost of 16 shifts (what one stem step would need)
  run(0, 0, 1, 1000, false);
  return 0;
}
