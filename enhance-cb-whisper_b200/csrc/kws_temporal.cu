// LEF temporal projector on the tensor cores (sm_100a):
//     Conv1d(P, P, 3, padding 1) -> BatchNorm1d (eval, folded) -> MaxPool1d(3, 2, 1) over the frame axis,
//     then the L2 normalisation and mask fold of the similarity prologue            (model.py:107-124, :152-166, :210-218)
//
// The convolution over time is an implicit GEMM whose three taps read the SAME shared-memory tile shifted by one
// row each: the tile of projected frames is stored K-chunk-major, [chunk of 8 features][row][16 B], so that rows are
// 16 bytes apart inside a chunk (no-swizzle canonical layout with SBO = 128 B) and "one frame later" is a +16 byte
// start address in the A descriptor -- the same trick the fused stem uses for its horizontal taps.
//
//     y[r, :] = b' + sum_{d=0..2} x[r + d - 1, :] W'_d^T          3 x (P/16) MMAs of 128 x P x 16 per tile
//
// Rows are the flat sequence of all frames of all items of one layer with one virtual zero row before and after every
// item (the Conv1d zero padding; virtual row v -> item v / (T+2), position v % (T+2), 0 and T+1 are the pads), so
// tiles run across item boundaries and short items (150-frame keywords) waste nothing.  Tile i computes the conv rows
// [126 i, 126 i + 128) and owns the pooled frames whose centre row lies in [126 i + 1, 126 i + 127).
//
// One CTA per (tile, layer); 256 threads: all load, one elected thread issues the MMAs, all 8 warps drain TMEM
// (bias added) into a padded fp32 tile in shared memory, then one warp per pooled frame does max-of-3, the norm
// and the fp16 store (128-byte rows).
#include "kws_common.cuh"
#include "../../include/kws_b200.h"

namespace kws {

constexpr int TP_THREADS = 256;
constexpr int TP_ROWS = 128;   // conv rows per tile (MMA M)
constexpr int TP_STEP = 126;   // tile stride in virtual rows
constexpr int TP_AROWS = 130;  // input rows held per tile: conv rows -1 .. +128

struct TemporalParams {
  const uint4* proj;   // 16-bit [C, B*T, P]
  const uint4* w16;    // 16-bit [C][3][P/8][P][8]  (kws_fold_temporal_weights)
  const float* bias;   // [C, P]
  const float* mask;   // [B, C, T2] or null
  __half* out;         // fp16 [C, B, T2, P]
  int B, T, P, T2, C;
  long long Rv;        // virtual rows per layer = B * (T + 2)
  float eps;
  uint32_t idesc;
  uint32_t tmem_cols;
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(TP_THREADS) temporal_mma_kernel(const TemporalParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int P = p.P, chunks = P >> 3;
  const int a_chunk_bytes = TP_AROWS * 16;
  uint8_t* sA = smem;                                       // [chunks][130][16 B]
  uint8_t* sW = sA + chunks * a_chunk_bytes;                // [3][chunks][P][16 B]
  float* sY = reinterpret_cast<float*>(sW + 3 * P * P * 2); // [128][P + 1]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sY + TP_ROWS * (P + 1) + ((TP_ROWS * (P + 1)) & 1));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.y;
  const long long v0 = (long long)blockIdx.x * TP_STEP;  // virtual row of conv row 0 of this tile
  const int Tp = p.T + 2;

  if (warp == 0) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  if (tid == 32) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  // ---- operands -> shared memory ----
  {
    const uint4* wsrc = p.w16 + (size_t)c * (3 * P * P / 8);
    uint4* wdst = reinterpret_cast<uint4*>(sW);
    for (int i = tid; i < 3 * P * P / 8; i += TP_THREADS) wdst[i] = __ldg(wsrc + i);
    const uint4* xsrc = p.proj + (size_t)c * p.B * p.T * chunks;
    for (int i = tid; i < TP_AROWS * chunks; i += TP_THREADS) {
      const int j = i / chunks, ch = i - j * chunks;
      const long long v = v0 - 1 + j;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (v >= 0 && v < p.Rv) {
        const long long b = v / Tp;
        const int pos = (int)(v - b * Tp);
        if (pos >= 1 && pos <= p.T) val = __ldg(xsrc + (b * p.T + (pos - 1)) * chunks + ch);
      }
      *reinterpret_cast<uint4*>(sA + ch * a_chunk_bytes + j * 16) = val;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- 3 taps x P/16 MMAs, one elected thread ----
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW);
      int n = 0;
      for (int d = 0; d < 3; ++d) {
        for (int kk = 0; kk < P / 16; ++kk, ++n) {
          const uint64_t adesc = make_smem_desc(a0 + 2 * kk * a_chunk_bytes + d * 16, a_chunk_bytes, 128, LAYOUT_NONE);
          const uint64_t bdesc = make_smem_desc(w0 + d * (P * P * 2) + 2 * kk * (P * 16), P * 16, 128, LAYOUT_NONE);
          umma_f16(tmem_base, adesc, bdesc, p.idesc, n != 0);
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
  }

  // ---- TMEM -> (+ bias) -> fp32 tile in shared memory; warps w and w+4 share a lane quarter, split the columns ----
  mbar_wait(bar, 0, 900);
  tc_fence_after();
  {
    const int quarter = warp & 3, half = warp >> 2;
    const int m = quarter * 32 + lane;
    const int col0 = half * (P >> 1);
    const float* bias = p.bias + (size_t)c * P;
    float* yrow = sY + m * (P + 1);
    for (int cc = 0; cc < (P >> 1); cc += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + col0 + cc, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) yrow[col0 + cc + e] = __uint_as_float(v[e]) + __ldg(bias + col0 + cc + e);
    }
  }
  tc_fence_before();
  __syncthreads();

  // ---- MaxPool1d(3,2,1) (-inf padding: out-of-item rows skipped), L2 norm, mask, fp16 ----
  for (int m = 1 + warp; m < TP_ROWS - 1; m += TP_THREADS / 32) {
    const long long q = v0 + m;
    if (q >= p.Rv) break;
    const long long b = q / Tp;
    const int t = (int)(q - b * Tp) - 1;  // frame of the centre row
    if (t < 0 || t >= p.T || (t & 1)) continue;
    const int t2 = t >> 1;
    float vals[4];  // P <= 128
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int col = lane + 32 * k;
      vals[k] = 0.f;
      if (col < P) {
        float mx = sY[m * (P + 1) + col];
        if (t - 1 >= 0) mx = fmaxf(mx, sY[(m - 1) * (P + 1) + col]);
        if (t + 1 < p.T) mx = fmaxf(mx, sY[(m + 1) * (P + 1) + col]);
        vals[k] = mx;
        ss = fmaf(mx, mx, ss);
      }
    }
    ss = warp_sum_f(ss);
    const float mk = p.mask ? __ldg(p.mask + ((size_t)b * p.C + c) * p.T2 + t2) : 1.f;
    const float scale = mk / fmaxf(sqrtf(ss), p.eps);
    __half* o = p.out + (((size_t)c * p.B + b) * p.T2 + t2) * P;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int col = lane + 32 * k;
      if (col < P) o[col] = __float2half_rn(vals[k] * scale);
    }
  }

  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// BN fold + 16-bit pack of the temporal conv: conv_w fp32 [C][P out][P in][3] -> w16 [C][3][P/8][P out][8 in]
// (the kernel's shared-memory image: B operand of tap d, K-chunk-major), b' = (conv_b - mean) * s + beta.
__global__ void fold_temporal_kernel(const float* __restrict__ w, const float* __restrict__ cb,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                     int C, int P, int bf16, uint16_t* __restrict__ w16, float* __restrict__ bf) {
  const int total = C * 3 * P * P;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    int r = i >> 3;
    const int po = r % P;
    r /= P;
    const int ch = r % (P / 8);
    r /= (P / 8);
    const int d = r % 3, c = r / 3;
    const int pi = ch * 8 + e;
    const float s = gamma[c * P + po] / sqrtf(var[c * P + po] + eps);
    const float v = w[(((size_t)c * P + po) * P + pi) * 3 + d] * s;
    w16[i] = (uint16_t)((bf16 ? pack_bf162(v, 0.f) : pack_half2_sat(v, 0.f)) & 0xffffu);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * P; i += gridDim.x * blockDim.x) {
    const float s = gamma[i] / sqrtf(var[i] + eps);
    bf[i] = (cb[i] - mean[i]) * s + beta[i];
  }
}

}  // namespace kws

using namespace kws;

extern "C" {

int kws_temporal(const void* proj16, int C, int B, int T, int P, int dtype16, const void* w16_packed,
                 const float* b_folded, const float* mask, float eps, void* out_f16, void* stream) {
  KWS_CHECK_ARG(proj16 && w16_packed && b_folded && out_f16, "temporal: null pointer");
  KWS_CHECK_ARG(C > 0 && B > 0 && T > 0, "temporal: non-positive dimension");
  KWS_CHECK_ARG(dtype16 == KWS_F16 || dtype16 == KWS_BF16, "temporal: bad dtype16 %d", dtype16);
  KWS_CHECK_ARG(P >= 32 && P <= 128 && P % 32 == 0, "temporal: P=%d must be a multiple of 32 in [32,128]", P);
  KWS_CHECK_ARG(C <= 65535, "temporal: C must be <= 65535 per launch");
  KWS_CHECK_ARG(((reinterpret_cast<uintptr_t>(proj16) | reinterpret_cast<uintptr_t>(w16_packed)) & 15) == 0,
                "temporal: pointers must be 16-byte aligned");
  TemporalParams p{};
  p.proj = reinterpret_cast<const uint4*>(proj16);
  p.w16 = reinterpret_cast<const uint4*>(w16_packed);
  p.bias = b_folded;
  p.mask = mask;
  p.out = reinterpret_cast<__half*>(out_f16);
  p.B = B, p.T = T, p.P = P, p.T2 = (T + 1) / 2, p.C = C;
  p.Rv = (long long)B * (T + 2);
  p.eps = eps;
  p.idesc = make_idesc_f16(TP_ROWS, P, (uint32_t)dtype16);
  p.tmem_cols = P <= 32 ? 32 : (P <= 64 ? 64 : 128);
  const long long tiles = (p.Rv - 1 + TP_STEP - 1) / TP_STEP;
  KWS_CHECK_ARG(tiles < (1ll << 31), "temporal: too many rows for one launch");
  const size_t n_y = (size_t)TP_ROWS * (P + 1);
  const size_t smem = (size_t)(P / 8) * TP_AROWS * 16 + (size_t)3 * P * P * 2 + (n_y + (n_y & 1)) * 4 + 16;
  KWS_CUDA(cudaFuncSetAttribute(temporal_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)tiles, (unsigned)C);
  temporal_mma_kernel<<<grid, TP_THREADS, smem, (cudaStream_t)stream>>>(p);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

int kws_fold_temporal_weights(const float* conv_w, const float* conv_b, const float* gamma, const float* beta,
                              const float* mean, const float* var, float eps, int C, int P, int dtype16,
                              void* w16_packed, float* b_folded, void* stream) {
  KWS_CHECK_ARG(conv_w && conv_b && gamma && beta && mean && var && w16_packed && b_folded,
                "fold_temporal: null pointer");
  KWS_CHECK_ARG(C > 0 && P > 0 && P % 8 == 0, "fold_temporal: need C > 0 and P a positive multiple of 8");
  KWS_CHECK_ARG(dtype16 == KWS_F16 || dtype16 == KWS_BF16, "fold_temporal: bad dtype16 %d", dtype16);
  fold_temporal_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(conv_w, conv_b, gamma, beta, mean, var, eps, C, P,
                                                             dtype16 == KWS_BF16, (uint16_t*)w16_packed, b_folded);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
