// ResNet stem Conv2d(C,64,7x7,stride 2,pad 3) + BatchNorm2d(eval, folded) + ReLU
// as a tap-decomposed tcgen05 implicit GEMM (sm_100a).
//
//   out[p, oc, oi, oj] = relu(bias[oc] + sum_{c,di,dj} W'[oc,c,di,dj] * S[p, c, 2oi+di-3, 2oj+dj-3])
//
// GEMM view per output row oi and tile of 128 output columns:
//   M = 128 output pixels (consecutive oj), N = 64 output channels,
//   K = 16 input channels per MMA, one MMA per tap (di,dj) -> 49 MMAs accumulate
//   into one 128x64 fp32 TMEM accumulator.
//
// The stride-2 window is made MMA-addressable by de-interleaving the input
// columns into two parity planes in shared memory, channels innermost:
//   plane[pj][chunk][jj][8 ch]  (16 bytes per pixel per 8-channel chunk)
// so that for tap dj the 128 A-rows (oj = 0..127) are the 128 consecutive
// pixels jj = oj + (dj >> 1) of plane pj = dj & 1: a K-major, no-swizzle UMMA
// operand with SBO = 128 B (8 pixels) and LBO = plane chunk stride.  A tap only
// changes the descriptor start address; nothing is materialised (no im2col).
//
// Input rows stream through a ring of row-pair slots (each input row is read
// from global exactly once per column tile); weights (49 taps x 2 KB) stay
// resident in shared memory for the life of the CTA.
//
// Roles (352 threads): warp 0 = TMEM allocator, warp 1 = MMA issuer,
// warps 2..6 = loaders (global fp16 planar -> smem parity planes; one thread
// per plane pixel, software-pipelined one row pair ahead),
// warps 7..10 = epilogue (TMEM -> +bias, ReLU -> global).
#include "kws_common.cuh"
#include "../../include/kws_b200.h"

namespace kws {

constexpr int STEM_THREADS = 352;
constexpr int LOADER_THREADS = 160;        // warps 2..6; threads 0..131 own one plane pixel each
constexpr int FIRST_EPI_WARP = 7;
constexpr int OC = 64;
constexpr int TILE_OJ = 128;               // output columns per tile (= UMMA M)
constexpr int PLANE_JJ = 132;              // pixels per parity plane row: x = j - (256 ct - 4), jj = x >> 1
constexpr int X_ORIGIN = 4;                // tile column x = 0 is input column 256 ct - 4 (even: 32-bit loads)
constexpr int CHUNK_BYTES = PLANE_JJ * 16;  // 2112: one 8-channel chunk of one plane row
constexpr int PLANE_BYTES = 2 * CHUNK_BYTES;  // 4224: 16 channels
constexpr int ROW_BYTES = 2 * PLANE_BYTES;    // 8448: both parities
constexpr int SLOT_BYTES = 2 * ROW_BYTES;     // 16896: a pair of input rows
constexpr int RING_SLOTS = 7;
constexpr int TAP_BYTES = 2 * OC * 16;  // 2048: [chunk][oc][8 ch] fp16
constexpr int W_BYTES = 49 * TAP_BYTES;  // 100352
constexpr int NUM_ACC = 4;
constexpr int STEM_TMEM_COLS = NUM_ACC * OC;  // 256

struct StemParams {
  const __half* feat;  // [pairs, C, Tk, pitch]
  const uint4* w;      // packed weights of this channel group
  const float* bias;   // [64]
  void* out;
  int out_mode;
  long long pairs;
  int C, Tk, Tu, pitch, Ho, Wo, col_tiles;
  long long num_items;
  // channel groups (C > 16): one launch per 16-channel group, partial sums chained through an fp32
  // [pairs,Ho,Wo,64] workspace; bias + ReLU + the final store happen in the last group's launch
  int ch0;             // first input channel of this launch's group
  const float* acc_in;  // partial sums of the previous groups (null for the first group)
  float* acc_out;       // where to leave partial sums (null for the last group: write `out`)
};

__global__ void __launch_bounds__(STEM_THREADS, 1) kws_stem_kernel(const StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = base;                   // W_BYTES
  uint8_t* s_ring = base + W_BYTES;      // RING_SLOTS * SLOT_BYTES  (W_BYTES is a multiple of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + RING_SLOTS * SLOT_BYTES);
  uint64_t* full_bar = bars;                              // [RING_SLOTS] loaders -> MMA
  uint64_t* empty_bar = bars + RING_SLOTS;                // [RING_SLOTS] MMA (commit) -> loaders
  uint64_t* afull_bar = bars + 2 * RING_SLOTS;            // [NUM_ACC]   MMA (commit) -> epilogue
  uint64_t* aempty_bar = bars + 2 * RING_SLOTS + NUM_ACC;  // [NUM_ACC]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RING_SLOTS + 2 * NUM_ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NP = p.Ho + 3;  // row pairs per item: ring rows rr = 0 .. 2Ho+5 <-> input rows rr-3

  // resident weights (generic-proxy stores, made visible to the UMMA proxy below)
  for (int i = threadIdx.x; i < W_BYTES / 16; i += STEM_THREADS) reinterpret_cast<uint4*>(s_w)[i] = p.w[i];
  fence_proxy_async();

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < RING_SLOTS; ++s) {
      mbar_init(&full_bar[s], LOADER_THREADS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(&afull_bar[a], 1);
      mbar_init(&aempty_bar[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, STEM_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(TILE_OJ, OC, 0);
      const uint64_t adesc0 = make_smem_desc(smem_u32(s_ring), CHUNK_BYTES, 128, LAYOUT_NONE);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(s_w), OC * 16, 128, LAYOUT_NONE);
      uint32_t pair_seq = 0;  // global sequence number of the first row pair of this item
      uint32_t row_seq = 0;   // global output-row counter -> accumulator buffer
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x, pair_seq += NP) {
        uint32_t waited = 0;  // row pairs of this item already known to be in smem
        for (int oi = 0; oi < p.Ho; ++oi, ++row_seq) {
          const uint32_t acc = row_seq % NUM_ACC, acc_par = (row_seq / NUM_ACC) & 1;
          mbar_wait(&aempty_bar[acc], acc_par ^ 1, 200 + acc);
          while (waited <= (uint32_t)oi + 3) {
            const uint32_t g = pair_seq + waited;
            mbar_wait(&full_bar[g % RING_SLOTS], (g / RING_SLOTS) & 1, 300 + (int)(g % RING_SLOTS));
            ++waited;
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * OC;
          // descriptors differ from the two bases only in the 16-byte-granular start address
#pragma unroll 1
          for (int di = 0; di < 7; ++di) {
            const uint32_t g = pair_seq + oi + (di >> 1);
            const uint64_t a_row = adesc0 + (uint64_t)(((g % RING_SLOTS) * SLOT_BYTES + (di & 1) * ROW_BYTES) >> 4);
            const uint64_t b_row = bdesc0 + (uint64_t)((di * 7 * TAP_BYTES) >> 4);
#pragma unroll
            for (int dj = 0; dj < 7; ++dj) {
              // input column 2 oj + dj - 3  ->  x = 2 ojl + dj + 1: plane (dj+1)&1, pixel ojl + ((dj+1)>>1)
              umma_f16(d_tmem, a_row + (uint64_t)((((dj + 1) & 1) * PLANE_BYTES + ((dj + 1) >> 1) * 16) >> 4),
                       b_row + (uint64_t)((dj * TAP_BYTES) >> 4), idesc, (di | dj) != 0);
            }
          }
          // row pair `oi` is dead once these MMAs retire; the accumulator is complete
          umma_commit(&empty_bar[(pair_seq + oi) % RING_SLOTS]);
          umma_commit(&afull_bar[acc]);
        }
        // the last three row pairs (bottom halo) are released at the end of the item
        for (int s = p.Ho; s < NP; ++s) {
          while (waited <= (uint32_t)s) {  // only possible when Ho == 0 (never), kept for symmetry
            const uint32_t g = pair_seq + waited;
            mbar_wait(&full_bar[g % RING_SLOTS], (g / RING_SLOTS) & 1, 350);
            ++waited;
          }
          umma_commit(&empty_bar[(pair_seq + s) % RING_SLOTS]);
        }
      }
    }
  } else if (warp >= 2 && warp < FIRST_EPI_WARP) {
    // ===================== loaders =====================
    // Thread jj owns plane pixel jj of both parities: one 32-bit load fetches input columns
    // (j, j+1) = (even x, odd x) of one channel; 16 channels x 2 rows = 32 loads in flight per
    // thread, issued one row pair ahead of the stores (register double buffer).
    const int jj = threadIdx.x - 64;  // 0..159, active < PLANE_JJ
    const bool active = jj < PLANE_JJ;
    const long long my_items = (p.num_items - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long total = my_items * NP;  // row pairs this CTA streams
    const long long ch_stride = (long long)p.Tk * p.pitch;

    auto load_pair = [&](long long n, uint32_t (&v)[32]) {
      const long long it = blockIdx.x + (n / NP) * gridDim.x;
      const int s = (int)(n % NP);
      const long long pair = it / p.col_tiles;
      const int ct = (int)(it % p.col_tiles);
      const int j = ct * (2 * TILE_OJ) - X_ORIGIN + 2 * jj;  // even
      const bool lo_ok = active && j >= 0 && j < p.Tu;
      const uint32_t keep = (j + 1 < p.Tu) ? 0xffffffffu : 0x0000ffffu;
      const __half* f_pair = p.feat + (pair * (long long)p.C + p.ch0) * ch_stride + j;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = 2 * s + h - 3;
        const bool ok = lo_ok && r >= 0 && r < p.Tk;
        const __half* f_row = f_pair + (long long)r * p.pitch;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          uint32_t x = 0;
          if (ok && c + p.ch0 < p.C) x = __ldg(reinterpret_cast<const unsigned int*>(f_row + c * ch_stride)) & keep;
          v[h * 16 + c] = x;
        }
      }
    };
    auto store_pair = [&](long long n, const uint32_t (&v)[32]) {
      const uint32_t slot = (uint32_t)(n % RING_SLOTS);
      mbar_wait(&empty_bar[slot], (uint32_t)((n / RING_SLOTS) & 1) ^ 1, 100 + (int)slot);
      if (active) {
        uint8_t* dst = s_ring + slot * SLOT_BYTES + jj * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {  // 8-channel chunk
            const uint32_t* q = &v[h * 16 + k * 8];
            // plane 0 <- low halves (even x), plane 1 <- high halves (odd x)
            *reinterpret_cast<uint4*>(dst + h * ROW_BYTES + k * CHUNK_BYTES) =
                make_uint4(__byte_perm(q[0], q[1], 0x5410), __byte_perm(q[2], q[3], 0x5410),
                           __byte_perm(q[4], q[5], 0x5410), __byte_perm(q[6], q[7], 0x5410));
            *reinterpret_cast<uint4*>(dst + h * ROW_BYTES + PLANE_BYTES + k * CHUNK_BYTES) =
                make_uint4(__byte_perm(q[0], q[1], 0x7632), __byte_perm(q[2], q[3], 0x7632),
                           __byte_perm(q[4], q[5], 0x7632), __byte_perm(q[6], q[7], 0x7632));
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&full_bar[slot]);
    };

    uint32_t va[32], vb[32];
    if (total > 0) load_pair(0, va);
    for (long long n = 0; n < total; n += 2) {
      if (n + 1 < total) load_pair(n + 1, vb);
      store_pair(n, va);
      if (n + 2 < total) load_pair(n + 2, va);
      if (n + 1 < total) store_pair(n + 1, vb);
    }
  } else if (warp >= FIRST_EPI_WARP) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int ojl = q * 32 + lane;
    uint32_t row_seq = 0;
    float bias_r[OC];
#pragma unroll
    for (int i = 0; i < OC; ++i) bias_r[i] = __ldg(p.bias + i);
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const long long pair = it / p.col_tiles;
      const int ct = (int)(it % p.col_tiles);
      const int oj = ct * TILE_OJ + ojl;
      const bool ok = oj < p.Wo;
      for (int oi = 0; oi < p.Ho; ++oi, ++row_seq) {
        const uint32_t acc = row_seq % NUM_ACC, acc_par = (row_seq / NUM_ACC) & 1;
        mbar_wait(&afull_bar[acc], acc_par, 400 + acc);
        tc_fence_after();
        const uint32_t t_row = tmem_base + acc * OC + ((uint32_t)(q * 32) << 16);
        uint32_t v[4][16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) tmem_ld16(t_row + ch * 16, v[ch]);
        tmem_ld_wait();
        // accumulator is in registers: hand the TMEM buffer back before the stores
        tc_fence_before();
        mbar_arrive(&aempty_bar[acc]);
        if (ok && (p.acc_in || p.acc_out)) {
          const long long px = ((pair * p.Ho + oi) * (long long)p.Wo + oj) * OC;
          if (p.acc_in) {
            const float4* a = reinterpret_cast<const float4*>(p.acc_in + px);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 f = a[ch * 4 + e];
                v[ch][4 * e] = __float_as_uint(__uint_as_float(v[ch][4 * e]) + f.x);
                v[ch][4 * e + 1] = __float_as_uint(__uint_as_float(v[ch][4 * e + 1]) + f.y);
                v[ch][4 * e + 2] = __float_as_uint(__uint_as_float(v[ch][4 * e + 2]) + f.z);
                v[ch][4 * e + 3] = __float_as_uint(__uint_as_float(v[ch][4 * e + 3]) + f.w);
              }
          }
          if (p.acc_out) {
            uint4* a = reinterpret_cast<uint4*>(p.acc_out + px);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                a[ch * 4 + e] = make_uint4(v[ch][4 * e], v[ch][4 * e + 1], v[ch][4 * e + 2], v[ch][4 * e + 3]);
          }
        }
        if (ok && !p.acc_out) {
          if (p.out_mode == KWS_STEM_OUT_NCHW_F32) {
            float* o = reinterpret_cast<float*>(p.out) + ((pair * OC) * p.Ho + oi) * (long long)p.Wo + oj;
            const long long oc_stride = (long long)p.Ho * p.Wo;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
#pragma unroll
              for (int e = 0; e < 16; ++e)
                o[(ch * 16 + e) * oc_stride] = fmaxf(__uint_as_float(v[ch][e]) + bias_r[ch * 16 + e], 0.f);
          } else {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                                ((pair * p.Ho + oi) * (long long)p.Wo + oj) * OC);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float a = fmaxf(__uint_as_float(v[ch][2 * e]) + bias_r[ch * 16 + 2 * e], 0.f);
                const float b = fmaxf(__uint_as_float(v[ch][2 * e + 1]) + bias_r[ch * 16 + 2 * e + 1], 0.f);
                pk[e] = pack_bf162(a, b);
              }
              o[ch * 2] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              o[ch * 2 + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, STEM_TMEM_COLS);
  }
}

constexpr size_t STEM_SMEM = 1024 + W_BYTES + RING_SLOTS * SLOT_BYTES + (2 * RING_SLOTS + 2 * NUM_ACC) * 8 + 16;

}  // namespace kws

using namespace kws;

extern "C" size_t kws_stem_workspace_bytes(int pairs, int C, int Tk, int Tu) {
  if (C <= 16 || pairs <= 0 || Tk <= 0 || Tu <= 0) return 0;
  return (size_t)pairs * ((Tk + 1) / 2) * ((Tu + 1) / 2) * OC * sizeof(float);
}

extern "C" int kws_stem(const void* feat_f16, int pairs, int C, int Tk, int Tu, int pitch16, const void* w_packed,
                        const float* bias, int out_mode, void* out, void* workspace, void* stream) {
  KWS_CHECK_ARG(feat_f16 && w_packed && bias && out, "stem: null pointer");
  KWS_CHECK_ARG(pairs > 0 && Tk > 0 && Tu > 0, "stem: non-positive dimension");
  KWS_CHECK_ARG(C > 0 && C <= 64, "stem: C=%d input channels out of (0,64]", C);
  KWS_CHECK_ARG(C <= 16 || workspace, "stem: C=%d > 16 needs a workspace of kws_stem_workspace_bytes()", C);
  KWS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "stem: workspace must be 16-byte aligned");
  KWS_CHECK_ARG(pitch16 >= Tu, "stem: pitch16=%d < Tu=%d", pitch16, Tu);
  KWS_CHECK_ARG(out_mode == KWS_STEM_OUT_NCHW_F32 || out_mode == KWS_STEM_OUT_NHWC_BF16, "stem: bad out_mode %d",
                out_mode);
  KWS_CHECK_ARG((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "stem: pointers must be 16-byte aligned");
  StemParams p{};
  p.feat = reinterpret_cast<const __half*>(feat_f16);
  p.w = reinterpret_cast<const uint4*>(w_packed);
  p.bias = bias;
  p.out = out;
  p.out_mode = out_mode;
  p.pairs = pairs;
  p.C = C, p.Tk = Tk, p.Tu = Tu, p.pitch = pitch16;
  p.Ho = (Tk + 1) / 2;
  p.Wo = (Tu + 1) / 2;
  p.col_tiles = (p.Wo + TILE_OJ - 1) / TILE_OJ;
  p.num_items = (long long)pairs * p.col_tiles;
  KWS_CUDA(cudaFuncSetAttribute(kws_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEM_SMEM));
  long long grid = p.num_items;
  const int sms = sm_count();
  if (grid > sms) grid = sms;
  const int groups = (C + 15) / 16;
  for (int g = 0; g < groups; ++g) {
    p.ch0 = g * 16;
    p.w = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(w_packed) + (size_t)g * W_BYTES);
    p.acc_in = g > 0 ? reinterpret_cast<const float*>(workspace) : nullptr;
    p.acc_out = g + 1 < groups ? reinterpret_cast<float*>(workspace) : nullptr;
    kws_stem_kernel<<<(int)grid, STEM_THREADS, STEM_SMEM, (cudaStream_t)stream>>>(p);
    KWS_CUDA(cudaGetLastError());
  }
  return 0;
}
