// tcgen05 / TMEM / TMA GEMM kernel of the efficient_kws path (sm_100a).
//
// One persistent, warp-specialised kernel computes D[128 x N] tiles of
//      D = A[rows_a, Kd] * B[rows_b, Kd]^T       (both operands K-major, 16-bit)
// with fp32 accumulation in tensor memory and three fused epilogues:
//   EPI_BIAS_RELU_BF16 : projector Linear(D,H)+ReLU           -> bf16 hidden
//   EPI_BIAS_OUT       : projector Linear(H,P) (+ normalise * mask -> fp16,
//                        or raw fp32 for the LEF temporal stage)
//   EPI_SIM            : cosine-similarity tile of one (keyword, utterance,
//                        layer): rows = utterance frames, cols = keyword frames,
//                        stored transposed as [Tk, Tu] fp32 and/or fp16
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// lane), warp 2 = TMEM allocator, warps 4..7 = epilogue (one TMEM lane quarter
// each).  smem ring of NUM_STAGES {A 128x64, B Nx64} 128B-swizzled tiles;
// two TMEM accumulator buffers so the epilogue of tile i overlaps the MMAs of
// tile i+1.
#include "kws_common.cuh"
#include "../../include/kws_b200.h"

namespace kws {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;  // columns between the two accumulator buffers

enum { EPI_BIAS_RELU_BF16 = 0, EPI_BIAS_OUT = 1, EPI_SIM = 2 };

struct GemmParams {
  int epi;
  int num_kblocks;  // Kd / 64
  int block_n;      // N of the tile (multiple of 16, <= 256)
  int num_stages;
  uint32_t idesc;
  long long num_items;
  // MLP tiling: item -> (batch c, m tile, n tile)
  int m_tiles, n_tiles;
  int rows;    // valid rows of A per batch (R, or Tu for SIM)
  int cols;    // valid output columns (H / P, or Tk for SIM)
  // outputs
  void* out;         // EPI 0: bf16 [C,R,cols]; EPI 1: fp16/fp32 [C,R,cols]; SIM: fp32 features or null
  void* out2;        // SIM: fp16 features or null
  const float* bias;  // [C, cols]
  const float* mask;  // EPI 1 normalised: [B, C, T] or null
  int out_mode;       // EPI 1: KWS_MLP_OUT_*
  int hidden_bf16;    // EPI 0 / RAW_16: 16-bit outputs stored as bf16 (else saturating fp16)
  int T, Cn;          // EPI 1 mask indexing: row r -> (b = r / T, t = r % T); Cn = layers
  float eps;
  // SIM
  int K, U, C, Tk, Tu, pitch16, diag;
  int operand_out;  // SIM: out2 = fp16 [C, pairs, Tu, block_n] (the accumulator tile as it is: a K-major operand)
};

struct Item {
  int a_row, a_batch, b_row, b_batch;
  int c, mt, nt;   // MLP
  long long pair;  // SIM
};

__device__ __forceinline__ Item decode(const GemmParams& p, long long it) {
  Item r;
  if (p.epi == EPI_SIM) {
    const int mt = (int)(it % p.m_tiles);
    long long q = it / p.m_tiles;
    const int c = (int)(q % p.C);
    q /= p.C;
    const int kw = (int)(q % p.K);
    const int u = p.diag ? kw : (int)(q / p.K);
    r.mt = mt, r.nt = 0, r.c = c;
    r.a_row = mt * BLOCK_M, r.a_batch = c * p.U + u;
    r.b_row = 0, r.b_batch = c * p.K + kw;
    r.pair = p.diag ? (long long)kw : (long long)kw * p.U + u;
  } else {
    const int nt = (int)(it % p.n_tiles);
    long long q = it / p.n_tiles;
    const int mt = (int)(q % p.m_tiles);
    const int c = (int)(q / p.m_tiles);
    r.mt = mt, r.nt = nt, r.c = c;
    r.a_row = mt * BLOCK_M, r.a_batch = c;
    r.b_row = nt * p.block_n, r.b_batch = c;
    r.pair = 0;
  }
  return r;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
kws_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A | B)] [barriers]
  const uint32_t stage_bytes = A_TILE_BYTES + p.block_n * BLOCK_K * 2;
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)p.num_stages * stage_bytes);
  uint64_t* full_bar = bars;                         // [stages]
  uint64_t* empty_bar = bars + p.num_stages;         // [stages]
  uint64_t* tfull_bar = bars + 2 * p.num_stages;     // [2]
  uint64_t* tempty_bar = bars + 2 * p.num_stages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.num_stages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const Item w = decode(p, it);
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          uint8_t* sa = tiles + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_3d(&map_a, &full_bar[stage], sa, kb * BLOCK_K, w.a_row, w.a_batch);
          tma_load_3d(&map_b, &full_bar[stage], sb, kb * BLOCK_K, w.b_row, w.b_batch);
          if (++stage == p.num_stages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t iter = 0;
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x, ++iter) {
        const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 200 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + A_TILE_BYTES;
          const uint64_t adesc = make_smem_desc(sa, 16, 1024, LAYOUT_SW128);
          const uint64_t bdesc = make_smem_desc(sb, 16, 1024, LAYOUT_SW128);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes (16 elements) inside the 128B swizzle row
            umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == p.num_stages) stage = 0, phase ^= 1;
        }
        umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t iter = 0;
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x, ++iter) {
      const Item w = decode(p, it);
      const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      const int row = w.a_row + q * 32 + lane;  // row within the batch
      const bool row_ok = row < p.rows;
      const int n_chunks = p.block_n >> 4;
      uint32_t v[16];

      if (p.epi == EPI_SIM && p.operand_out) {
        // the tile as a K-major fp16 operand of a following GEMM: row = utterance frame, block_n keyword frames
        // contiguous (config #4: contracted with the resize's height map inside the fused kernel)
        const long long n_pairs = p.diag ? (long long)p.K : (long long)p.K * p.U;
        __half* o = reinterpret_cast<__half*>(p.out2) + (((long long)w.c * n_pairs + w.pair) * p.Tu + row) * p.block_n;
        for (int ch = 0; ch < n_chunks; ++ch) {
          tmem_ld16(t_row + ch * 16, v);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) pk[e] = pack_half2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
            uint4* dst = reinterpret_cast<uint4*>(o + ch * 16);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      } else if (p.epi == EPI_SIM) {
        // row = utterance frame j, column = keyword frame i; out[pair][c][i][j]
        float* o32 = p.out ? reinterpret_cast<float*>(p.out) + ((w.pair * p.C + w.c) * p.Tk) * (long long)p.Tu + row
                           : nullptr;
        __half* o16 = p.out2 ? reinterpret_cast<__half*>(p.out2) +
                                   ((w.pair * p.C + w.c) * p.Tk) * (long long)p.pitch16 + row
                             : nullptr;
        for (int ch = 0; ch < n_chunks; ++ch) {
          tmem_ld16(t_row + ch * 16, v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int i = ch * 16 + e;
              if (i < p.Tk) {
                const float s = __uint_as_float(v[e]);
                if (o32) o32[(long long)i * p.Tu] = s;
                if (o16) o16[(long long)i * p.pitch16] = __float2half_rn(s);
              }
            }
          }
        }
      } else if (p.epi == EPI_BIAS_RELU_BF16) {
        const int col0 = w.nt * p.block_n;
        const float* bias = p.bias + (long long)w.c * p.cols + col0;
        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + ((long long)w.c * p.rows + row) * p.cols + col0;
        for (int ch = 0; ch < n_chunks; ++ch) {
          tmem_ld16(t_row + ch * 16, v);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float a = fmaxf(__uint_as_float(v[2 * e]) + __ldg(bias + ch * 16 + 2 * e), 0.f);
              const float b = fmaxf(__uint_as_float(v[2 * e + 1]) + __ldg(bias + ch * 16 + 2 * e + 1), 0.f);
              pk[e] = p.hidden_bf16 ? pack_bf162(a, b) : pack_half2_sat(a, b);
            }
            uint4* dst = reinterpret_cast<uint4*>(o + ch * 16);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      } else {  // EPI_BIAS_OUT (single n tile: block_n == cols)
        const float* bias = p.bias + (long long)w.c * p.cols;
        if (p.out_mode == KWS_MLP_OUT_RAW_F32) {
          float* o = reinterpret_cast<float*>(p.out) + ((long long)w.c * p.rows + row) * p.cols;
          for (int ch = 0; ch < n_chunks; ++ch) {
            tmem_ld16(t_row + ch * 16, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                float4 f;
                f.x = __uint_as_float(v[e]) + __ldg(bias + ch * 16 + e);
                f.y = __uint_as_float(v[e + 1]) + __ldg(bias + ch * 16 + e + 1);
                f.z = __uint_as_float(v[e + 2]) + __ldg(bias + ch * 16 + e + 2);
                f.w = __uint_as_float(v[e + 3]) + __ldg(bias + ch * 16 + e + 3);
                *reinterpret_cast<float4*>(o + ch * 16 + e) = f;
              }
            }
          }
        } else if (p.out_mode == KWS_MLP_OUT_RAW_16) {
          // un-normalised 16-bit rows for the temporal projector (fp16 saturates instead of overflowing)
          uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + ((long long)w.c * p.rows + row) * p.cols;
          for (int ch = 0; ch < n_chunks; ++ch) {
            tmem_ld16(t_row + ch * 16, v);
            tmem_ld_wait();
            if (row_ok) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float a = __uint_as_float(v[2 * e]) + __ldg(bias + ch * 16 + 2 * e);
                const float b = __uint_as_float(v[2 * e + 1]) + __ldg(bias + ch * 16 + 2 * e + 1);
                pk[e] = p.hidden_bf16 ? pack_bf162(a, b) : pack_half2_sat(a, b);
              }
              uint4* dst = reinterpret_cast<uint4*>(o + ch * 16);
              dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        } else {
          // pass 1: squared norm of the row; pass 2: re-read TMEM, scale, store
          float ss = 0.f;
          for (int ch = 0; ch < n_chunks; ++ch) {
            tmem_ld16(t_row + ch * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float a = __uint_as_float(v[e]) + __ldg(bias + ch * 16 + e);
              ss = fmaf(a, a, ss);
            }
          }
          float scale = 0.f;
          if (row_ok) {
            float m = 1.f;
            if (p.mask) {
              const int b = row / p.T, t = row - b * p.T;
              m = p.mask[((long long)b * p.Cn + w.c) * p.T + t];
            }
            scale = m / fmaxf(sqrtf(ss), p.eps);
          }
          __half* o = reinterpret_cast<__half*>(p.out) + ((long long)w.c * p.rows + row) * p.cols;
          for (int ch = 0; ch < n_chunks; ++ch) {
            tmem_ld16(t_row + ch * 16, v);
            tmem_ld_wait();
            if (row_ok) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float a = (__uint_as_float(v[2 * e]) + __ldg(bias + ch * 16 + 2 * e)) * scale;
                const float b = (__uint_as_float(v[2 * e + 1]) + __ldg(bias + ch * 16 + 2 * e + 1)) * scale;
                pk[e] = pack_half2(a, b);
              }
              uint4* dst = reinterpret_cast<uint4*>(o + ch * 16);
              dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      }
      // release the accumulator buffer to the MMA warp
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static size_t gemm_smem_bytes(int block_n, int stages) {
  return 1024 + (size_t)stages * (A_TILE_BYTES + block_n * BLOCK_K * 2) + (2 * stages + 4) * 8 + 16;
}

static int pick_stages(int block_n, int num_kblocks) {
  int s = 6;
  while (s > 2 && gemm_smem_bytes(block_n, s) > 200 * 1024) --s;
  if (s > num_kblocks + 2) s = num_kblocks + 2;  // tiny K: still prefetch the next items
  if (s < 2) s = 2;
  return s;
}

static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, GemmParams& p, cudaStream_t st) {
  p.num_stages = pick_stages(p.block_n, p.num_kblocks);
  const size_t smem = gemm_smem_bytes(p.block_n, p.num_stages);
  KWS_CUDA(cudaFuncSetAttribute(kws_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = p.num_items;
  const int sms = sm_count();
  if (grid > sms) grid = sms;
  if (grid <= 0) return 0;
  kws_gemm_kernel<<<(int)grid, GEMM_THREADS, smem, st>>>(ma, mb, p);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

// K-major 16-bit operand [batches, rows, kd] as a 3-D map with a (64, box_rows, 1) box
static int operand_map(CUtensorMap* m, const void* base, int kd, int rows, long long batches, int box_rows) {
  const uint64_t dims[3] = {(uint64_t)kd, (uint64_t)rows, (uint64_t)batches};
  const uint64_t strides[2] = {(uint64_t)kd * 2, (uint64_t)kd * 2 * (uint64_t)rows};
  const uint32_t box[3] = {BLOCK_K, (uint32_t)box_rows, 1};
  return make_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, base, dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace kws

using namespace kws;

extern "C" {

int kws_mlp(const void* x16, int C, int B, int T, int D, int H, int P, int dtype16, const void* w1_16,
            const float* b1, const void* w2_16, const float* b2, void* hidden16, const float* mask, float eps,
            int out_mode, void* out, void* stream) {
  KWS_CHECK_ARG(x16 && w1_16 && b1 && w2_16 && b2 && hidden16 && out, "mlp: null pointer");
  KWS_CHECK_ARG(dtype16 == KWS_F16 || dtype16 == KWS_BF16, "mlp: bad dtype16 %d", dtype16);
  KWS_CHECK_ARG(C > 0 && B > 0 && T > 0, "mlp: non-positive dimension");
  KWS_CHECK_ARG(D % 64 == 0 && D >= 64, "mlp: D=%d must be a multiple of 64", D);
  KWS_CHECK_ARG(H % 64 == 0 && H >= 64, "mlp: H=%d must be a multiple of 64", H);
  KWS_CHECK_ARG(P % 16 == 0 && P >= 16 && P <= 256, "mlp: P=%d must be a multiple of 16 in [16,256]", P);
  KWS_CHECK_ARG(out_mode == KWS_MLP_OUT_NORM_F16 || out_mode == KWS_MLP_OUT_RAW_F32 || out_mode == KWS_MLP_OUT_RAW_16,
                "mlp: bad out_mode %d", out_mode);
  const long long R = (long long)B * T;
  KWS_CHECK_ARG(R < (1ll << 31), "mlp: B*T too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int m_tiles = (int)((R + BLOCK_M - 1) / BLOCK_M);
  // ---- GEMM 1: hidden = relu(x W1^T + b1) ----
  {
    CUtensorMap ma, mb;
    const int bn = (H % 128 == 0) ? 128 : 64;
    if (int e = operand_map(&ma, x16, D, (int)R, C, BLOCK_M)) return e;
    if (int e = operand_map(&mb, w1_16, D, H, C, bn)) return e;
    GemmParams p{};
    p.epi = EPI_BIAS_RELU_BF16;
    p.num_kblocks = D / BLOCK_K;
    p.block_n = bn;
    p.idesc = make_idesc_f16(BLOCK_M, bn, (uint32_t)dtype16);
    p.hidden_bf16 = dtype16 == KWS_BF16;
    p.m_tiles = m_tiles;
    p.n_tiles = H / bn;
    p.num_items = (long long)C * m_tiles * p.n_tiles;
    p.rows = (int)R;
    p.cols = H;
    p.out = hidden16;
    p.bias = b1;
    if (int e = launch_gemm(ma, mb, p, st)) return e;
  }
  // ---- GEMM 2: out = hidden W2^T + b2 (+ normalise * mask) ----
  {
    CUtensorMap ma, mb;
    if (int e = operand_map(&ma, hidden16, H, (int)R, C, BLOCK_M)) return e;
    if (int e = operand_map(&mb, w2_16, H, P, C, P)) return e;
    GemmParams p{};
    p.epi = EPI_BIAS_OUT;
    p.num_kblocks = H / BLOCK_K;
    p.block_n = P;
    p.idesc = make_idesc_f16(BLOCK_M, P, (uint32_t)dtype16);
    p.m_tiles = m_tiles;
    p.n_tiles = 1;
    p.num_items = (long long)C * m_tiles;
    p.rows = (int)R;
    p.cols = P;
    p.out = out;
    p.bias = b2;
    p.mask = mask;
    p.out_mode = out_mode;
    p.hidden_bf16 = dtype16 == KWS_BF16;
    p.T = T;
    p.Cn = C;
    p.eps = eps;
    if (int e = launch_gemm(ma, mb, p, st)) return e;
  }
  return 0;
}

int kws_sim(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk, int pair_mode,
            float* feat_f32, void* feat_f16, int pitch16, void* stream) {
  KWS_CHECK_ARG(kwd_n && utt_n, "sim: null operand");
  KWS_CHECK_ARG(feat_f32 || feat_f16, "sim: no output requested");
  KWS_CHECK_ARG(C > 0 && K > 0 && U > 0 && Tk > 0 && Tu > 0, "sim: non-positive dimension");
  KWS_CHECK_ARG(Dk % 64 == 0 && Dk >= 64, "sim: Dk=%d must be a multiple of 64", Dk);
  KWS_CHECK_ARG(Tk <= 256, "sim: Tk=%d > 256 keyword frames not supported", Tk);
  KWS_CHECK_ARG(pair_mode == KWS_PAIRS_ALL || pair_mode == KWS_PAIRS_DIAG, "sim: bad pair_mode %d", pair_mode);
  KWS_CHECK_ARG(pair_mode == KWS_PAIRS_ALL || U == K, "sim: KWS_PAIRS_DIAG needs U == K (got K=%d U=%d)", K, U);
  KWS_CHECK_ARG(!feat_f16 || pitch16 >= Tu, "sim: pitch16=%d < Tu=%d", pitch16, Tu);
  const int bn = (Tk + 15) & ~15;
  CUtensorMap ma, mb;
  if (int e = operand_map(&ma, utt_n, Dk, Tu, (long long)C * U, BLOCK_M)) return e;
  if (int e = operand_map(&mb, kwd_n, Dk, Tk, (long long)C * K, bn)) return e;
  GemmParams p{};
  p.epi = EPI_SIM;
  p.num_kblocks = Dk / BLOCK_K;
  p.block_n = bn;
  p.idesc = make_idesc_f16(BLOCK_M, bn, 0);
  p.m_tiles = (Tu + BLOCK_M - 1) / BLOCK_M;
  p.n_tiles = 1;
  p.diag = pair_mode == KWS_PAIRS_DIAG;
  p.num_items = (long long)K * (p.diag ? 1 : U) * C * p.m_tiles;
  p.rows = Tu;
  p.cols = Tk;
  p.out = feat_f32;
  p.out2 = feat_f16;
  p.K = K, p.U = U, p.C = C, p.Tk = Tk, p.Tu = Tu, p.pitch16 = pitch16;
  return launch_gemm(ma, mb, p, (cudaStream_t)stream);
}

int kws_sim_operand(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk, void* out_f16,
                    void* stream) {
  KWS_CHECK_ARG(kwd_n && utt_n && out_f16, "sim_operand: null pointer");
  KWS_CHECK_ARG(C > 0 && K > 0 && U > 0 && Tk > 0 && Tu > 0, "sim_operand: non-positive dimension");
  KWS_CHECK_ARG(Dk % 64 == 0 && Dk >= 64, "sim_operand: Dk=%d must be a multiple of 64", Dk);
  KWS_CHECK_ARG(Tk % 16 == 0 && Tk <= 256, "sim_operand: Tk=%d must be a multiple of 16, <= 256 (zero-pad the bank)", Tk);
  KWS_CHECK_ARG((reinterpret_cast<uintptr_t>(out_f16) & 15) == 0, "sim_operand: out must be 16-byte aligned");
  CUtensorMap ma, mb;
  if (int e = operand_map(&ma, utt_n, Dk, Tu, (long long)C * U, BLOCK_M)) return e;
  if (int e = operand_map(&mb, kwd_n, Dk, Tk, (long long)C * K, Tk)) return e;
  GemmParams p{};
  p.epi = EPI_SIM;
  p.operand_out = 1;
  p.num_kblocks = Dk / BLOCK_K;
  p.block_n = Tk;
  p.idesc = make_idesc_f16(BLOCK_M, Tk, 0);
  p.m_tiles = (Tu + BLOCK_M - 1) / BLOCK_M;
  p.n_tiles = 1;
  p.num_items = (long long)K * U * C * p.m_tiles;
  p.rows = Tu;
  p.cols = Tk;
  p.out2 = out_f16;
  p.K = K, p.U = U, p.C = C, p.Tk = Tk, p.Tu = Tu;
  return launch_gemm(ma, mb, p, (cudaStream_t)stream);
}

}  // extern "C"
