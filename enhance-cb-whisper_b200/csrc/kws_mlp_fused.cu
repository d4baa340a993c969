// Fused per-layer projector (sm_100a): projector[c] = Linear(D, H) -> ReLU -> Linear(H, P) of the reference
// (src/efficient_kws/model.py:92-104, applied :146-150), layer selection, fp32 -> 16-bit cast, both GEMMs and the
// L2-normalise * mask epilogue (model.py:214-216, :187-191) in ONE persistent kernel.  The [C, R, H] hidden activation
// never exists in HBM and the raw fp32 embeddings are read exactly once (H <= 384) by the kernel that consumes them.
//
// Work item = (layer c, tile of 128 rows).  Per item and per hidden pass (NH <= 384 hidden units: all of H for
// D <= 768, two passes for D = 1024 / 1280):
//   loaders (8 warps)  : fp32 rows of x [B, Cin, T, D] straight from HBM with 128-bit streaming loads (16 lanes = one
//                        contiguous 256-byte run of a row), converted to fp16 / bf16 in registers and stored as the
//                        128B-swizzled K-major A operand (128 rows x 64) of a pipeline stage;
//   TMA producer       : the matching W1 k-block (NH x 64) into the same stage; W2 (P x H) once per layer, resident;
//   MMA issuer         : GEMM1  hidden[128, NH] += x_tile . W1_kblock^T   (tcgen05, fp32 accumulators in TMEM columns
//                        [0, NH), N split into two instructions when NH > 256), then, as the epilogue hands hidden
//                        tiles back, GEMM2  out[128, P] += hidden_tile_j . W2_j^T  into TMEM columns [448, 448 + P);
//   epilogue (4 warps) : drains the hidden accumulator 64 columns at a time: + b1, ReLU, 16-bit, into a 128B-swizzled
//                        shared-memory tile that is the A operand of GEMM2 (two tiles, ping-pong); after the last
//                        pass: + b2, row norm, * mask / max(norm, eps) -> fp16 (LE), or raw 16-bit / fp32 (LEF, parity).
//
// Shared-memory port budget per k-block at D = 768 (NH = 384): MMA operand reads 80 KB + W1 TMA writes 48 KB + x
// stores 16 KB = 144 KB = 1125 cycles against 768 cycles of tensor math: the port, not the tensor pipe, bounds a
// 1-CTA design (2-CTA MMAs that split W1 across a CTA pair are the next step).
#include "kws_common.cuh"
#include "../../include/kws_b200.h"

namespace kws {

constexpr int MF_THREADS = 512;          // 16 warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue, 8-15 loaders
constexpr int MF_LOADER_WARPS = 8;
constexpr int MF_BM = 128;
constexpr int MF_BK = 64;
constexpr int MF_A_BYTES = MF_BM * MF_BK * 2;   // 16 KB: x tile of a stage / one hidden tile
constexpr int MF_TMEM_OUT = 448;                // GEMM2 accumulator columns [448, 448 + P), P <= 64
constexpr int MF_MAX_NH = 384;
constexpr int MF_MAX_LAYERS = 64;

struct MlpFusedParams {
  const float* x;       // [B, Cin, T, D] fp32
  int B, Cin, T, D, C, H, P;
  int32_t lidx[MF_MAX_LAYERS];
  long long R;          // rows per layer = B * T
  int m_tiles;
  int nkb;              // D / 64
  int NH, npass;        // hidden units per pass, passes
  int nsplit;           // GEMM1 instructions per k16 step: 1 (NH <= 256) or 2 (NH / 2 each)
  int stages;           // pipeline stages (2 or 3)
  int bf16;             // operand type: 0 fp16 | 1 bf16
  const float* b1;      // [C, H]
  const float* b2;      // [C, P]
  const float* mask;    // [B, C, T] or null (NORM mode)
  float eps;
  int out_mode;         // KWS_MLP_OUT_*
  void* out;            // [C, R, P]
  long long num_items;
};

__device__ __forceinline__ float4 mf_ldg(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void mf_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(MF_THREADS, 1)
kws_mlp_fused_kernel(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
                     const MlpFusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int stage_bytes = MF_A_BYTES + p.NH * 128;           // x tile | W1 k-block (NH rows x 128 B)
  uint8_t* s_stage = smem_raw;                                // [stages][stage_bytes], each 1024-aligned
  uint8_t* s_hid = s_stage + p.stages * stage_bytes;          // [2][16 KB] hidden tiles (A operand of GEMM2)
  uint8_t* s_w2 = s_hid + 2 * MF_A_BYTES;                     // [H / 64][P x 128 B] resident W2 of the current layer
  float* s_b1 = reinterpret_cast<float*>(s_w2 + (p.H / 64) * p.P * 128);  // [H]
  float* s_b2 = s_b1 + p.H;                                   // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b2 + 64);
  uint64_t* full = bars;             // [4] loaders (8 warp arrivals) + TMA tx -> MMA
  uint64_t* empty = full + 4;        // [4] MMA commit -> loaders, TMA
  uint64_t* hacc_full = empty + 4;   // GEMM1 of a pass complete -> epilogue
  uint64_t* hacc_empty = hacc_full + 1;  // epilogue has drained the hidden accumulator (4 warp arrivals) -> MMA
  uint64_t* hfull = hacc_empty + 1;  // [2] hidden tile written (4 warp arrivals) -> MMA
  uint64_t* hempty = hfull + 2;      // [2] GEMM2 on the tile complete (commit) -> epilogue
  uint64_t* out_full = hempty + 2;   // GEMM2 of an item complete -> epilogue
  uint64_t* out_empty = out_full + 1;  // epilogue has read the output accumulator (4 warp arrivals) -> MMA
  uint64_t* w2_full = out_empty + 1;   // W2 of the item's layer resident (TMA tx) -> MMA
  uint64_t* w2_free = w2_full + 1;     // all GEMM2 MMAs of a layer's items complete (commit) -> TMA may overwrite W2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w2_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if ((smem_u32(smem_raw) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[kws] mlp_fused: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w1);
    tma_prefetch_desc(&map_w2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(&full[s], MF_LOADER_WARPS + 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(hacc_full, 1);
    mbar_init(hacc_empty, 4);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&hfull[s], 4);
      mbar_init(&hempty[s], 1);
    }
    mbar_init(out_full, 1);
    mbar_init(out_empty, 4);
    mbar_init(w2_full, 1);
    mbar_init(w2_free, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nhb = p.NH / 64;  // hidden 64-column blocks per pass

  if (warp == 0) {
    // ===================== TMA producer: W1 k-blocks per stage, W2 per layer change =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, nchg = 0;  // nchg: W2 loads issued = layer changes seen by this CTA
      int prev_c = -1;
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const int c = (int)(it / p.m_tiles);
        if (c != prev_c) {
          // w2_free completes one phase per layer change (committed by the MMA issuer after the last item of a layer),
          // so producer and issuer stay in lock-step on it whatever the run-ahead of the W1 pipeline
          if (nchg > 0) mbar_wait(w2_free, (nchg - 1) & 1, 900);  // the previous layer's GEMM2s no longer read W2
          ++nchg;
          mbar_arrive_expect_tx(w2_full, (uint32_t)((p.H / 64) * p.P * 128));
          for (int j = 0; j < p.H / 64; ++j) tma_load_3d(&map_w2, w2_full, s_w2 + j * p.P * 128, j * 64, 0, c);
          prev_c = c;
        }
        for (int ps = 0; ps < p.npass; ++ps) {
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1, 100 + stage);
            uint8_t* sb = s_stage + stage * stage_bytes + MF_A_BYTES;
            mbar_arrive_expect_tx(&full[stage], (uint32_t)(p.NH * 128));
            const int half = p.NH / p.nsplit;  // TMA boxes of <= 256 rows
            for (int h = 0; h < p.nsplit; ++h)
              tma_load_3d(&map_w1, &full[stage], sb + h * half * 128, kb * 64, ps * p.NH + h * half, c);
            if (++stage == p.stages) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const int half = p.NH / p.nsplit;
      const uint32_t idesc1 = make_idesc_f16(MF_BM, (uint32_t)half, (uint32_t)p.bf16);
      const uint32_t idesc2 = make_idesc_f16(MF_BM, (uint32_t)p.P, (uint32_t)p.bf16);
      int stage = 0;
      uint32_t phase = 0, seq = 0, hseq = 0, pseq = 0;  // hseq: hidden tiles consumed, pseq: passes issued
      int prev_c = -1;
      uint32_t w2_loads = 0;
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x, ++seq) {
        const int c = (int)(it / p.m_tiles);
        for (int ps = 0; ps < p.npass; ++ps, ++pseq) {
          // the epilogue has drained the hidden accumulator of the previous pass
          mbar_wait(hacc_empty, (pseq & 1) ^ 1, 200);
          tc_fence_after();
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&full[stage], phase, 300 + stage);
            tc_fence_after();
            const uint32_t sa = smem_u32(s_stage + stage * stage_bytes);
            const uint64_t adesc = make_smem_desc(sa, 16, 1024, LAYOUT_SW128);
            const uint64_t bdesc = make_smem_desc(sa + MF_A_BYTES, 16, 1024, LAYOUT_SW128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc1, (kb | k) != 0);
              if (p.nsplit == 2)
                umma_f16(tmem_base + (uint32_t)half, adesc + (uint64_t)(k * 2),
                         bdesc + (uint64_t)((half * 128) >> 4) + (uint64_t)(k * 2), idesc1, (kb | k) != 0);
            }
            umma_commit(&empty[stage]);
            if (++stage == p.stages) stage = 0, phase ^= 1;
          }
          umma_commit(hacc_full);
          // GEMM2 on the hidden tiles of this pass, as the epilogue produces them
          for (int j = 0; j < nhb; ++j, ++hseq) {
            if (ps == 0 && j == 0) {
              if (c != prev_c) {
                mbar_wait(w2_full, w2_loads & 1, 910);
                ++w2_loads;
                prev_c = c;
              }
              mbar_wait(out_empty, (seq & 1) ^ 1, 210);  // the previous item's output accumulator has been read
            }
            mbar_wait(&hfull[hseq & 1], (hseq >> 1) & 1, 400 + (int)(hseq & 1));
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(smem_u32(s_hid + (hseq & 1) * MF_A_BYTES), 16, 1024, LAYOUT_SW128);
            const uint64_t bdesc = make_smem_desc(smem_u32(s_w2 + (ps * nhb + j) * p.P * 128), 16, 1024, LAYOUT_SW128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + MF_TMEM_OUT, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc2,
                       (ps | j | k) != 0);
            umma_commit(&hempty[hseq & 1]);
          }
        }
        umma_commit(out_full);
        const long long nit = it + gridDim.x;
        if (nit >= p.num_items || (int)(nit / p.m_tiles) != c) umma_commit(w2_free);  // last item of this layer here
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: hidden drain (bias + ReLU -> 16-bit A tiles) and the output =====================
    const int q = warp & 3;
    const int rl = q * 32 + lane;  // row of the tile = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t seq = 0, hseq = 0, pseq = 0;
    int prev_c = -1;
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x, ++seq) {
      const int c = (int)(it / p.m_tiles);
      const int mt = (int)(it - (long long)c * p.m_tiles);
      if (c != prev_c) {  // biases of the layer -> shared memory (no L1 with the full carve-out)
        mf_bar_sync(2, 128);  // every epilogue thread is done with the previous layer's biases
        for (int i = rl; i < p.H; i += 128) s_b1[i] = __ldg(p.b1 + (long long)c * p.H + i);
        if (rl < p.P) s_b2[rl] = __ldg(p.b2 + (long long)c * p.P + rl);
        mf_bar_sync(2, 128);
        prev_c = c;
      }
      for (int ps = 0; ps < p.npass; ++ps, ++pseq) {
        mbar_wait(hacc_full, pseq & 1, 500);
        tc_fence_after();
        for (int j = 0; j < nhb; ++j, ++hseq) {
          uint8_t* srow = s_hid + (hseq & 1) * MF_A_BYTES + rl * 128;
          const float* bj = s_b1 + ps * p.NH + j * 64;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {  // 32 columns per TMEM round trip (register budget: 128 per thread)
            uint32_t v[2][16];
            tmem_ld16(t_lane + j * 64 + hh * 32, v[0]);
            tmem_ld16(t_lane + j * 64 + hh * 32 + 16, v[1]);
            // the GEMM2 MMAs that read this tile buffer two tiles ago have retired
            if (hh == 0) mbar_wait(&hempty[hseq & 1], ((hseq >> 1) & 1) ^ 1, 600 + (int)(hseq & 1));
            tmem_ld_wait();
            if (hh == 1 && j == nhb - 1) {  // last read of the hidden accumulator of this pass
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(hacc_empty);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float a = fmaxf(__uint_as_float(v[h][2 * e]) + bj[hh * 32 + h * 16 + 2 * e], 0.f);
                const float b = fmaxf(__uint_as_float(v[h][2 * e + 1]) + bj[hh * 32 + h * 16 + 2 * e + 1], 0.f);
                pk[e] = p.bf16 ? pack_bf162(a, b) : pack_half2_sat(a, b);
              }
              const int c16 = hh * 4 + 2 * h;
              *reinterpret_cast<uint4*>(srow + (((c16) ^ (rl & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(srow + (((c16 + 1) ^ (rl & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&hfull[hseq & 1]);
        }
      }
      // ---- output of the item: + b2, (normalise * mask), store; two passes over TMEM, 16 columns at a time ----
      mbar_wait(out_full, seq & 1, 700);
      tc_fence_after();
      const long long row = (long long)mt * MF_BM + rl;
      const bool row_ok = row < p.R;
      const int nch = p.P >> 4;
      const bool norm = p.out_mode == KWS_MLP_OUT_NORM_F16;
      float scale = 1.f;
      if (norm) {
        float ss = 0.f;
        for (int h = 0; h < nch; ++h) {
          uint32_t v[16];
          tmem_ld16(t_lane + MF_TMEM_OUT + h * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float a = __uint_as_float(v[e]) + s_b2[h * 16 + e];
            ss = fmaf(a, a, ss);
          }
        }
        float m = 1.f;
        if (p.mask && row_ok) {
          const long long b = row / p.T;
          const int t = (int)(row - b * p.T);
          m = p.mask[(b * p.C + c) * p.T + t];
        }
        scale = m / fmaxf(sqrtf(ss), p.eps);
      }
      for (int h = 0; h < nch; ++h) {
        uint32_t v[16];
        tmem_ld16(t_lane + MF_TMEM_OUT + h * 16, v);
        tmem_ld_wait();
        if (h == nch - 1) {  // last read of the output accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(out_empty);
        }
        if (!row_ok) continue;
        if (p.out_mode == KWS_MLP_OUT_RAW_F32) {
          float* o = reinterpret_cast<float*>(p.out) + ((long long)c * p.R + row) * p.P + h * 16;
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            *reinterpret_cast<float4*>(o + e) =
                make_float4(__uint_as_float(v[e]) + s_b2[h * 16 + e], __uint_as_float(v[e + 1]) + s_b2[h * 16 + e + 1],
                            __uint_as_float(v[e + 2]) + s_b2[h * 16 + e + 2], __uint_as_float(v[e + 3]) + s_b2[h * 16 + e + 3]);
        } else {
          uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + ((long long)c * p.R + row) * p.P + h * 16;
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float a = (__uint_as_float(v[2 * e]) + s_b2[h * 16 + 2 * e]) * scale;
            const float b = (__uint_as_float(v[2 * e + 1]) + s_b2[h * 16 + 2 * e + 1]) * scale;
            pk[e] = norm ? pack_half2(a, b) : (p.bf16 ? pack_bf162(a, b) : pack_half2_sat(a, b));
          }
          *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== loaders: fp32 x rows -> 16-bit swizzled A tiles =====================
    // warp lw owns rows [16 lw, 16 lw + 16) of the tile; per k-block a lane loads 8 float4: row 2 i + (lane >> 4),
    // columns 4 (lane & 15) .. +3 -- 16 lanes read one contiguous 256-byte run.  Two register sets: the loads of
    // k-block n+1 are in flight while k-block n is converted and stored.
    const int lw = warp - 8;
    const int sub = lane >> 4, c4 = lane & 15;
    int stage = 0;
    uint32_t phase = 0;
    // flattened (item, pass, k-block) sequence of this CTA
    long long it_l = blockIdx.x;
    int ps_l = 0, kb_l = 0;
    bool have = it_l < p.num_items;
    const float4* xbase = reinterpret_cast<const float4*>(p.x) + c4;
    uint32_t rowo[8];  // row start in units of 64 floats (rows are D floats, D % 64 == 0); ~0u = beyond the last row
    auto setup_rows = [&](long long it) {
      const int c = (int)(it / p.m_tiles);
      const int mt = (int)(it - (long long)c * p.m_tiles);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long r = (long long)mt * MF_BM + lw * 16 + 2 * i + sub;
        if (r < p.R) {
          const long long b = r / p.T;
          const int t = (int)(r - b * p.T);
          rowo[i] = (uint32_t)((((b * p.Cin + p.lidx[c]) * p.T + t) * (long long)p.D) >> 6);
        } else {
          rowo[i] = ~0u;
        }
      }
    };
    auto issue = [&](float4 (&r)[8], int kb) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r[i] = rowo[i] != ~0u ? mf_ldg(xbase + ((size_t)rowo[i] + kb) * 16)  // 64 floats = 16 float4
                              : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto advance = [&]() {  // next (item, pass, k-block); refreshes the row pointers on an item change
      if (++kb_l == p.nkb) {
        kb_l = 0;
        if (++ps_l == p.npass) {
          ps_l = 0;
          it_l += gridDim.x;
          have = it_l < p.num_items;
          if (have) setup_rows(it_l);
        }
      }
    };
    auto store = [&](const float4 (&r)[8]) {
      mbar_wait(&empty[stage], phase ^ 1, 800 + stage);
      uint8_t* sa = s_stage + stage * stage_bytes;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = lw * 16 + 2 * i + sub;
        uint2 o;
        if (p.bf16) {
          o.x = pack_bf162(r[i].x, r[i].y), o.y = pack_bf162(r[i].z, r[i].w);
        } else {
          o.x = pack_half2_sat(r[i].x, r[i].y), o.y = pack_half2_sat(r[i].z, r[i].w);
        }
        *reinterpret_cast<uint2*>(sa + row * 128 + (((c4 >> 1) ^ (row & 7)) << 4) + (c4 & 1) * 8) = o;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);
      if (++stage == p.stages) stage = 0, phase ^= 1;
    };
    float4 ra[8], rb[8];
    if (have) {
      setup_rows(it_l);
      issue(ra, kb_l);
      advance();
    }
    // invariant at loop top: set A holds an issued k-block; `have` says whether another one follows
    bool have_a = blockIdx.x < p.num_items;
    while (have_a) {
      bool have_b = have;
      if (have_b) {
        issue(rb, kb_l);
        advance();
      }
      store(ra);
      if (!have_b) break;
      have_a = have;
      if (have_a) {
        issue(ra, kb_l);
        advance();
      }
      store(rb);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static size_t mf_smem_bytes(int NH, int stages, int H, int P) {
  return (size_t)stages * (MF_A_BYTES + NH * 128) + 2 * MF_A_BYTES + (size_t)(H / 64) * P * 128 + (size_t)(H + 64) * 4 +
         24 * 8 + 16;
}

// hidden units per pass: the largest NH <= 384 with NH % 64 == 0 that divides H into <= 4 equal passes
static bool mf_plan(int D, int H, int P, int* NH, int* npass, int* stages) {
  if (D % 64 != 0 || D < 64 || H % 64 != 0 || H < 64 || P % 16 != 0 || P < 16 || P > 64) return false;
  for (int n = 1; n <= 4; ++n) {
    if (H % n != 0) continue;
    const int nh = H / n;
    if (nh > MF_MAX_NH || nh % 64 != 0) continue;
    if (nh > 256 && (nh / 2) % 16 != 0) continue;
    for (int s = 3; s >= 2; --s) {
      if (mf_smem_bytes(nh, s, H, P) <= 232448) {
        *NH = nh, *npass = n, *stages = s;
        return true;
      }
    }
  }
  return false;
}

}  // namespace kws

using namespace kws;

extern "C" int kws_mlp_fused_supported(int D, int H, int P) {
  int nh, np, st;
  return mf_plan(D, H, P, &nh, &np, &st) ? 1 : 0;
}

extern "C" int kws_mlp_fused(const float* x, int B, int Cin, int T, int D, const int32_t* layer_idx, int C, int H, int P,
                             int dtype16, const void* w1_16, const float* b1, const void* w2_16, const float* b2,
                             const float* mask, float eps, int out_mode, void* out, void* stream) {
  KWS_CHECK_ARG(x && layer_idx && w1_16 && b1 && w2_16 && b2 && out, "mlp_fused: null pointer");
  KWS_CHECK_ARG(dtype16 == KWS_F16 || dtype16 == KWS_BF16, "mlp_fused: bad dtype16 %d", dtype16);
  KWS_CHECK_ARG(B > 0 && Cin > 0 && T > 0 && C > 0 && C <= MF_MAX_LAYERS, "mlp_fused: bad dimension");
  KWS_CHECK_ARG(out_mode == KWS_MLP_OUT_NORM_F16 || out_mode == KWS_MLP_OUT_RAW_F32 || out_mode == KWS_MLP_OUT_RAW_16,
                "mlp_fused: bad out_mode %d", out_mode);
  KWS_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "mlp_fused: pointers must be 16-byte aligned");
  MlpFusedParams p{};
  KWS_CHECK_ARG(mf_plan(D, H, P, &p.NH, &p.npass, &p.stages),
                "mlp_fused: shape D=%d H=%d P=%d not covered (use kws_cast_rows16 + kws_mlp)", D, H, P);
  for (int i = 0; i < C; ++i) {
    KWS_CHECK_ARG(layer_idx[i] >= 0 && layer_idx[i] < Cin, "mlp_fused: layer_idx[%d]=%d out of [0,%d)", i, layer_idx[i], Cin);
    p.lidx[i] = layer_idx[i];
  }
  p.x = x, p.B = B, p.Cin = Cin, p.T = T, p.D = D, p.C = C, p.H = H, p.P = P;
  p.R = (long long)B * T;
  KWS_CHECK_ARG(p.R < (1ll << 31), "mlp_fused: B*T too large");
  KWS_CHECK_ARG((((long long)B * Cin * T * D) >> 6) < 0xffffffffll, "mlp_fused: x too large for 32-bit row offsets (1 TB)");
  p.m_tiles = (int)((p.R + MF_BM - 1) / MF_BM);
  p.nkb = D / MF_BK;
  p.nsplit = p.NH > 256 ? 2 : 1;
  p.bf16 = dtype16 == KWS_BF16;
  p.b1 = b1, p.b2 = b2, p.mask = mask, p.eps = eps, p.out_mode = out_mode, p.out = out;
  p.num_items = (long long)C * p.m_tiles;
  CUtensorMap m1, m2;
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)H, (uint64_t)C};
    const uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)D * 2 * (uint64_t)H};
    const uint32_t box[3] = {64, (uint32_t)(p.NH / p.nsplit), 1};
    if (int e = make_tensor_map(&m1, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, w1_16, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return e;
  }
  {
    const uint64_t dims[3] = {(uint64_t)H, (uint64_t)P, (uint64_t)C};
    const uint64_t strides[2] = {(uint64_t)H * 2, (uint64_t)H * 2 * (uint64_t)P};
    const uint32_t box[3] = {64, (uint32_t)P, 1};
    if (int e = make_tensor_map(&m2, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, w2_16, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return e;
  }
  const size_t smem = mf_smem_bytes(p.NH, p.stages, H, P);
  KWS_CUDA(cudaFuncSetAttribute(kws_mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = p.num_items;
  const int sms = sm_count();
  if (grid > sms) grid = sms;
  kws_mlp_fused_kernel<<<(int)grid, MF_THREADS, smem, (cudaStream_t)stream>>>(m1, m2, p);
  KWS_CUDA(cudaGetLastError());
  return 0;
}
