// [v1, superseded by kws_fused.cu -- kept as kws_sim_stem_v1* (not in the public header) for A/B timing and
// cross-checks while the N=128 tap-pair kernel is brought up]
// Fused similarity + ResNet-stem kernel (sm_100a): the layer-wise cosine-similarity 'image'
// of a (keyword, utterance) pair is produced tile by tile in tensor memory, converted to fp16
// straight into the shared-memory operand layout of the stem convolution, and consumed there by
// the tap-decomposed tcgen05 implicit GEMM.  The [pairs, C, Tk, Tu] tensor of the reference
// (model.py:174-191) never exists in HBM.
//
//   S_c[i, j]          = < kwd_n[c, k, i, :], utt_n[c, u, j, :] >                 (model.py:210-218)
//   out[oc, oi, oj]    = relu(bias[oc] + sum_{c,di,dj} W'[oc,c,di,dj] S_c[2oi+di-3, 2oj+dj-3])
//                                                        (HF modeling_resnet.py:39-54, BN folded)
//
// Work item = (pair, column tile of 61 output columns); an item walks down the image in steps of
// two output rows.  One stem MMA covers M = 128 = 2 output rows x 64 pixel slots (61 used),
// N = 64 output channels, K = 16 input channels, for one tap (di, dj); 49 taps accumulate in TMEM.
//
// Shared-memory operand of the stem ("ring"): input rows are kept de-interleaved by column
// parity (plane) and by row parity (rp), channels innermost in chunks of 8:
//   block[k8][rp][plane] : ring of NR row slots x (64 pixels x 16 B)
// For tap (di, dj) the 128 A-rows are 128 consecutive 16-byte pixels starting at
//   slot(r0) * 1024 + (dj >> 1) * 16      in block[.][(di+1)&1][dj&1],   r0 = 4P + di - 3:
// rows 0..63 read input row r0 (output row 2P), rows 64..127 run on into the next slot, which
// holds input row r0 + 2 (output row 2P+1).  A tap only changes the descriptor start address.
// Slot NR mirrors slot 0 so the run never wraps.
//
// The similarity GEMM works on chunks of 16 input rows: for each layer c, D[128 px, 16 rows] =
// utt tile (128 x Dk, TMA, OOB columns zero-filled) x kwd rows (16 x Dk)^T, fp32 in one of two
// TMEM regions; 4 converter warps (thread = pixel) read 4 rows x C layers at a time, pack fp16
// and store 16-byte (8-channel) words into the ring, one quantum (4 rows) per stem step.
//
// Roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer (similarity + stem, similarity
// stages are issued opportunistically between tap groups so the tensor pipe never waits on
// them), warp 2 TMEM allocator, warps 4..7 stem epilogue, warps 8..11 converters.
#include "kws_common.cuh"
#include "../../include/kws_b200.h"

namespace kws {

constexpr int F_THREADS = 384;
constexpr int F_OC = 64;
constexpr int F_TILE_OJ = 61;                      // output columns per item (2*61 + 5 = 127 <= 128 input px)
constexpr int F_NR = 8;                            // ring slots per block = 4 quanta of 2 slots
constexpr int F_BLOCK = (F_NR + 1) * 1024 + 64;    // 9280: +mirror slot, +64 keeps plane 1 on other banks
constexpr int F_RING_BYTES = 8 * F_BLOCK;          // [k8 2][rp 2][plane 2]
constexpr int F_TAP_BYTES = 2 * F_OC * 16;         // 2048: [k8][oc][8 ch] fp16
constexpr int F_W_BYTES = 49 * F_TAP_BYTES;        // 100352
constexpr int F_NS = 3;                            // similarity operand stages
constexpr int F_A_BYTES = 128 * 128;               // utt tile 128 px x 64 dims (SW128)
constexpr int F_B_BYTES = 16 * 128;                // kwd tile 16 rows x 64 dims (SW128)
constexpr int F_STAGE = F_A_BYTES + F_B_BYTES;     // 18432
constexpr int F_MAX_C = 12;
constexpr int F_SIM_COLS = F_MAX_C * 16;           // 192 TMEM columns per similarity region
constexpr int F_TMEM_SIM = 2 * F_OC;               // stem accumulators in columns [0,128)
constexpr int F_NBAR = 2 * F_NS + 2 + 2 + 4 + 4 + 2 + 2;

struct FusedParams {
  const uint4* w;     // packed stem weights (kws_pack_stem_weights, one 16-channel group)
  const float* bias;  // [64]
  void* out;
  int out_mode;
  int C, K, U, Tk, Tu, nkb, Ho, Wo, col_tiles;  // K, U: operand batch sizes (tensor-map extents)
  int k0, u0, nk, nu;  // scored sub-range: keywords [k0, k0+nk) x utterances [u0, u0+nu); out pair = (k-k0)*nu + (u-u0)
  int nP;        // stem steps per item = ceil(Ho / 2)
  int nQ;        // quanta (4 input rows) converted per item = nP + 2
  int n_chunks;  // similarity chunks (16 input rows) per item = ceil(nQ / 4)
  int diag;
  long long num_items;
  long long* dbg;  // optional [grid][8] cycle counters of the MMA issuer (development aid), or null
};

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ bool mbar_poll(uint64_t* bar, uint32_t parity, bool blocking, int tag) {
  if (mbar_try_wait(bar, parity)) return true;
  if (!blocking) return false;
  mbar_wait(bar, parity, tag);
  return true;
}

struct ItemCoord {
  long long pair;
  int ct, kw, u;
};
__device__ __forceinline__ ItemCoord decode_item(const FusedParams& p, long long it) {
  ItemCoord r;
  r.pair = it / p.col_tiles;
  r.ct = (int)(it - r.pair * p.col_tiles);
  if (p.diag) {
    r.kw = r.u = p.k0 + (int)r.pair;
  } else {
    const int kl = (int)(r.pair / p.nu);
    r.kw = p.k0 + kl;
    r.u = p.u0 + (int)(r.pair - (long long)kl * p.nu);
  }
  return r;
}

__global__ void __launch_bounds__(F_THREADS, 1)
kws_fused_v1_kernel(const __grid_constant__ CUtensorMap map_utt, const __grid_constant__ CUtensorMap map_kwd,
                 const FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_ops = base;                          // F_NS * F_STAGE (each 1024-aligned)
  uint8_t* s_w = s_ops + F_NS * F_STAGE;          // F_W_BYTES
  uint8_t* s_ring = s_w + F_W_BYTES;              // F_RING_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + F_RING_BYTES);
  uint64_t* ofull = bars;                 // [F_NS] TMA -> MMA (similarity operands)
  uint64_t* oempty = ofull + F_NS;        // [F_NS] MMA commit -> TMA
  uint64_t* sfull = oempty + F_NS;        // [2] MMA commit -> converters (similarity region ready)
  uint64_t* sempty = sfull + 2;           // [2] converters -> MMA
  uint64_t* qfull = sempty + 2;           // [4] converters -> MMA (ring quantum written)
  uint64_t* qempty = qfull + 4;           // [4] MMA commit -> converters
  uint64_t* afull = qempty + 4;           // [2] MMA commit -> epilogue (stem accumulator ready)
  uint64_t* aempty = afull + 2;           // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + F_NBAR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < F_W_BYTES / 16; i += F_THREADS) reinterpret_cast<uint4*>(s_w)[i] = p.w[i];
  fence_proxy_async();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_utt);
    tma_prefetch_desc(&map_kwd);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < F_NS; ++s) {
      mbar_init(&ofull[s], 1);
      mbar_init(&oempty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sfull[s], 1);
      mbar_init(&sempty[s], 128);
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 128);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&qfull[s], 128);
      mbar_init(&qempty[s], 1);
    }
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
  if (tmem_base != 0) {  // all 512 columns of this SM's TMEM: the allocation can only start at column 0
    if (threadIdx.x == 0) printf("[kws] unexpected TMEM base 0x%x\n", tmem_base);
    __trap();
  }
  const int stages_per_chunk = p.C * p.nkb;

  if (warp == 0) {
    // ===================== TMA producer: similarity operands =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const ItemCoord w = decode_item(p, it);
        const int jbase = 2 * F_TILE_OJ * w.ct - 3;  // input column of pixel x = 0 (OOB columns read as zero)
        for (int n = 0; n < p.n_chunks; ++n) {
          for (int c = 0; c < p.C; ++c) {
            for (int kb = 0; kb < p.nkb; ++kb) {
              mbar_wait(&oempty[stage], phase ^ 1, 100 + stage);
              uint8_t* sa = s_ops + stage * F_STAGE;
              mbar_arrive_expect_tx(&ofull[stage], F_STAGE);
              tma_load_3d(&map_utt, &ofull[stage], sa, kb * 64, jbase, c * p.U + w.u);
              tma_load_3d(&map_kwd, &ofull[stage], sa + F_A_BYTES, kb * 64, 16 * n - 3, c * p.K + w.kw);
              if (++stage == F_NS) stage = 0, phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop with identical (warp-uniform) values; only lane 0 executes the
    // tcgen05 instructions (predicated inside the asm), so descriptor arithmetic stays in uniform
    // registers and one MMA costs a handful of issue slots.
    if (elect_one()) {
    const uint32_t idesc_sim = make_idesc_f16(128, 16, 0);
    const uint32_t idesc_stem = make_idesc_f16(128, F_OC, 0);
    const uint32_t ops_u32 = smem_u32(s_ops);
    const uint64_t adesc0 = make_smem_desc(smem_u32(s_ring), 4 * F_BLOCK, 128, LAYOUT_NONE);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(s_w), F_OC * 16, 128, LAYOUT_NONE);
    const uint64_t sdesc0 = make_smem_desc(0, 16, 1024, LAYOUT_SW128);
    // similarity cursor: runs ahead of the stem cursor (across items) by about one chunk
    long long s_it = blockIdx.x;
    int s_chunk = 0, s_stage = 0;
    uint32_t s_g = 0;  // similarity chunks fully issued so far (global)
    int o_stage = 0;
    uint32_t o_phase = 0;
    // one operand stage = 4 MMAs of one layer / k-block; non-blocking calls return false when the
    // TMEM region or the operands are not there yet (the stem MMAs go on, the call is retried)
    auto sim_issue = [&](bool blocking) -> bool {
      const uint32_t buf = s_g & 1;
      if (s_stage == 0 && !mbar_poll(&sempty[buf], ((s_g >> 1) & 1) ^ 1, blocking, 500 + buf)) return false;
      if (!mbar_poll(&ofull[o_stage], o_phase, blocking, 300 + o_stage)) return false;
      tc_fence_after();
      const int c = s_stage / p.nkb, kb = s_stage - c * p.nkb;
      const uint32_t d = F_TMEM_SIM + buf * F_SIM_COLS + c * 16;  // TMEM base is 0 (checked at start)
      const uint32_t sa = ops_u32 + o_stage * F_STAGE;
      const uint64_t adesc = sdesc0 + (uint64_t)(sa >> 4);
      const uint64_t bdesc = adesc + (uint64_t)(F_A_BYTES >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_sim, (kb | k) != 0);
      umma_commit(&oempty[o_stage]);
      if (++o_stage == F_NS) o_stage = 0, o_phase ^= 1;
      if (++s_stage == stages_per_chunk) {
        umma_commit(&sfull[buf]);
        s_stage = 0;
        ++s_g;
        if (++s_chunk == p.n_chunks) s_chunk = 0, s_it += gridDim.x;
      }
      return true;
    };

    uint32_t g_chunk0 = 0;  // global index of the current item's chunk 0
    uint32_t qbase = 0;     // global index of the current item's quantum 0
    uint32_t acc_seq = 0;   // global stem step counter -> accumulator buffer
    long long tm_sim = 0, tm_acc = 0, tm_q = 0, tm_issue = 0;
    const long long tm_start = clock64();
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      int waited = 0;  // quanta of this item known to be in the ring
      for (int P = 0; P < p.nP; ++P, ++acc_seq) {
        // chunks: `need` holds quantum P+2 and must be issued now; `ahead` (one chunk further,
        // possibly chunk 0 of the next item) is issued one stage per tap group while the stem runs
        int need = (P + 2) >> 2;
        if (need > p.n_chunks - 1) need = p.n_chunks - 1;
        int ahead = ((P + 2) >> 2) + 1;
        if (ahead > p.n_chunks) ahead = p.n_chunks;
        const uint32_t g_need = g_chunk0 + (uint32_t)need, g_ahead = g_chunk0 + (uint32_t)ahead;
        const long long t0 = clock64();
        while (s_g <= g_need && s_it < p.num_items) sim_issue(true);
        const long long t1 = clock64();
        const uint32_t acc = acc_seq & 1;
        mbar_wait(&aempty[acc], ((acc_seq >> 1) & 1) ^ 1, 200 + acc);
        const long long t2 = clock64();
        while (waited <= P + 2 && waited < p.nQ) {
          const uint32_t G = qbase + waited;
          mbar_wait(&qfull[G & 3], (G >> 2) & 1, 400 + (int)(G & 3));
          ++waited;
        }
        const long long t3 = clock64();
        tm_sim += t1 - t0, tm_acc += t2 - t1, tm_q += t3 - t2;
        tc_fence_after();
        const uint32_t d = acc * F_OC;
        const uint32_t slot_base = 2 * (qbase + P);
#pragma unroll
        for (int di = 0; di < 7; ++di) {
          const uint32_t slot0 = (slot_base + (di >> 1)) & (F_NR - 1);
          const uint64_t a_row = adesc0 + (uint64_t)(((((di + 1) & 1) * 2 * F_BLOCK) >> 4) + slot0 * 64);
          const uint64_t b_row = bdesc0 + (uint64_t)((di * 7 * F_TAP_BYTES) >> 4);
#pragma unroll
          for (int dj = 0; dj < 7; ++dj) {
            umma_f16(d, a_row + (uint64_t)((((dj & 1) * F_BLOCK) + (dj >> 1) * 16) >> 4),
                       b_row + (uint64_t)((dj * F_TAP_BYTES) >> 4), idesc_stem, (di | dj) != 0);
          }
          if (s_g <= g_ahead && s_it < p.num_items) sim_issue(false);
        }
        umma_commit(&qempty[(qbase + P) & 3]);  // quantum P is dead once these MMAs retire
        umma_commit(&afull[acc]);
        tm_issue += clock64() - t3;
      }
      // the two tail quanta were read by the last step only
      for (int q = p.nP; q < p.nQ; ++q) umma_commit(&qempty[(qbase + q) & 3]);
      qbase += p.nQ;
      g_chunk0 += p.n_chunks;
    }
    if (p.dbg) {
      long long* o = p.dbg + (size_t)blockIdx.x * 8;
      o[0] = clock64() - tm_start, o[1] = tm_sim, o[2] = tm_acc, o[3] = tm_q, o[4] = tm_issue;
    }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================== stem epilogue =====================
    const int q = warp & 3;
    const int row_sel = q >> 1;
    const int ojl = (q & 1) * 32 + lane;
    float bias_r[F_OC];
#pragma unroll
    for (int i = 0; i < F_OC; ++i) bias_r[i] = __ldg(p.bias + i);
    uint32_t acc_seq = 0;
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const ItemCoord w = decode_item(p, it);
      const int oj = w.ct * F_TILE_OJ + ojl;
      const bool col_ok = ojl < F_TILE_OJ && oj < p.Wo;
      for (int P = 0; P < p.nP; ++P, ++acc_seq) {
        const uint32_t acc = acc_seq & 1;
        mbar_wait(&afull[acc], (acc_seq >> 1) & 1, 600 + acc);
        tc_fence_after();
        const uint32_t t_row = tmem_base + acc * F_OC + ((uint32_t)(q * 32) << 16);
        uint32_t v[4][16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) tmem_ld16(t_row + ch * 16, v[ch]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&aempty[acc]);
        const int oi = 2 * P + row_sel;
        if (col_ok && oi < p.Ho) {
          if (p.out_mode == KWS_STEM_OUT_NCHW_F32) {
            float* o = reinterpret_cast<float*>(p.out) + ((w.pair * F_OC) * p.Ho + oi) * (long long)p.Wo + oj;
            const long long oc_stride = (long long)p.Ho * p.Wo;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
#pragma unroll
              for (int e = 0; e < 16; ++e)
                o[(ch * 16 + e) * oc_stride] = fmaxf(__uint_as_float(v[ch][e]) + bias_r[ch * 16 + e], 0.f);
          } else {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                                ((w.pair * p.Ho + oi) * (long long)p.Wo + oj) * F_OC);
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
  } else if (warp >= 8) {
    // ===================== converters: TMEM similarity rows -> fp16 ring =====================
    const int q4 = warp & 3;
    const int x = q4 * 32 + lane;  // pixel of the 128-wide input window
    uint8_t* dst_px = s_ring + (x & 1) * F_BLOCK + (x >> 1) * 16;
    const uint32_t t_lane = tmem_base + F_TMEM_SIM + ((uint32_t)(q4 * 32) << 16);
    uint32_t g = 0;   // global similarity chunk counter
    uint32_t Gq = 0;  // global quantum counter
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      for (int q = 0; q < p.nQ; ++q, ++Gq) {
        const int qq = q & 3;
        const uint32_t buf = g & 1;
        if (qq == 0) {
          mbar_wait(&sfull[buf], (g >> 1) & 1, 700 + buf);
          tc_fence_after();
        }
        uint32_t v[F_MAX_C][4];
#pragma unroll
        for (int c = 0; c < F_MAX_C; ++c) {
          if (c < p.C) {
            tmem_ld4(t_lane + buf * F_SIM_COLS + c * 16 + qq * 4, v[c]);
          } else {
            v[c][0] = v[c][1] = v[c][2] = v[c][3] = 0u;  // +0.0f
          }
        }
        tmem_ld_wait();
        if (qq == 3 || q == p.nQ - 1) {  // last quantum read from this region
          tc_fence_before();
          mbar_arrive(&sempty[buf]);
          ++g;
        }
        mbar_wait(&qempty[Gq & 3], ((Gq >> 2) & 1) ^ 1, 800 + (int)(Gq & 3));
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          // input row r = 4q + t - 3: parity (t+1)&1, ring slot (2 Gq + (t >> 1)) mod NR
          const uint32_t slot = (2 * Gq + (t >> 1)) & (F_NR - 1);
          uint8_t* d0 = dst_px + ((t + 1) & 1) * 2 * F_BLOCK + slot * 1024;
          const uint4 lo = make_uint4(pack_half2(__uint_as_float(v[0][t]), __uint_as_float(v[1][t])),
                                      pack_half2(__uint_as_float(v[2][t]), __uint_as_float(v[3][t])),
                                      pack_half2(__uint_as_float(v[4][t]), __uint_as_float(v[5][t])),
                                      pack_half2(__uint_as_float(v[6][t]), __uint_as_float(v[7][t])));
          const uint4 hi = make_uint4(pack_half2(__uint_as_float(v[8][t]), __uint_as_float(v[9][t])),
                                      pack_half2(__uint_as_float(v[10][t]), __uint_as_float(v[11][t])), 0u, 0u);
          *reinterpret_cast<uint4*>(d0) = lo;
          *reinterpret_cast<uint4*>(d0 + 4 * F_BLOCK) = hi;
          if (slot == 0) {  // mirror of slot 0 after the last slot: tap windows never wrap
            *reinterpret_cast<uint4*>(d0 + F_NR * 1024) = lo;
            *reinterpret_cast<uint4*>(d0 + 4 * F_BLOCK + F_NR * 1024) = hi;
          }
        }
        fence_proxy_async();
        mbar_arrive(&qfull[Gq & 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

constexpr size_t F_SMEM = 1024 + (size_t)F_NS * F_STAGE + F_W_BYTES + F_RING_BYTES + F_NBAR * 8 + 16;
static_assert(F_SMEM <= 232448, "fused kernel exceeds the 227 KB shared-memory limit");

}  // namespace kws

using namespace kws;

static long long* g_fused_v1_dbg = nullptr;
// development aid (not part of the public header): device buffer [148][8] receiving the issuer's cycle counters
extern "C" void kws_debug_set_fused_v1_counters(long long* dev_buf) { g_fused_v1_dbg = dev_buf; }

extern "C" int kws_sim_stem_v1_supported(int C, int Tk, int Tu, int Dk) {
  return C > 0 && C <= F_MAX_C && Dk >= 64 && Dk % 64 == 0 && Tk > 0 && Tu > 0;
}

extern "C" int kws_sim_stem_v1_range(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk,
                                     int pair_mode, int k0, int nk, int u0, int nu, const void* w_packed,
                                     const float* bias, int out_mode, void* out, void* stream);

extern "C" int kws_sim_stem_v1(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk,
                            int pair_mode, const void* w_packed, const float* bias, int out_mode, void* out,
                            void* stream) {
  return kws_sim_stem_v1_range(kwd_n, utt_n, C, K, U, Tk, Tu, Dk, pair_mode, 0, K, 0, U, w_packed, bias, out_mode, out,
                            stream);
}

extern "C" int kws_sim_stem_v1_range(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk,
                                  int pair_mode, int k0, int nk, int u0, int nu, const void* w_packed,
                                  const float* bias, int out_mode, void* out, void* stream) {
  KWS_CHECK_ARG(kwd_n && utt_n && w_packed && bias && out, "sim_stem: null pointer");
  KWS_CHECK_ARG(C > 0 && K > 0 && U > 0 && Tk > 0 && Tu > 0, "sim_stem: non-positive dimension");
  KWS_CHECK_ARG(k0 >= 0 && nk > 0 && k0 + nk <= K, "sim_stem: keyword range [%d,%d) outside [0,%d)", k0, k0 + nk, K);
  KWS_CHECK_ARG(u0 >= 0 && nu > 0 && u0 + nu <= U, "sim_stem: utterance range [%d,%d) outside [0,%d)", u0, u0 + nu, U);
  KWS_CHECK_ARG(C <= F_MAX_C, "sim_stem: C=%d > %d layers (use kws_sim + kws_stem)", C, F_MAX_C);
  KWS_CHECK_ARG(Dk % 64 == 0 && Dk >= 64, "sim_stem: Dk=%d must be a multiple of 64", Dk);
  KWS_CHECK_ARG(pair_mode == KWS_PAIRS_ALL || pair_mode == KWS_PAIRS_DIAG, "sim_stem: bad pair_mode %d", pair_mode);
  KWS_CHECK_ARG(pair_mode == KWS_PAIRS_ALL || (U == K && k0 == u0 && nk == nu),
                "sim_stem: KWS_PAIRS_DIAG needs U == K and equal ranges (got K=%d U=%d)", K, U);
  KWS_CHECK_ARG(out_mode == KWS_STEM_OUT_NCHW_F32 || out_mode == KWS_STEM_OUT_NHWC_BF16, "sim_stem: bad out_mode %d",
                out_mode);
  KWS_CHECK_ARG((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "sim_stem: pointers must be 16-byte aligned");
  CUtensorMap mu, mk;
  {
    const uint64_t dims[3] = {(uint64_t)Dk, (uint64_t)Tu, (uint64_t)C * U};
    const uint64_t strides[2] = {(uint64_t)Dk * 2, (uint64_t)Dk * 2 * (uint64_t)Tu};
    const uint32_t box[3] = {64, 128, 1};
    if (int e = make_tensor_map(&mu, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, utt_n, dims, strides, box,
                                CU_TENSOR_MAP_SWIZZLE_128B))
      return e;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Dk, (uint64_t)Tk, (uint64_t)C * K};
    const uint64_t strides[2] = {(uint64_t)Dk * 2, (uint64_t)Dk * 2 * (uint64_t)Tk};
    const uint32_t box[3] = {64, 16, 1};
    if (int e = make_tensor_map(&mk, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, kwd_n, dims, strides, box,
                                CU_TENSOR_MAP_SWIZZLE_128B))
      return e;
  }
  FusedParams p{};
  p.w = reinterpret_cast<const uint4*>(w_packed);
  p.bias = bias;
  p.out = out;
  p.out_mode = out_mode;
  p.C = C, p.K = K, p.U = U, p.Tk = Tk, p.Tu = Tu, p.nkb = Dk / 64;
  p.k0 = k0, p.u0 = u0, p.nk = nk, p.nu = nu;
  p.Ho = (Tk + 1) / 2;
  p.Wo = (Tu + 1) / 2;
  p.col_tiles = (p.Wo + F_TILE_OJ - 1) / F_TILE_OJ;
  p.nP = (p.Ho + 1) / 2;
  p.nQ = p.nP + 2;
  p.n_chunks = (p.nQ + 3) / 4;
  p.diag = pair_mode == KWS_PAIRS_DIAG;
  p.num_items = (long long)nk * (p.diag ? 1 : nu) * p.col_tiles;
  KWS_CUDA(cudaFuncSetAttribute(kws_fused_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
  long long grid = p.num_items;
  const int sms = sm_count();
  if (grid > sms) grid = sms;
  p.dbg = g_fused_v1_dbg;
  kws_fused_v1_kernel<<<(int)grid, F_THREADS, F_SMEM, (cudaStream_t)stream>>>(mu, mk, p);
  KWS_CUDA(cudaGetLastError());
  return 0;
}
