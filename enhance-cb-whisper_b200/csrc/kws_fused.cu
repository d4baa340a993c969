// Fused similarity + ResNet-stem kernel (sm_100a): the layer-wise cosine-similarity 'image'
// of a (keyword, utterance) pair is produced tile by tile in tensor memory, converted to fp16
// straight into the shared-memory operand layout of the stem convolution, and consumed there by
// a tap-decomposed tcgen05 implicit GEMM.  The [pairs, C, Tk, Tu] tensor of the reference
// (model.py:174-191) never exists in HBM.
//
//   S_c[i, j]          = < kwd_n[c, k, i, :], utt_n[c, u, j, :] >                 (model.py:210-218)
//   out[oc, oi, oj]    = relu(bias[oc] + sum_{c,di,dj} W'[oc,c,di,dj] S_c[2oi+di-3, 2oj+dj-3])
//                                                        (HF modeling_resnet.py:39-54, BN folded)
//
// Work item = (pair, column tile of 60 output columns); an item walks down the image in steps of
// two output rows (M = 128 = 2 output rows x 64 pixel slots, 60 used).
//
// An SS-mode tcgen05.mma is paced by shared-memory operand bytes (measured: max(M*N/256,
// (bytes A + bytes B)/128) cycles, tools/umma_probe2.cu), so the stem is arranged to read the pixel
// operand as rarely as possible:
//   * N = 128: one MMA applies TWO taps to one read of the pixel operand.  Columns 0..63 of the
//     accumulator ("half a") take tap dj, columns 64..127 ("half b") take tap dj+4 of the same input
//     pixels; half b is therefore the partial sum of the output pixel two slots to the LEFT, and the
//     epilogue adds D_a[x] + D_b[x+2] (warp shuffle).
//   * K = 16 = two 8-element chunks whose distance (the descriptor's leading byte offset) is free:
//     16 bytes = "the same rows one pixel further" (tap dj+2), or the distance to the other
//     column-parity plane.
//   * channels 8..11 of a pixel share a 16-byte chunk with channels 8..11 of its right neighbour in
//     the same plane ("Y' chunk"), so 12 layers cost 1.5 chunks per tap instead of 2.
//   Per kernel row di the 7 taps x 12 channels become 3 MMAs (2 for C <= 8) instead of 7:
//     (X plane0 | X plane0 + 1 px)  a: dj 0, 2   b: dj 4, 6      channels 0..7
//     (Y' plane0 | Y' plane1)       a: dj 0,2 | 1,3   b: dj 4,6 | 5,-   channels 8..11
//     (X plane1 | X plane1 + 1 px)  a: dj 1, 3   b: dj 5, -      channels 0..7
//
// Shared-memory operand of the stem ("ring"): input rows are kept de-interleaved by column
// parity (plane) and by row parity (rp), in blocks [k8 (X, Y')][rp][plane] of NR row slots x
// (64 pixels x 16 B).  For kernel row di the 128 A-rows are 128 consecutive 16-byte pixels starting
// at slot(r0) * 1024 (+16 per pixel of shift), r0 = 4P + di - 3: rows 0..63 read input row r0 (output
// row 2P), rows 64..127 run on into the next slot, which holds input row r0 + 2 (output row 2P+1).
// Slot NR mirrors slot 0 so the run never wraps.
//
// The similarity GEMM works on chunks of 16 input rows: for each layer c, D[128 px, 16 rows] =
// utt tile (128 x Dk, TMA, OOB columns zero-filled = the conv's zero padding) x kwd rows (16 x Dk)^T,
// fp32 in TMEM.  4 converter warps (thread = pixel) pull a chunk out of TMEM layer pair by layer pair
// (each pair's tiles are handed back to the similarity issuer at once, so the next chunk is computed
// while this one is still being pulled), keep it as packed fp16 in registers and feed the ring one
// quantum (4 rows) per stem step.
//
// TMEM: stem accumulators [0,128) and [128,256) (double-buffered), similarity region [256, 256+16C), bias [448,512).
// Roles (384 threads): warp 0 TMA producer, warp 1 stem MMA issuer, warp 2 TMEM allocator + similarity MMA
// issuer (one elected thread each: two independent instruction streams into the one tensor pipe, so the
// barrier polls of one never starve it), warps 4..7 stem epilogue (bf16 results staged in swizzled shared
// memory, one TMA store per warp and step), warps 8..11 converters.  Warp 3 is idle except in the POOL instances, where it
// is the pool warp (MaxPool2d(3,2,1) on the staged tile; see the POOL template flag).
//
// Template instances (the role loops are sensitive to every instruction, so whatever varies per shape or mode is a
// template parameter): NHWC (bf16 channels-last via TMA stores | fp32 NCHW, parity), ROWS (similarity chunk height 16 |
// 32 | 48), MULTI (channel-group passes for more than 12 layers: fp16 partial sums chained through the output tiles),
// S12 (12 layers x Dk = 64 as compile-time constants), RAGGED (keyword length table: rows beyond a keyword skipped and
// filled with relu(bias)), POOL (the max-pool behind the stem applied before the activation leaves the SM).
// Up to 4 layers the stem issues ONE MMA per kernel row (channels 0..3 in the Y' position) instead of two.
//
// Two regimes: in bursts the kernel is bound by two loops of ~2 k cycles per step that overlap
// imperfectly (operands -> similarity -> converters -> ring -> stem MMAs; accumulator -> epilogue -> store); run back to
// back, as the benchmark does, every shape sits at the 1 kW power cap and throughput is energy per pair, ~45 % of
// which are the stem MMAs and their shared-memory operands at 12 layers.
//
// Budget (C = 12, 150 x 1500): the kernel is bound by the shared-memory port (128 B/clk), which every
// operand read of an SS-mode MMA goes through.  Per step (2 output rows x 60 px): stem MMA operands
// 21 x 8 KB = 168 KB, similarity MMA operands 55 KB, TMA operand writes 55 KB, converter stores 16 KB,
// epilogue staging 16 KB written + 16 KB read by the TMA store: ~340 KB = 2.7 k cycles, against 1.8 k
// cycles of tensor math (21 x 64 + 12 x 39.5).  Measured: ~2.5 k cycles per step.
#include <stdlib.h>

#include "kws_common.cuh"
#include "../../include/kws_b200.h"

// cycle counters of the roles (development aid): compiled in only with -DKWS_FUSED_TIMERS, they cost registers
#ifdef KWS_FUSED_TIMERS
#define KWS_CLK() clock64()
// event trace of CTA 0: dbg[148*32 + (role*KWS_TRACE_STEPS + step)*8 + k] = clock64()
#define KWS_TRACE_STEPS 640
#define KWS_TRACE(role, step, k)                                                                   \
  do {                                                                                             \
    if (p.dbg && blockIdx.x == 0 && (step) < KWS_TRACE_STEPS)                                      \
      p.dbg[148 * 32 + ((role)*KWS_TRACE_STEPS + (step)) * 8 + (k)] = clock64();                   \
  } while (0)
#else
#define KWS_CLK() 0ll
#define KWS_TRACE(role, step, k) \
  do {                           \
  } while (0)
#endif

#ifdef KWS_DEBUG_HOOKS
#define KWS_WHATIF(bit) ((p.whatif & (bit)) != 0)
#else
#define KWS_WHATIF(bit) false
#endif

namespace kws {

constexpr int G_THREADS = 384;
constexpr int G_OC = 64;
constexpr int G_TILE_OJ = 60;                      // output columns per item (2*59 + 6 = 124 <= 127 input px); slots 60..63 idle
// POOL instances (MaxPool2d(3,2,1) behind the stem fused in, HF modeling_resnet.py:67,76-77): an item is 29 pooled columns =
// stem columns [58 ct - 1, 58 ct + 57] (59 slots, slot 0 = the left neighbour the first pooling window needs)
constexpr int G_POOL_OJ = 29;
constexpr int G_TILE_OJ_POOL = 2 * G_POOL_OJ;
constexpr int G_NR = 8;                            // ring slots per block = 4 quanta of 2 slots
constexpr int G_BLOCK = (G_NR + 1) * 1024 + 64;    // 9280: +mirror slot, +64 keeps plane 1 on other banks
constexpr int G_RING_BYTES = 8 * G_BLOCK;          // [k8 2][rp 2][plane 2]
constexpr int G_MMA_W_BYTES = 2 * 128 * 16;        // 4096: one MMA's B operand [chunk 2][n 128][8 ch] fp16
constexpr int G_MAX_MMA = 3;                       // stem MMAs per kernel row
constexpr int G_W_BYTES = 7 * G_MAX_MMA * G_MMA_W_BYTES;  // 86016
constexpr int G_NS = 3;                            // similarity operand stages (single pass)
constexpr int G_A_BYTES = 128 * 128;               // utt tile 128 px x 64 dims (SW128)
// Similarity chunk = ROWS keyword frames per MMA (N = ROWS): 16 for up to 12 layers, 32 for <= 6, 48 for <= 4
// (the TMEM region holds C x ROWS <= 192 columns, the converters C/2 x ROWS <= 96 packed registers).  Wide
// models with few layers (the L variant: Dk = 384..1280) spend most of their time in the similarity GEMM;
// a larger N reads the utterance tile once per 48 instead of 16 frames.
__host__ __device__ constexpr int g_stage_bytes(int rows) { return G_A_BYTES + rows * 128; }
constexpr int G_MAX_C = 12;
constexpr int G_ACC_COLS = 128;                    // one stem accumulator: half a | half b
constexpr int G_TMEM_SIM = 2 * G_ACC_COLS;         // similarity region starts at column 256 (<= 192 columns)
constexpr int G_TMEM_BIAS = 448;                   // 64 columns: the folded-BN bias, replicated in every lane
constexpr int G_OUT_STAGE = 4 * 32 * 128;          // per epilogue warp: 32 pixels x 64 bf16 (SW128), source of its TMA stores
constexpr int G_NPAIR = G_MAX_C / 2;                 // similarity tiles are handed over per pair of layers
constexpr int G_NBAR = 2 * G_NS + 2 * G_NPAIR + 4 + 4 + 2 + 2 + 8 + 2;

struct FusedParams {
  const uint4* w;     // fused stem weights (kws_pack_stem_fused)
  const float* bias;  // [64]
  void* out;
  int out_mode;
  int C, K, U, Tk, Tu, nkb, Ho, Wo, col_tiles;  // K, U: operand batch sizes (tensor-map extents)
  int k0, u0, nk, nu;  // scored sub-range: keywords [k0, k0+nk) x utterances [u0, u0+nu); out pair = (k-k0)*nu + (u-u0)
  int c0;        // first layer of this pass in the operand banks (C layers [c0, c0+C) are processed)
  int C_total;   // layers in the operand banks (tensor-map batch = layer * bank size + item)
  int acc_mode;  // 0 single pass | 1 first pass: write fp16 partial sums | 2 middle: += | 3 last: +=, bias, ReLU, bf16
  int reduce_mid;  // middle passes add their partial sums with a TMA reduce-add store (fp16 add in L2) instead of
                   // loading the previous sums, adding in registers and storing
  int n_mma;     // stem MMAs per kernel row: 1 (C <= 4: the Y' MMA alone), 2 (C <= 8) or 3
  int nP;        // stem steps per item = ceil(Ho / 2)
  int nQ;        // quanta (4 input rows) converted per item = nP + 2
  int n_chunks;  // similarity chunks (ROWS input rows = ROWS/4 quanta) per item
  int w_bytes;   // bytes of this pass's stem weights in shared memory (7 * n_mma * 4096)
  int diag;
  int per_kw_u;  // KWS_PAIRS_PER_KEYWORD: utterance-side operand of pair (k, u) is bank item k * per_kw_u + u (else 0)
  int prefetch;  // multi-pass partial sums of step n+1: 1 = loaded into a second staging set during step n;
                 // 2 = pulled into L2 only (TMA prefetch), loaded and awaited in step n+1; 0 = neither
  const int32_t* kwd_len;  // optional [K]: valid frames of every keyword (rows >= len are zero in kwd_n); output rows whose
                           // receptive field lies beyond are relu(bias): no similarity, no stem MMAs, constant fill
  long long num_items;
  void* pool_out;  // POOL instances: bf16 channels-last [pairs, Hp, Wp, 64] (max-pooled stem activation)
  int Hp, Wp;      // pooled image: ceil(Ho / 2) x ceil(Wo / 2)
  int whatif;    // development what-if switches (KWS_DEBUG_HOOKS builds only; results are WRONG when set): 1 epilogue
                 // releases the accumulator at once and does nothing else | 2 no half-b shift (no mailbox, no shuffles) |
                 // 4 no TMA store | 8 stem issues kernel row 0 only | 16 similarity issues one of four k-steps |
                 // 32 only half of every utterance tile is loaded (half the L2 -> SM operand bytes)
  long long* dbg;  // optional [grid][16] cycle counters (issuer 0-4, epilogue 5-7, converter 8-11; development aid), or null
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

// packed fp32x2 add (sm_100) and fused convert + ReLU: 3 instructions per two outputs in the bf16 epilogue
__device__ __forceinline__ uint64_t pack_b64(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint32_t relu_bf16x2(uint64_t v) {  // {lo, hi} fp32 -> bf16x2 (lo in the low half), max(., 0)
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}

__device__ __forceinline__ uint32_t f16x2_of(uint64_t v) {  // {lo, hi} fp32 -> fp16x2 (partial sums between passes)
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}
__device__ __forceinline__ uint64_t f32x2_of_f16x2(uint32_t h) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
  return pack_b64(__float_as_uint(f.x), __float_as_uint(f.y));
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
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
    if (p.per_kw_u) r.u += r.kw * p.per_kw_u;
  }
  return r;
}

// Ragged keywords: the live part of an item.  Keyword frames >= L are zero rows of the operand bank (the 0/1 frame mask
// is folded in as a row scale), so similarity rows >= L are exactly zero and an output row oi whose receptive field
// [2oi-3, 2oi+3] starts at or beyond L is exactly relu(bias).  Live output rows: oi with max(0, 2oi-3) < L.
struct ItemShape {
  int L;   // valid keyword frames (<= Tk)
  int nP;  // live stem steps (2 output rows each); steps [nP, p.nP) are constant fills
  int nQ;  // quanta (4 input rows) the live steps read
};
template <bool RAGGED>
__device__ __forceinline__ int item_len(const FusedParams& p, long long it) {
  if constexpr (!RAGGED) return 0;  // unused
  if (p.kwd_len == nullptr || it >= p.num_items) return p.Tk;
  const int L = __ldg(p.kwd_len + decode_item(p, it).kw);
  return L < 0 ? 0 : (L > p.Tk ? p.Tk : L);
}
template <bool RAGGED>
__device__ __forceinline__ ItemShape item_shape(const FusedParams& p, int L) {
  ItemShape r;
  if constexpr (!RAGGED) {  // dense: launch constants, no registers
    r.L = p.Tk, r.nP = p.nP, r.nQ = p.nQ;
    return r;
  }
  r.L = L;
  if (L >= p.Tk) {
    r.nP = p.nP;
  } else if (L <= 0) {
    r.nP = 0;
  } else {
    int rows = (L + 2) / 2 + 1;
    if (rows > p.Ho) rows = p.Ho;
    r.nP = (rows + 1) >> 1;
  }
  r.nQ = r.nP > 0 ? r.nP + 2 : 0;
  return r;
}

// NHWC: bf16 channels-last through TMA stores; else fp32 NCHW with direct stores (parity).  ROWS: see above.
// MULTI: channel-group passes (acc_mode 1..3) compiled in; single-pass launches use the leaner instance.
// NHM: output channels / 16 per TMEM round trip of the multi-pass epilogue (1 | 2).
// S12: the shape parameters of the 12-layer, Dk = 64 case (LE / LEF models: cfg2, the full groups of cfg3 / cfg5) are
// compile-time constants -- the issuer loops are sensitive to every instruction (a run-time stage count cost 10 %).
// RAGGED: keyword length table honoured (p.kwd_len); the dense instances carry none of its bookkeeping -- the role loops
// are sensitive to every live register and instruction.
// POOL: the 3x3 / stride-2 / pad-1 max-pool that follows the stem (HF ResNetEmbeddings.pooler) is applied before the
// activation leaves the SM: the epilogue warps stage the bf16 tile as before, the otherwise idle warp 3 reduces it
// (3 columns x 2 rows from the staging tile + the previous step's second row, carried in registers) and writes one
// pooled row per step: 1.8 instead of 7.2 MB per pair, no separate max-pool pass.
template <bool NHWC, int ROWS, bool MULTI, int NHM = 2, bool S12 = false, bool RAGGED = false, bool POOL = false>
__global__ void __launch_bounds__(G_THREADS, 1)
kws_fused_kernel(const __grid_constant__ CUtensorMap map_utt, const __grid_constant__ CUtensorMap map_kwd,
                 const __grid_constant__ CUtensorMap map_out_lo, const __grid_constant__ CUtensorMap map_out_hi,
                 const FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int kC = S12 ? 12 : p.C;                                  // layers of this pass
  const int kNkb = S12 ? 1 : p.nkb;                               // 64-wide k-blocks per layer
  const int kNmma = S12 ? 3 : p.n_mma;                            // stem MMAs per kernel row
  const int kWbytes = S12 ? 7 * 3 * G_MMA_W_BYTES : p.w_bytes;    // stem weights in shared memory
  uint8_t* base = smem_raw;  // no alignment slack to spare: the declared 1024-byte alignment is checked below
  constexpr int G_STAGE = g_stage_bytes(ROWS);
  constexpr int NPAIR = 96 / ROWS;                // layer pairs the converters can hold: 6 / 3 / 2
  constexpr int QPC = ROWS / 4;                   // quanta per similarity chunk
  constexpr int NS = G_NS;                        // operand stages
  static_assert(!POOL || NHWC, "the fused max-pool exists for the bf16 channels-last output only");
  constexpr int TILE = POOL ? G_TILE_OJ_POOL : G_TILE_OJ;  // stem columns an item advances by
  constexpr int COL_OFF = POOL ? 1 : 0;                    // slot s holds stem column TILE * ct + s - COL_OFF
  constexpr int SLOTS = POOL ? G_TILE_OJ_POOL + 1 : G_TILE_OJ;  // pixel slots of a row that carry output
  uint8_t* s_ops = base;                          // NS * G_STAGE (each 1024-aligned)
  uint8_t* s_w = s_ops + NS * G_STAGE;            // p.w_bytes (multiple of 4096)
  uint8_t* s_ostage = s_w + kWbytes;              // G_OUT_STAGE (4 x 4 KB, each 1024-aligned); two sets when MULTI
  uint8_t* s_ring = s_ostage + ((MULTI && p.prefetch == 1) ? 2 : 1) * G_OUT_STAGE;  // G_RING_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + G_RING_BYTES - 64);  // from the last block's (unused) bank pad on
  uint64_t* ofull = bars;                 // [NS] TMA -> MMA (similarity operands)
  uint64_t* oempty = ofull + NS;          // [NS] MMA commit -> TMA
  uint64_t* sfull = oempty + NS;          // [G_NPAIR] MMA commit -> converters (similarity tiles of a layer pair ready)
  uint64_t* sempty = sfull + G_NPAIR;     // [G_NPAIR] converters -> MMA (tiles pulled into registers)
  uint64_t* qfull = sempty + G_NPAIR;     // [4] converters -> MMA (ring quantum written)
  uint64_t* qempty = qfull + 4;           // [4] MMA commit -> converters
  uint64_t* afull = qempty + 4;           // [2] MMA commit -> epilogue (stem accumulator ready)
  uint64_t* aempty = afull + 2;           // [2] epilogue -> MMA
  uint64_t* pload = aempty + 2;           // [4][2] TMA load of the previous pass's partial sums -> epilogue warp, per staging set
  uint64_t* pfull = pload + 8;            // POOL: epilogue warps -> pool warp (staging tile of a step written)
  uint64_t* pempty = pfull + 1;           // POOL: pool warp -> epilogue warps (staging tile read)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G_NBAR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  {
    const int w16 = 7 * kNmma * (G_MMA_W_BYTES / 16);
    for (int i = threadIdx.x; i < w16; i += G_THREADS) reinterpret_cast<uint4*>(s_w)[i] = p.w[i];
    // the ring is zeroed once: chunks nobody writes (Y' of the last pixel, unused planes for C <= 8) must hold
    // finite values, they only ever meet zero weights or discarded pixel slots
    for (int i = threadIdx.x; i < (G_RING_BYTES - 64) / 16; i += G_THREADS)  // the last block's pad holds the barriers
      reinterpret_cast<uint4*>(s_ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    if ((smem_u32(smem_raw) & 1023u) != 0) {
      if (threadIdx.x == 0) printf("[kws] dynamic shared memory base 0x%x is not 1024-byte aligned\n", smem_u32(smem_raw));
      __trap();
    }
  }
  fence_proxy_async();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_utt);
    tma_prefetch_desc(&map_kwd);
    tma_prefetch_desc(&map_out_lo);
    tma_prefetch_desc(&map_out_hi);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&ofull[s], 1);
      mbar_init(&oempty[s], 1);
    }
    for (int s = 0; s < G_NPAIR; ++s) {
      mbar_init(&sfull[s], 1);
      mbar_init(&sempty[s], 128);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 128);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&qfull[s], 128);
      mbar_init(&qempty[s], 1);
      mbar_init(&pload[2 * s], 1);
      mbar_init(&pload[2 * s + 1], 1);
    }
    mbar_init(pfull, 128);
    mbar_init(pempty, 1);
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
  const int stages_per_chunk = kC * kNkb;

  if (warp == 0) {
    // ===================== TMA producer: similarity operands =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      [[maybe_unused]] uint32_t tr_stage = 0;
      int L_next = item_len<RAGGED>(p, blockIdx.x);
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const ItemCoord w = decode_item(p, it);
        const ItemShape sh = item_shape<RAGGED>(p, L_next);
        L_next = item_len<RAGGED>(p, it + gridDim.x);
        const int n_chunks = (sh.nQ + QPC - 1) / QPC;
        const int jbase = 2 * (TILE * w.ct - COL_OFF) - 3;  // input column of pixel x = 0 (OOB columns read as zero)
        for (int n = 0; n < n_chunks; ++n) {
          if (ROWS * n - 3 >= sh.L) continue;  // chunk entirely below the image / the keyword (zero rows): nothing to load
          for (int c = 0; c < kC; ++c) {
            for (int kb = 0; kb < kNkb; ++kb) {
              mbar_wait(&oempty[stage], phase ^ 1, 100 + stage);
              KWS_TRACE(3, tr_stage, 2);  // slot free seen by the producer
              uint8_t* sa = s_ops + stage * G_STAGE;
              mbar_arrive_expect_tx(&ofull[stage], KWS_WHATIF(32) ? G_STAGE - G_A_BYTES / 2 : G_STAGE);
              tma_load_3d(&map_utt, &ofull[stage], sa, kb * 64, jbase, (p.c0 + c) * p.U + w.u);
              tma_load_3d(&map_kwd, &ofull[stage], sa + G_A_BYTES, kb * 64, ROWS * n - 3, (p.c0 + c) * p.K + w.kw);
              KWS_TRACE(3, tr_stage, 3);  // loads issued
              ++tr_stage;
              if (++stage == NS) stage = 0, phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== similarity MMA issuer (one elected thread) =====================
    // Runs as far ahead as the operand stages and the single TMEM region allow; its barrier polls never
    // hold up the stem issuer (warp 1), and its MMAs fill the tensor pipe while that one polls.
    if (elect_one()) {
      const uint32_t idesc_sim = make_idesc_f16(128, ROWS, 0);
      const uint32_t ops_u32 = smem_u32(s_ops);
      const uint64_t sdesc0 = make_smem_desc(0, 16, 1024, LAYOUT_SW128);
      int o_stage = 0;
      uint32_t o_phase = 0, g = 0;
      [[maybe_unused]] uint32_t tr_s = 0;
      int L_next = item_len<RAGGED>(p, blockIdx.x);
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const ItemShape sh = item_shape<RAGGED>(p, L_next);
        L_next = item_len<RAGGED>(p, it + gridDim.x);
        const int n_chunks = (sh.nQ + QPC - 1) / QPC;
        for (int n = 0; n < n_chunks; ++n) {
          // All-zero chunk (entirely below the image / the keyword): no operands, no MMAs and NO tile hand-shake -- the
          // converters know the same thing from the same shape and write zeros without touching the tile barriers, so
          // the chunk counter g counts real chunks only.  (With the hand-shake kept, the next item's first chunk could
          // not start before the converters had reached the zero chunk, i.e. finished storing the chunk before it:
          // a start-up bubble of ~1.5 steps per item on 75-row images.)
          if (ROWS * n - 3 >= sh.L) continue;
          for (int st = 0; st < stages_per_chunk; ++st) {
            const int c = st / kNkb, kb = st - c * kNkb;
            // the converters have pulled the previous chunk's tiles of this layer pair out of TMEM
            if (st == 0) KWS_TRACE(2, g, 2);
            if (kb == 0 && (c & 1) == 0) mbar_wait(&sempty[c >> 1], (g & 1) ^ 1, 500 + (c >> 1));
            if (st == 0) KWS_TRACE(2, g, 3);
            mbar_wait(&ofull[o_stage], o_phase, 300 + o_stage);
            if (st == 0) KWS_TRACE(2, g, 4);
            KWS_TRACE(3, tr_s, 4);
            tc_fence_after();
            const uint32_t d = G_TMEM_SIM + c * ROWS;  // TMEM base is 0 (checked at start)
            const uint32_t sa = ops_u32 + o_stage * G_STAGE;
            const uint64_t adesc = sdesc0 + (uint64_t)(sa >> 4);
            const uint64_t bdesc = adesc + (uint64_t)(G_A_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (KWS_WHATIF(16) && k > 0) break;
              umma_f16(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_sim, (kb | k) != 0);
            }
            umma_commit(&oempty[o_stage]);
            KWS_TRACE(3, tr_s, 5);
            ++tr_s;
            if (++o_stage == NS) o_stage = 0, o_phase ^= 1;
            // layer pair complete (or last layer of an odd C): hand its tiles to the converters
            if (kb == kNkb - 1 && ((c & 1) == 1 || c == kC - 1)) umma_commit(&sfull[c >> 1]);
            if (st == stages_per_chunk - 1) KWS_TRACE(2, g, 5);
          }
          ++g;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== stem MMA issuer (one elected thread) =====================
    if (elect_one()) {
      const uint32_t idesc_stem = make_idesc_f16(128, 2 * G_OC, 0);
      // A descriptors: SBO = 128 (8 pixels x 16 B); LBO = distance between the two 8-element K chunks
      const uint64_t adesc_px = make_smem_desc(smem_u32(s_ring), 16, 128, LAYOUT_NONE);        // +1 pixel
      const uint64_t adesc_pl = make_smem_desc(smem_u32(s_ring), G_BLOCK, 128, LAYOUT_NONE);   // other plane
      const uint64_t bdesc0 = make_smem_desc(smem_u32(s_w), 128 * 16, 128, LAYOUT_NONE);
      const int n_mma = kNmma;
      uint32_t qbase = 0;    // global index of the current item's quantum 0
      uint32_t acc_seq = 0;  // global stem step counter -> accumulator buffer
      long long tm_acc = 0, tm_q = 0, tm_issue = 0;
      const long long tm_start = KWS_CLK();
      int L_next = item_len<RAGGED>(p, blockIdx.x);
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const ItemShape sh = item_shape<RAGGED>(p, L_next);
        L_next = item_len<RAGGED>(p, it + gridDim.x);
        int waited = 0;  // quanta of this item known to be in the ring
        for (int P = 0; P < sh.nP; ++P, ++acc_seq) {
          const uint32_t acc = acc_seq & 1;
          const long long t1 = KWS_CLK();
          KWS_TRACE(0, acc_seq, 0);
          mbar_wait(&aempty[acc], ((acc_seq >> 1) & 1) ^ 1, 200 + acc);
          KWS_TRACE(0, acc_seq, 1);
          const long long t2 = KWS_CLK();
          // kernel rows 0..5 read quanta P and P+1 only; quantum P+2 (input row 4P+5) is first touched by di = 6
          while (waited <= P + 1 && waited < sh.nQ) {  // first step of an item only
            const uint32_t G = qbase + waited;
            mbar_wait(&qfull[G & 3], (G >> 2) & 1, 400 + (int)(G & 3));
            ++waited;
          }
          const long long t3 = KWS_CLK();
          KWS_TRACE(0, acc_seq, 2);
          tm_acc += t2 - t1, tm_q += t3 - t2;
          tc_fence_after();
          const uint32_t d = acc * G_ACC_COLS;
          const uint32_t slot_base = 2 * (qbase + P);
#pragma unroll
          for (int di = 0; di < 7; ++di) {
            if (di == 6) {
              const long long u0 = KWS_CLK();
              KWS_TRACE(0, acc_seq, 3);
              while (waited <= P + 2 && waited < sh.nQ) {
                const uint32_t G = qbase + waited;
                mbar_wait(&qfull[G & 3], (G >> 2) & 1, 400 + (int)(G & 3));
                ++waited;
              }
              tm_q += KWS_CLK() - u0;
              KWS_TRACE(0, acc_seq, 5);
              tc_fence_after();
            }
            if (KWS_WHATIF(8) && di != 0 && di != 3) continue;
            const uint32_t slot0 = (slot_base + (di >> 1)) & (G_NR - 1);
            // block [k8][rp][plane]: rp = (di+1)&1; byte offsets in 16-byte units
            const uint32_t row16 = ((((di + 1) & 1) * 2 * G_BLOCK) >> 4) + slot0 * 64;
            const uint64_t b_row = bdesc0 + (uint64_t)((di * n_mma * G_MMA_W_BYTES) >> 4);
            if (n_mma == 1) {
              // up to 4 layers: (Y' plane0 | Y' plane1) carries channels 0..3 of all seven taps -- ONE MMA per kernel row
              // instead of two half-empty X MMAs (the stem MMAs are what the power-capped kernel spends its energy on)
              umma_f16(d, adesc_pl + (uint64_t)(row16 + ((4 * G_BLOCK) >> 4)), b_row, idesc_stem, di != 0);
              if (di == 3) umma_commit(&qempty[(qbase + P) & 3]);
              continue;
            }
            // (X plane0 | +1 px): half a taps 0,2; half b taps 4,6
            umma_f16(d, adesc_px + (uint64_t)row16, b_row, idesc_stem, di != 0);
            if (n_mma == 3) {
              // (Y' plane0 | Y' plane1): channels 8..11 of all seven taps
              umma_f16(d, adesc_pl + (uint64_t)(row16 + ((4 * G_BLOCK) >> 4)), b_row + (uint64_t)(G_MMA_W_BYTES >> 4),
                       idesc_stem, 1);
            }
            // (X plane1 | +1 px): half a taps 1,3; half b tap 5
            umma_f16(d, adesc_px + (uint64_t)(row16 + (G_BLOCK >> 4)),
                     b_row + (uint64_t)(((n_mma - 1) * G_MMA_W_BYTES) >> 4), idesc_stem, 1);
            // quantum P (input rows 4P-3..4P) is read by kernel rows 0..3 only: hand its ring slots back early
            if (di == 3) umma_commit(&qempty[(qbase + P) & 3]);
          }
          umma_commit(&afull[acc]);
          KWS_TRACE(0, acc_seq, 6);
          tm_issue += KWS_CLK() - t3;
        }
        // the two tail quanta were read by the last step only
        for (int q = sh.nP; q < sh.nQ; ++q) umma_commit(&qempty[(qbase + q) & 3]);
        qbase += sh.nQ;
      }
      if (p.dbg) {
        long long* o = p.dbg + (size_t)blockIdx.x * 32;
        o[0] = KWS_CLK() - tm_start, o[1] = 0, o[2] = tm_acc, o[3] = tm_q, o[4] = tm_issue;
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================== stem epilogue =====================
    // lane of TMEM = pixel slot; out[x] = relu(D_a[x] + D_b[x+2] + bias) (half b belongs to the pixel two
    // slots to the left).  bf16 channels-last results are staged per warp in shared memory (32 pixels x 128 B,
    // 128B-swizzled) and leave through one TMA store per warp and step: full 128-byte lines, edges clipped
    // by the tensor map.  The staging buffer doubles as the mailbox through which pixels 30, 31 of a row
    // receive half b of pixels 32, 33 (lanes 0, 1 of the neighbouring warp).
    const int q = warp & 3;
    const int row_sel = q >> 1;
    const int ojl = (q & 1) * 32 + lane;
    const bool lo_warp = (q & 1) == 0;
    constexpr bool nhwc = NHWC;
    uint8_t* const stage0 = s_ostage + q * (32 * 128);  // MULTI: set (step & 1) at + G_OUT_STAGE
    // mailboxes (2 lanes x 32 fp32 per channel group) live in the hi warp's never-stored pixel rows 28..31
    float* mbox0 = reinterpret_cast<float*>(s_ostage + (q | 1) * (32 * 128) + 28 * 128);
    float* mbox1 = mbox0 + 64;
    const CUtensorMap* my_map = lo_warp ? &map_out_lo : &map_out_hi;
    // The bias lives in TMEM (64 columns, every lane holds all 64 values): with the whole L1 carved into shared
    // memory a per-step bias read would cost either an L2 round trip or shared-memory port wavefronts.
    {
      const uint32_t t_bias = tmem_base + G_TMEM_BIAS + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t b[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) b[e] = __float_as_uint(__ldg(p.bias + ch * 16 + e));
        tmem_st16(t_bias + ch * 16, b);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    uint32_t acc_seq = 0;
    const bool loads_prev = MULTI && (p.acc_mode == 3 || (p.acc_mode == 2 && !p.reduce_mid));
    uint32_t pl_seq[2] = {0u, 0u};  // partial-sum tiles consumed so far from each staging set (phase of pload[q][set])
    // Partial sums of the previous channel-group pass: the tile of step n+1 is TMA-loaded into the other staging set
    // while step n is processed (a load issued and awaited inside one step would put its whole latency, ~2 k cycles
    // behind the tensor core's operand reads, on the epilogue's critical path).
    auto load_prev = [&](uint32_t set, int ct, int oi, int pair) {
      mbar_arrive_expect_tx(&pload[2 * q + set], (lo_warp ? 32u : (uint32_t)(G_TILE_OJ - 32)) * 128u);
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
          "%6}], [%2];" ::"r"(smem_u32(stage0 + set * G_OUT_STAGE)),
          "l"(reinterpret_cast<uint64_t>(my_map)), "r"(smem_u32(&pload[2 * q + set])), "r"(0),
          "r"(ct * TILE + (q & 1) * 32 - COL_OFF), "r"(oi), "r"(pair)
          : "memory");
    };
    auto tile_ok_at = [&](int ct, int oi) { return oi < p.Ho && ct * TILE + (q & 1) * 32 - COL_OFF < p.Wo; };
    if (MULTI && nhwc && p.prefetch == 1 && loads_prev && lane == 0 && (long long)blockIdx.x < p.num_items) {
      const ItemCoord w0 = decode_item(p, blockIdx.x);
      if (tile_ok_at(w0.ct, row_sel)) load_prev(0u, w0.ct, row_sel, (int)w0.pair);
    }
    long long te_wait = 0, te_ld = 0, te_rest = 0, te_pl = 0, te_lds = 0;
    long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int L_next = item_len<RAGGED>(p, blockIdx.x);
    bool stage_const = false;  // this warp's staging tile holds the constant relu(bias) tile of the fill steps
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const ItemCoord w = decode_item(p, it);
      const ItemShape sh = item_shape<RAGGED>(p, L_next);
      L_next = item_len<RAGGED>(p, it + gridDim.x);
      const int oj = w.ct * TILE + ojl - COL_OFF;
      const bool col_ok = ojl < SLOTS && oj >= 0 && oj < p.Wo;
      for (int P = 0; P < sh.nP; ++P, ++acc_seq) {
        stage_const = false;
        const uint32_t acc = acc_seq & 1;
        const long long e0 = KWS_CLK();
        mbar_wait(&afull[acc], (acc_seq >> 1) & 1, 600 + acc);
        const long long e1 = KWS_CLK();
        if (warp == 4 && lane == 0) KWS_TRACE(1, acc_seq, 0);
        te_wait += e1 - e0;
        tc_fence_after();
        if (KWS_WHATIF(1)) {
          tc_fence_before();
          mbar_arrive(&aempty[acc]);
          continue;
        }
        const uint32_t t_row = tmem_base + acc * G_ACC_COLS + ((uint32_t)(q * 32) << 16);
        const int oi = 2 * P + row_sel;
        const bool ok = col_ok && oi < p.Ho;
        const bool tile_ok = tile_ok_at(w.ct, oi);  // this warp stores a tile this step
        const uint32_t set = (MULTI && p.prefetch == 1) ? (acc_seq & 1u) : 0u;
        uint8_t* const my_stage = stage0 + set * G_OUT_STAGE;
        float* o32 = nullptr;
        long long oc_stride = 0;
        if (!nhwc) {
          o32 = reinterpret_cast<float*>(p.out) + ((w.pair * G_OC) * p.Ho + oi) * (long long)p.Wo + oj;
          oc_stride = (long long)p.Ho * p.Wo;
        }
        uint8_t* srow = my_stage + lane * 128;
        // 32 output channels per TMEM round trip; 16 in the multi-pass instance, whose partial-sum registers would
        // otherwise spill (with the L1 carved into shared memory a spill is an L2 round trip)
        constexpr int NH = MULTI ? NHM : 2, NG = 4 / NH;
#pragma unroll 1  // a real loop: the role loops are instruction-cache bound, not ILP bound
        for (int grp = 0; grp < NG; ++grp) {
          uint32_t va[NH][16], vb[NH][16], vbias[NH][16];
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            tmem_ld16(t_row + G_OC + grp * (16 * NH) + h * 16, vb[h]);
            tmem_ld16(t_row + grp * (16 * NH) + h * 16, va[h]);
            tmem_ld16(tmem_base + G_TMEM_BIAS + ((uint32_t)(q * 32) << 16) + grp * (16 * NH) + h * 16, vbias[h]);
          }
          if (grp == 0) {
            // POOL: the pool warp must have read the previous step's tile out of the staging buffers before they are
            // written again: here when the partial sums of a previous pass are TMA-loaded into them, else as late as
            // possible (just before this step's first store, below)
            if constexpr (POOL && MULTI) mbar_wait(pempty, (acc_seq & 1) ^ 1, 970);
            // the previous step's TMA store must have read this warp's staging buffer before it is reused
            if (lane == 0) {
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              if (nhwc && MULTI && loads_prev && p.prefetch != 1) {
                if (tile_ok) load_prev(0u, w.ct, oi, (int)w.pair);  // awaited below, in this step
                if (p.prefetch == 2) {  // next step's tile: HBM -> L2 now, so that its load is an L2 hit
                  int nct = w.ct, noi = oi + 2, npair = (int)w.pair;
                  bool have = P + 1 < sh.nP;
                  if (!have && it + gridDim.x < p.num_items && item_shape<RAGGED>(p, L_next).nP > 0) {
                    const ItemCoord wn = decode_item(p, it + gridDim.x);
                    nct = wn.ct, noi = row_sel, npair = (int)wn.pair, have = true;
                  }
                  if (have && tile_ok_at(nct, noi))
                    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                                     reinterpret_cast<uint64_t>(my_map)),
                                 "r"(0), "r"(nct * TILE + (q & 1) * 32 - COL_OFF), "r"(noi), "r"(npair)
                                 : "memory");
                }
              } else if (nhwc && MULTI && loads_prev) {
                // the other staging set is free now (its store has been read): prefetch the next step's partial sums
                if (P + 1 < sh.nP) {
                  if (tile_ok_at(w.ct, oi + 2)) load_prev(set ^ 1u, w.ct, oi + 2, (int)w.pair);
                } else if (it + gridDim.x < p.num_items) {
                  const ItemCoord wn = decode_item(p, it + gridDim.x);
                  if (tile_ok_at(wn.ct, row_sel)) load_prev(set ^ 1u, wn.ct, row_sel, (int)wn.pair);
                }
              }
            }
            __syncwarp();
          }
          tmem_ld_wait();
          const long long f0 = KWS_CLK();
          if (grp == NG - 1) {  // last TMEM read of this accumulator
            tc_fence_before();
            mbar_arrive(&aempty[acc]);
            if (warp == 4 && lane == 0) KWS_TRACE(1, acc_seq, 1);
            te_ld += KWS_CLK() - e1;
          }
          float* mbox = (grp & 1) == 0 ? mbox0 : mbox1;
          long long f1 = f0;
          if (!KWS_WHATIF(2)) {
          if (!lo_warp && lane < 2) {
            float4* dst = reinterpret_cast<float4*>(mbox + lane * 32);
#pragma unroll
            for (int h = 0; h < NH; ++h)
#pragma unroll
              for (int e = 0; e < 16; e += 4)
                dst[h * 4 + (e >> 2)] = make_float4(__uint_as_float(vb[h][e]), __uint_as_float(vb[h][e + 1]),
                                                    __uint_as_float(vb[h][e + 2]), __uint_as_float(vb[h][e + 3]));
          }
          named_bar_sync(1 + row_sel, 64);  // mailbox written (and the other mailbox has been read)
          f1 = KWS_CLK();
#pragma unroll
          for (int h = 0; h < NH; ++h)
#pragma unroll
            for (int e = 0; e < 16; ++e) vb[h][e] = __shfl_down_sync(0xffffffffu, vb[h][e], 2);
          if (lo_warp && lane >= 30) {
            const float4* src = reinterpret_cast<const float4*>(mbox + (lane - 30) * 32);
#pragma unroll
            for (int h = 0; h < NH; ++h)
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 f = src[h * 4 + (e >> 2)];
                vb[h][e] = __float_as_uint(f.x), vb[h][e + 1] = __float_as_uint(f.y);
                vb[h][e + 2] = __float_as_uint(f.z), vb[h][e + 3] = __float_as_uint(f.w);
              }
          }
          }
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            const int ch = grp * NH + h;
            if constexpr (nhwc) {
              const bool with_bias = !MULTI || p.acc_mode == 0 || p.acc_mode == 3;  // single or last channel-group pass
              const bool add_prev = MULTI && loads_prev && tile_ok;
              uint8_t* c0p = srow + (((2 * ch) ^ (lane & 7)) << 4);
              uint8_t* c1p = srow + (((2 * ch + 1) ^ (lane & 7)) << 4);
              uint32_t prev[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
              if (add_prev) {  // fp16 partial sums of the previous passes, TMA-loaded into the staging tile
                const long long g0 = KWS_CLK();
                if (grp == 0 && h == 0) mbar_wait(&pload[2 * q + set], pl_seq[set] & 1, 950 + q);
                const long long g1 = KWS_CLK();
                const uint4 a = *reinterpret_cast<const uint4*>(c0p), b = *reinterpret_cast<const uint4*>(c1p);
                prev[0] = a.x, prev[1] = a.y, prev[2] = a.z, prev[3] = a.w;
                prev[4] = b.x, prev[5] = b.y, prev[6] = b.z, prev[7] = b.w;
#ifdef KWS_FUSED_TIMERS
                asm volatile("" ::"r"(prev[0]), "r"(prev[7]));  // the loads have landed before the clock is read
                te_pl += g1 - g0, te_lds += KWS_CLK() - g1;
#endif
              }
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                uint64_t s0 = add_f32x2(pack_b64(va[h][e], va[h][e + 1]), pack_b64(vb[h][e], vb[h][e + 1]));
                uint64_t s1 = add_f32x2(pack_b64(va[h][e + 2], va[h][e + 3]), pack_b64(vb[h][e + 2], vb[h][e + 3]));
                if (with_bias) {
                  s0 = add_f32x2(s0, pack_b64(vbias[h][e], vbias[h][e + 1]));
                  s1 = add_f32x2(s1, pack_b64(vbias[h][e + 2], vbias[h][e + 3]));
                }
                if (add_prev) {
                  s0 = add_f32x2(s0, f32x2_of_f16x2(prev[e >> 1]));
                  s1 = add_f32x2(s1, f32x2_of_f16x2(prev[(e >> 1) + 1]));
                }
                o[e >> 1] = with_bias ? relu_bf16x2(s0) : f16x2_of(s0);
                o[(e >> 1) + 1] = with_bias ? relu_bf16x2(s1) : f16x2_of(s1);
              }
              // 16-byte chunks 2ch, 2ch+1 of this pixel's 128-byte row, 128B swizzle (chunk ^ (row & 7));
              // rows 28..31 of the hi warp are idle pixel slots and hold the mailboxes: never written here
              if constexpr (POOL && !MULTI) {
                if (grp == 0 && h == 0) mbar_wait(pempty, (acc_seq & 1) ^ 1, 970);
              }
              if (lo_warp || lane < SLOTS - 32) {
                if (POOL && !ok) {  // outside the image: 0 never wins a max over post-ReLU values
#pragma unroll
                  for (int e = 0; e < 8; ++e) o[e] = 0u;
                }
                *reinterpret_cast<uint4*>(c0p) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(c1p) = make_uint4(o[4], o[5], o[6], o[7]);
              }
            } else {
              float r[16];
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                r[e] = fmaxf(__uint_as_float(va[h][e]) + __uint_as_float(vb[h][e]) + __uint_as_float(vbias[h][e]), 0.f);
                r[e + 1] = fmaxf(__uint_as_float(va[h][e + 1]) + __uint_as_float(vb[h][e + 1]) + __uint_as_float(vbias[h][e + 1]), 0.f);
                r[e + 2] = fmaxf(__uint_as_float(va[h][e + 2]) + __uint_as_float(vb[h][e + 2]) + __uint_as_float(vbias[h][e + 2]), 0.f);
                r[e + 3] = fmaxf(__uint_as_float(va[h][e + 3]) + __uint_as_float(vb[h][e + 3]) + __uint_as_float(vbias[h][e + 3]), 0.f);
              }
              if (ok) {
#pragma unroll
                for (int e = 0; e < 16; ++e) o32[(ch * 16 + e) * oc_stride] = r[e];
              }
            }
          }
          const long long f2 = KWS_CLK();
          tp[(grp & 1) * 3 + 0] += f0 - (grp == 0 ? e1 : tp[7]);
          tp[(grp & 1) * 3 + 1] += f1 - f0;
          tp[(grp & 1) * 3 + 2] += f2 - f1;
          tp[7] = f2;
        }
        if constexpr (nhwc) {
          if (MULTI && loads_prev && tile_ok) ++pl_seq[set];
          if constexpr (POOL) {
            mbar_arrive(pfull);  // this thread's part of the tile is in the staging buffer: over to the pool warp
          } else {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && tile_ok && !KWS_WHATIF(4)) {
            if (MULTI && p.acc_mode == 2 && p.reduce_mid) {
              // out (fp16 partial sums of the earlier passes) += this pass's partial sums, added where the data lives
              asm volatile(
                  "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                      reinterpret_cast<uint64_t>(my_map)),
                  "r"(smem_u32(my_stage)), "r"(0), "r"(w.ct * G_TILE_OJ + (q & 1) * 32), "r"(oi), "r"((int)w.pair)
                  : "memory");
            } else {
              asm volatile(
                  "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                      reinterpret_cast<uint64_t>(my_map)),
                  "r"(smem_u32(my_stage)), "r"(0), "r"(w.ct * G_TILE_OJ + (q & 1) * 32), "r"(oi), "r"((int)w.pair)
                  : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          }
        }
        te_rest += KWS_CLK() - e1;
        if (warp == 4 && lane == 0) KWS_TRACE(1, acc_seq, 2);
        tp[6] += KWS_CLK() - tp[7];
      }
      // ---- fill steps (ragged keywords): output rows beyond the keyword are exactly relu(bias) --------------------
      // Only the pass that writes final values fills (single pass / last channel-group pass); the partial sums of the
      // earlier passes are never read for these rows.
      if (RAGGED && !POOL && sh.nP < p.nP && (!MULTI || p.acc_mode == 0 || p.acc_mode == 3)) {
        const uint32_t t_bias = tmem_base + G_TMEM_BIAS + ((uint32_t)(q * 32) << 16);
        if constexpr (nhwc) {
          if (!stage_const) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging tile free again
            __syncwarp();
            uint8_t* srow = stage0 + lane * 128;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
              uint32_t vbias[16];
              tmem_ld16(t_bias + ch * 16, vbias);
              tmem_ld_wait();
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 16; e += 2) o[e >> 1] = relu_bf16x2(pack_b64(vbias[e], vbias[e + 1]));
              if (lo_warp || lane < G_TILE_OJ - 32) {
                *reinterpret_cast<uint4*>(srow + (((2 * ch) ^ (lane & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(srow + (((2 * ch + 1) ^ (lane & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
              }
            }
            fence_proxy_async();
            __syncwarp();
            stage_const = true;
          }
          if (lane == 0) {
            for (int P = sh.nP; P < p.nP; ++P) {
              const int oi = 2 * P + row_sel;
              if (!tile_ok_at(w.ct, oi)) continue;
              asm volatile(
                  "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                      reinterpret_cast<uint64_t>(my_map)),
                  "r"(smem_u32(stage0)), "r"(0), "r"(w.ct * G_TILE_OJ + (q & 1) * 32), "r"(oi), "r"((int)w.pair)
                  : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          __syncwarp();
        } else {
          const long long oc_stride = (long long)p.Ho * p.Wo;
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t vbias[16];
            tmem_ld16(t_bias + ch * 16, vbias);
            tmem_ld_wait();
            if (col_ok) {
              for (int P = sh.nP; P < p.nP; ++P) {
                const int oi = 2 * P + row_sel;
                if (oi >= p.Ho) continue;
                float* o32 = reinterpret_cast<float*>(p.out) + ((w.pair * G_OC) * p.Ho + oi) * (long long)p.Wo + oj;
#pragma unroll
                for (int e = 0; e < 16; ++e) o32[(ch * 16 + e) * oc_stride] = fmaxf(__uint_as_float(vbias[e]), 0.f);
              }
            }
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before exit
    if (p.dbg && warp == 4 && lane == 0) {
      long long* o = p.dbg + (size_t)blockIdx.x * 32;
      o[5] = te_wait, o[6] = te_ld, o[7] = te_rest, o[12] = te_pl, o[13] = te_lds;
      for (int i = 0; i < 7; ++i) o[16 + i] = tp[i];
    }
  } else if (POOL && warp == 3) {
    // ===================== pool warp: MaxPool2d(3, 2, 1) on the staged bf16 tile =====================
    // Step P stages stem rows 2P and 2P+1 (slot s = stem column TILE ct + s - 1, zeros outside the image); pooled row P =
    // max over rows 2P-1 .. 2P+1 and, for pooled column c of the item, slots 2c .. 2c+2.  Row 2P-1 is the previous
    // step's second row: its column maxima are carried in registers.  Lane = (16-byte channel chunk j, pixel group g):
    // pooled columns g, g+4, ..., 8 units per lane; one warp-wide LDS.128 reads four whole 128-byte pixel rows.
    const int j = lane & 7, g4 = lane >> 3;
    uint32_t rb[4];  // relu(bias) of this lane's 8 channels, bf16x2 (the constant rows of ragged keywords)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      rb[e] = relu_bf16x2(pack_b64(__float_as_uint(__ldg(p.bias + 8 * j + 2 * e)), __float_as_uint(__ldg(p.bias + 8 * j + 2 * e + 1))));
    auto bmax = [](uint32_t a, uint32_t b) {
      uint32_t r;
      asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
      return r;
    };
    // 16-byte chunk j of slot sl = 2 (g4 + 4 i) + t of staged row r: the four 32-pixel staging tiles are contiguous
    // ([row r][half]), so the byte offset is r * 8192 + sl * 128 + swizzle, and the 128B swizzle (chunk ^ (slot & 7))
    // does not depend on i: three base addresses + compile-time offsets
    const uint8_t* pbase[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) pbase[t] = s_ostage + (2 * g4 + t) * 128 + ((j ^ ((2 * g4 + t) & 7)) << 4);
    auto px = [&](int r, int i, int t) { return *reinterpret_cast<const uint4*>(pbase[t] + r * (2 * 32 * 128) + i * (8 * 128)); };
    __nv_bfloat16* const pout = reinterpret_cast<__nv_bfloat16*>(p.pool_out);
    uint32_t seq = 0;
    int L_next = item_len<RAGGED>(p, blockIdx.x);
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const ItemCoord w = decode_item(p, it);
      const ItemShape sh = item_shape<RAGGED>(p, L_next);
      L_next = item_len<RAGGED>(p, it + gridDim.x);
      uint4 carry[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) carry[i] = make_uint4(0u, 0u, 0u, 0u);
      __nv_bfloat16* const obase = pout + ((size_t)w.pair * p.Hp * p.Wp + (size_t)w.ct * G_POOL_OJ + g4) * G_OC + 8 * j;
      const int cols = p.Wp - w.ct * G_POOL_OJ;  // pooled columns of this item that exist
      for (int P = 0; P < sh.nP; ++P, ++seq) {
        mbar_wait(pfull, seq & 1, 980);
        uint4 o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = g4 + 4 * i;
          if (c < G_POOL_OJ) {
            const uint4 a0 = px(0, i, 0), a1 = px(0, i, 1), a2 = px(0, i, 2);
            const uint4 b0 = px(1, i, 0), b1 = px(1, i, 1), b2 = px(1, i, 2);
            uint4 h1;
            h1.x = bmax(bmax(b0.x, b1.x), b2.x), h1.y = bmax(bmax(b0.y, b1.y), b2.y);
            h1.z = bmax(bmax(b0.z, b1.z), b2.z), h1.w = bmax(bmax(b0.w, b1.w), b2.w);
            o[i].x = bmax(bmax(bmax(a0.x, a1.x), a2.x), bmax(h1.x, carry[i].x));
            o[i].y = bmax(bmax(bmax(a0.y, a1.y), a2.y), bmax(h1.y, carry[i].y));
            o[i].z = bmax(bmax(bmax(a0.z, a1.z), a2.z), bmax(h1.z, carry[i].z));
            o[i].w = bmax(bmax(bmax(a0.w, a1.w), a2.w), bmax(h1.w, carry[i].w));
            carry[i] = h1;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(pempty);  // the staging tile may be overwritten
        __nv_bfloat16* orow = obase + (size_t)P * p.Wp * G_OC;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = g4 + 4 * i;
          if (c < G_POOL_OJ && c < cols) *reinterpret_cast<uint4*>(orow + i * (4 * G_OC)) = o[i];
        }
      }
      // ragged keywords: stem rows >= 2 nP are exactly relu(bias) (zero similarity rows), so pooled row nP is the maximum
      // of the carried row and that constant, and the rows below it are the constant
      if (RAGGED && sh.nP < p.nP && (!MULTI || p.acc_mode == 0 || p.acc_mode == 3)) {
        for (int P = sh.nP; P < p.nP; ++P) {
          __nv_bfloat16* orow = obase + (size_t)P * p.Wp * G_OC;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = g4 + 4 * i;
            if (c < G_POOL_OJ && c < cols) {
              uint4 v = make_uint4(rb[0], rb[1], rb[2], rb[3]);
              if (P == sh.nP) v = make_uint4(bmax(v.x, carry[i].x), bmax(v.y, carry[i].y), bmax(v.z, carry[i].z), bmax(v.w, carry[i].w));
              *reinterpret_cast<uint4*>(orow + i * (4 * G_OC)) = v;
            }
          }
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== converters: TMEM similarity rows -> fp16 ring =====================
    // A whole chunk (4 quanta x 4 rows x C layers) is pulled out of TMEM and packed to fp16 registers as soon
    // as its MMAs retire, so the single similarity region is handed back for the next chunk at once; the packed
    // rows then enter the ring at the pace the stem frees slots (quantum P is released after kernel row 3 of step P).
    const int q4 = warp & 3;
    const int x = q4 * 32 + lane;  // pixel of the 128-wide input window
    uint8_t* dst_px = s_ring + (x & 1) * G_BLOCK + (x >> 1) * 16;
    const bool has_left = (x >> 1) > 0;  // Y' chunk of the left neighbour (same plane) takes our channels 8..11 too
    const uint32_t t_lane = tmem_base + G_TMEM_SIM + ((uint32_t)(q4 * 32) << 16);
    const bool yp = kNmma == 3;
    const bool yonly = kNmma == 1;  // up to 4 layers: channels 0..3 travel in the Y' position, the X blocks are not used
    uint32_t g = 0;   // global similarity chunk counter
    uint32_t Gq = 0;  // global quantum counter
    long long tc_sfull = 0, tc_qempty = 0, tc_ld = 0, tc_st = 0;
    int L_next = item_len<RAGGED>(p, blockIdx.x);
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const ItemShape sh = item_shape<RAGGED>(p, L_next);
      L_next = item_len<RAGGED>(p, it + gridDim.x);
      for (int q0 = 0; q0 < sh.nQ; q0 += QPC) {
        const int nq = sh.nQ - q0 < QPC ? sh.nQ - q0 : QPC;
        const bool zero_chunk = 4 * q0 - 3 >= sh.L;  // entirely below the image / keyword: not computed, no hand-shake
        const long long c0 = KWS_CLK();
        const long long c1 = c0;
        if (warp == 8 && lane == 0) KWS_TRACE(2, g, 0);
        // pull the chunk layer pair by layer pair (ROWS rows x 2 layers per TMEM round trip); each pair's tiles go
        // back to the similarity issuer at once, so the next chunk is computed while this one is still being pulled
        uint32_t h2[ROWS][NPAIR];  // [row][layer pair] fp16x2
#pragma unroll
        for (int j = 0; j < NPAIR; ++j) {
          if (zero_chunk) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) h2[r][j] = 0u;
          } else if (2 * j < kC) {
            mbar_wait(&sfull[j], g & 1, 700 + j);
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < ROWS / 16; ++b) {  // 16 rows x 2 layers per TMEM round trip (32 transient registers)
              uint32_t v0[16], v1[16];
              tmem_ld16(t_lane + (2 * j) * ROWS + 16 * b, v0);
              if (2 * j + 1 < kC) {
                tmem_ld16(t_lane + (2 * j + 1) * ROWS + 16 * b, v1);
              } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) v1[r] = 0u;
              }
              tmem_ld_wait();
              if (b == ROWS / 16 - 1) {
                tc_fence_before();
                mbar_arrive(&sempty[j]);
              }
#pragma unroll
              for (int r = 0; r < 16; ++r) h2[16 * b + r][j] = pack_half2(__uint_as_float(v0[r]), __uint_as_float(v1[r]));
            }
          } else {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) h2[r][j] = 0u;
          }
        }
        const long long c2 = KWS_CLK();
        if (warp == 8 && lane == 0) KWS_TRACE(2, g, 1);
        tc_sfull += c1 - c0, tc_ld += c2 - c1;
        // 16 rows (4 quanta) per trip of a real loop: the row registers are shifted down by 16 after each trip, so
        // the store code exists once (static register indices, small instruction footprint) for any chunk height
#pragma unroll 1
        for (int q16 = 0; q16 < QPC; q16 += 4) {
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          if (q16 + qq < nq) {
            const long long c3 = KWS_CLK();
            mbar_wait(&qempty[Gq & 3], ((Gq >> 2) & 1) ^ 1, 800 + (int)(Gq & 3));
            const long long c4 = KWS_CLK();
            if (warp == 8 && lane == 0) KWS_TRACE(3, Gq, 0);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              // input row r = 4q + t - 3: parity (t+1)&1, ring slot (2 Gq + (t >> 1)) mod NR
              const uint32_t slot = (2 * Gq + (t >> 1)) & (G_NR - 1);
              uint8_t* d0 = dst_px + ((t + 1) & 1) * 2 * G_BLOCK + slot * 1024;
              const uint4 vx = make_uint4(h2[4 * qq + t][0], NPAIR > 1 ? h2[4 * qq + t][NPAIR > 1 ? 1 : 0] : 0u,
                                          NPAIR > 2 ? h2[4 * qq + t][NPAIR > 2 ? 2 : 0] : 0u,
                                          NPAIR > 3 ? h2[4 * qq + t][NPAIR > 3 ? 3 : 0] : 0u);
              if (!yonly) {
                *reinterpret_cast<uint4*>(d0) = vx;
                if (slot == 0) *reinterpret_cast<uint4*>(d0 + G_NR * 1024) = vx;  // mirror: tap windows never wrap
              }
              if ((NPAIR > 5 && yp) || yonly) {
                const uint2 vy = yonly ? make_uint2(vx.x, vx.y)
                                       : make_uint2(h2[4 * qq + t][NPAIR > 5 ? 4 : 0], h2[4 * qq + t][NPAIR > 5 ? 5 : 0]);
                uint8_t* dy = d0 + 4 * G_BLOCK;
                *reinterpret_cast<uint2*>(dy) = vy;                    // own chunk, elements 0..3
                if (has_left) *reinterpret_cast<uint2*>(dy - 8) = vy;  // left neighbour's chunk, elements 4..7
                if (slot == 0) {
                  *reinterpret_cast<uint2*>(dy + G_NR * 1024) = vy;
                  if (has_left) *reinterpret_cast<uint2*>(dy + G_NR * 1024 - 8) = vy;
                }
              }
            }
            fence_proxy_async();
            mbar_arrive(&qfull[Gq & 3]);
            if (warp == 8 && lane == 0) KWS_TRACE(3, Gq, 1);
            ++Gq;
            tc_qempty += c4 - c3, tc_st += KWS_CLK() - c4;
          }
        }
        if (ROWS > 16) {
#pragma unroll
          for (int r = 0; r + 16 < ROWS; ++r)
#pragma unroll
            for (int j = 0; j < NPAIR; ++j) h2[r][j] = h2[r + 16][j];
        }
        }
        if (!zero_chunk) ++g;
      }
    }
    if (p.dbg && warp == 8 && lane == 0) {
      long long* o = p.dbg + (size_t)blockIdx.x * 32;
      o[8] = tc_sfull, o[9] = tc_ld, o[10] = tc_qempty, o[11] = tc_st;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

constexpr size_t g_smem_bytes(int rows, int n_mma, bool two_sets = false) {
  return (size_t)G_NS * g_stage_bytes(rows) + (size_t)7 * n_mma * G_MMA_W_BYTES +
         (two_sets ? 2 : 1) * G_OUT_STAGE + G_RING_BYTES - 64 + G_NBAR * 8 + 16;
}
static_assert(g_smem_bytes(16, 3) <= 232448 && g_smem_bytes(32, 2) <= 232448 && g_smem_bytes(48, 2) <= 232448 &&
                  g_smem_bytes(16, 2, true) <= 232448,
              "fused kernel exceeds the 227 KB shared-memory limit");

// Fused-kernel weight layout: [di 7][m n_mma][chunk 2][n 128][e 8] fp16, BN scale folded.
//   n < 64: half a (output channel n); n >= 64: half b (output channel n - 64, tap dj + 4)
//   m = 0            : channels e,      dj = 2*chunk      + 4*half          (X, even plane)
//   m = n_mma - 1    : channels e,      dj = 1 + 2*chunk  + 4*half          (X, odd plane)
//   m = 1 (n_mma = 3): channels 8+(e&3), dj = chunk + 2*(e>>2) + 4*half     (Y', chunk = plane)
//   n_mma = 1 (C <= 4): the Y' layout alone with channels e&3
// taps dj > 6 and channels >= C are zero.
__global__ void pack_stem_fused_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ mean,
                                       const float* __restrict__ var, float eps, int C_total, int c_base, int C,
                                       int n_mma, __half* __restrict__ wp, float* __restrict__ bias) {
  const int total = 7 * n_mma * 2 * 128 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7, n = (i >> 3) & 127, chunk = (i >> 10) & 1;
    const int m = (i >> 11) % n_mma, di = (i >> 11) / n_mma;
    const int half = n >> 6, oc = n & 63;
    int dj, ch;
    if (n_mma == 1) {  // up to 4 layers: the Y' layout with channels 0..3
      dj = chunk + 2 * (e >> 2) + 4 * half, ch = e & 3;
    } else if (m == 0) {
      dj = 2 * chunk + 4 * half, ch = e;
    } else if (m == n_mma - 1) {
      dj = 1 + 2 * chunk + 4 * half, ch = e;
    } else {
      dj = chunk + 2 * (e >> 2) + 4 * half, ch = 8 + (e & 3);
    }
    float v = 0.f;
    if (dj < 7 && ch < C)
      v = w[(((size_t)oc * C_total + c_base + ch) * 7 + di) * 7 + dj] * (gamma[oc] / sqrtf(var[oc] + eps));
    wp[i] = __float2half_rn(v);
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64) bias[i] = beta[i] - mean[i] * gamma[i] / sqrtf(var[i] + eps);
}

}  // namespace kws

using namespace kws;

// Development hooks (cycle counters, forced chunk heights, multi-pass variants) exist only in the -DKWS_DEBUG_HOOKS
// flavour of the library (build.py --debug-hooks -> libkws_b200_dbg.so, used by tools/ and by the variant tests).
// The release library has NO mutable global state here: every knob below is a compile-time constant, no getenv,
// no exported setter, so weights packed by one call can never meet a kernel configured by another.
#ifdef KWS_DEBUG_HOOKS
#define KWS_KNOB static int
#else
#define KWS_KNOB static constexpr int
#endif
#ifdef KWS_DEBUG_HOOKS
static long long* g_fused_dbg = nullptr;
#else
static constexpr long long* g_fused_dbg = nullptr;
#endif
KWS_KNOB g_fused_grid_limit = 0;
KWS_KNOB g_fused_whatif = 0;
KWS_KNOB g_fused_rows = 0;
KWS_KNOB g_fused_s12 = 1;  // 12-layer / Dk = 64 compile-time specialisation
static int fused_n_mma(int C) { return C <= 4 ? 1 : (C <= 8 ? 2 : 3); }
// C > 12 layers are processed in passes over channel groups of 12 layers; the passes chain their
// partial sums through the output buffer itself (fp16, same tiles), the last one adds bias + ReLU -> bf16.
constexpr int G_MAX_C_TOTAL = 64;
// Multi-pass variants (measured at the cfg3 slab, pairs/s): layers per pass 12 | 8, output channels / 16 per epilogue
// TMEM round trip 2 | 1, partial-sum prefetch 0 none | 1 into a second staging set (8-layer groups only) | 2 into L2 only.
//   12,2,2: 185 k (shipped)   12,2,0: 178 k   12,1,0: 162 k   8,2,1: 157 k   8,2,2: 149 k   8,1,1: 143 k   8,2,0: 135 k
// The multi-pass epilogue (TMA load + add of the previous partial sums) is what bounds these shapes: the stem issuer
// waits for accumulators, so fewer, fatter passes win even though the 2-trip epilogue spills 8 registers.
KWS_KNOB g_multi_group = 12;
KWS_KNOB g_multi_nh = 2;
KWS_KNOB g_multi_prefetch = 2;
KWS_KNOB g_multi_reduce = 1;       // middle passes: TMA reduce-add store (1) | load + add + store (0)
KWS_KNOB g_multi_small_first = 0;
KWS_KNOB g_multi_last_pf1 = 1;     // last pass of <= 8 layers (28 KB of weights less): partial sums of step n+1 are loaded into a
                                   // second staging set during step n (+4 % at cfg3: profiles/r02_zero_chunk.log); dense un-pooled only  // remainder group first (no measurable difference; kept as a variant)
#ifdef KWS_DEBUG_HOOKS
extern "C" void kws_debug_set_fused_whatif(int bits) { g_fused_whatif = bits; }
extern "C" void kws_debug_set_fused_s12(int on) { g_fused_s12 = on ? 1 : 0; }
// force the similarity chunk height (16 | 32 | 48) instead of choosing it from the layer count
extern "C" void kws_debug_set_fused_rows(int rows) { g_fused_rows = rows; }
// cap the number of CTAs (to separate per-SM limits from chip-wide L2 limits)
extern "C" void kws_debug_set_fused_grid_limit(int n) { g_fused_grid_limit = n; }
// device buffer [148][32] receiving the roles' cycle counters (needs -DKWS_FUSED_TIMERS as well)
extern "C" void kws_debug_set_fused_counters(long long* dev_buf) { g_fused_dbg = dev_buf; }
extern "C" void kws_debug_set_fused_reduce(int on) { g_multi_reduce = on ? 1 : 0; }
extern "C" void kws_debug_set_fused_small_first(int on) { g_multi_small_first = on ? 1 : 0; }
extern "C" void kws_debug_set_fused_last_pf1(int on) { g_multi_last_pf1 = on ? 1 : 0; }
// NOTE: weights must be re-packed (kws_pack_stem_fused) after changing the group size or order.
extern "C" void kws_debug_set_fused_multi(int group, int nh, int prefetch) {
  g_multi_group = group == 12 ? 12 : 8;
  g_multi_nh = nh == 2 ? 2 : 1;
  g_multi_prefetch = prefetch == 2 ? 2 : ((prefetch == 1 && g_multi_group == 8) ? 1 : 0);
}
static void multi_cfg_from_env() {  // KWS_FUSED_MULTI="group,nh,prefetch[,reduce[,small_first]]", KWS_FUSED_S12, read once
  static bool done = false;
  if (done) return;
  done = true;
  if (const char* e = getenv("KWS_FUSED_S12")) g_fused_s12 = atoi(e);
  if (const char* e = getenv("KWS_FUSED_MULTI")) {
    int g = 8, nh = 1, pf = 1, red = 1, sf = 1;
    const int n = sscanf(e, "%d,%d,%d,%d,%d", &g, &nh, &pf, &red, &sf);
    if (n >= 3) kws_debug_set_fused_multi(g, nh, pf);
    if (n >= 4) kws_debug_set_fused_reduce(red);
    if (n >= 5) kws_debug_set_fused_small_first(sf);
  }
}
#else
static inline void multi_cfg_from_env() {}
#endif
static int fused_group_size(int C) {
  multi_cfg_from_env();
  return C <= G_MAX_C ? G_MAX_C : g_multi_group;
}
static int fused_groups(int C) { return (C + fused_group_size(C) - 1) / fused_group_size(C); }
// Development variant (kws_debug_set_fused_small_first): the remainder group first (cfg3: 8+12+12 instead of 12+12+8), so
// that the last pass, whose epilogue (load + add + bias + ReLU) is the slowest, runs on a full group whose similarity
// pipeline would hide it.  Measured on one box: 198.0 k vs 198.3 k pairs/s -- no difference (the cfg3 kernel sits at the
// 1 kW power cap: it is energy per pair that counts, not which role waits); the plain order stays the default.
static int fused_group_layers(int C, int g) {
  const int gs = fused_group_size(C), n = fused_groups(C);
  if (!g_multi_small_first) return g < n - 1 ? gs : C - g * gs;
  return g == 0 ? C - gs * (n - 1) : gs;
}
static int fused_group_first(int C, int g) {  // first layer of group g
  if (!g_multi_small_first) return g * fused_group_size(C);
  return g == 0 ? 0 : fused_group_layers(C, 0) + fused_group_size(C) * (g - 1);
}
static size_t fused_group_bytes(int Cg) { return (size_t)7 * fused_n_mma(Cg) * G_MMA_W_BYTES; }

extern "C" size_t kws_stem_fused_weight_bytes(int C) {
  if (C <= 0 || C > G_MAX_C_TOTAL) return 0;
  size_t n = 0;
  for (int g = 0; g < fused_groups(C); ++g) n += fused_group_bytes(fused_group_layers(C, g));
  return n;
}

extern "C" int kws_pack_stem_fused(const float* conv_w, const float* gamma, const float* beta, const float* mean,
                                   const float* var, float eps, int C, void* w_fused, float* bias, void* stream) {
  KWS_CHECK_ARG(conv_w && gamma && beta && mean && var && w_fused && bias, "pack_stem_fused: null pointer");
  KWS_CHECK_ARG(C > 0 && C <= G_MAX_C_TOTAL, "pack_stem_fused: C=%d out of (0,%d]", C, G_MAX_C_TOTAL);
  uint8_t* dst = reinterpret_cast<uint8_t*>(w_fused);
  for (int g = 0; g < fused_groups(C); ++g) {
    const int Cg = fused_group_layers(C, g);
    pack_stem_fused_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(conv_w, gamma, beta, mean, var, eps, C, fused_group_first(C, g), Cg,
                                                                 fused_n_mma(Cg), (__half*)dst, bias);
    KWS_CUDA(cudaGetLastError());
    dst += fused_group_bytes(Cg);
  }
  return 0;
}

// 1: fused kernel available for both output modes; 2: bf16 channels-last only (C > 12: multi-pass); 0: no
extern "C" int kws_sim_stem_supported(int C, int Tk, int Tu, int Dk) {
  if (!(C > 0 && C <= G_MAX_C_TOTAL && Dk >= 64 && Dk % 64 == 0 && Tk > 0 && Tu > 0)) return 0;
  return C <= G_MAX_C ? 1 : 2;
}

extern "C" int kws_sim_stem(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk,
                            int pair_mode, const void* w_fused, const float* bias, int out_mode, void* out,
                            void* stream) {
  return kws_sim_stem_range(kwd_n, utt_n, C, K, U, Tk, Tu, Dk, pair_mode, 0, K, 0, U, w_fused, bias, out_mode, out,
                            stream);
}

extern "C" int kws_sim_stem_range(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk,
                                  int pair_mode, int k0, int nk, int u0, int nu, const void* w_fused,
                                  const float* bias, int out_mode, void* out, void* stream) {
  return kws_sim_stem_ragged(kwd_n, utt_n, nullptr, C, K, U, Tk, Tu, Dk, pair_mode, k0, nk, u0, nu, w_fused, bias,
                             out_mode, out, stream);
}

static int sim_stem_core(const void* kwd_n, const void* utt_n, const int32_t* kwd_len, int C, int K, int U, int Tk, int Tu,
                         int Dk, int pair_mode, int k0, int nk, int u0, int nu, const void* w_fused, const float* bias,
                         int out_mode, void* out, void* workspace, void* stream);

extern "C" int kws_sim_stem_ragged(const void* kwd_n, const void* utt_n, const int32_t* kwd_len, int C, int K, int U,
                                   int Tk, int Tu, int Dk, int pair_mode, int k0, int nk, int u0, int nu,
                                   const void* w_fused, const float* bias, int out_mode, void* out, void* stream) {
  KWS_CHECK_ARG(out_mode != KWS_STEM_OUT_POOL_NHWC_BF16, "sim_stem: the pooled output is kws_sim_stem_pool's");
  return sim_stem_core(kwd_n, utt_n, kwd_len, C, K, U, Tk, Tu, Dk, pair_mode, k0, nk, u0, nu, w_fused, bias, out_mode, out,
                       nullptr, stream);
}

// Bytes of the partial-sum workspace kws_sim_stem_pool needs for `pairs` pairs per call: 0 up to 12 layers (single pass).
extern "C" size_t kws_sim_stem_pool_workspace_bytes(int C, long long pairs, int Tk, int Tu) {
  if (C <= G_MAX_C || pairs <= 0 || Tk <= 0 || Tu <= 0) return 0;
  return (size_t)pairs * ((Tk + 1) / 2) * ((Tu + 1) / 2) * G_OC * 2;
}

extern "C" int kws_sim_stem_pool(const void* kwd_n, const void* utt_n, const int32_t* kwd_len, int C, int K, int U, int Tk,
                                 int Tu, int Dk, int pair_mode, int k0, int nk, int u0, int nu, const void* w_fused,
                                 const float* bias, void* out, void* workspace, void* stream) {
  return sim_stem_core(kwd_n, utt_n, kwd_len, C, K, U, Tk, Tu, Dk, pair_mode, k0, nk, u0, nu, w_fused, bias,
                       KWS_STEM_OUT_POOL_NHWC_BF16, out, workspace, stream);
}

static int sim_stem_core(const void* kwd_n, const void* utt_n, const int32_t* kwd_len, int C, int K, int U, int Tk, int Tu,
                         int Dk, int pair_mode, int k0, int nk, int u0, int nu, const void* w_fused, const float* bias,
                         int out_mode, void* out, void* workspace, void* stream) {
  KWS_CHECK_ARG(kwd_n && utt_n && w_fused && bias && out, "sim_stem: null pointer");
  const bool pool = out_mode == KWS_STEM_OUT_POOL_NHWC_BF16;
  KWS_CHECK_ARG(!pool || C <= G_MAX_C || workspace, "sim_stem_pool: C=%d > %d layers needs the partial-sum workspace", C, G_MAX_C);
  KWS_CHECK_ARG(!pool || g_multi_prefetch != 1, "sim_stem_pool: needs partial-sum prefetch mode 0 or 2");
  KWS_CHECK_ARG(!pool || (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sim_stem_pool: workspace must be 16-byte aligned");
  KWS_CHECK_ARG(kwd_len == nullptr || pair_mode != KWS_PAIRS_PER_KEYWORD,
                "sim_stem: a keyword length table cannot be combined with KWS_PAIRS_PER_KEYWORD");
  KWS_CHECK_ARG(kwd_len == nullptr || g_multi_prefetch != 1, "sim_stem: ragged keywords need prefetch mode 0 or 2");
  KWS_CHECK_ARG(C > 0 && K > 0 && U > 0 && Tk > 0 && Tu > 0, "sim_stem: non-positive dimension");
  KWS_CHECK_ARG(k0 >= 0 && nk > 0 && k0 + nk <= K, "sim_stem: keyword range [%d,%d) outside [0,%d)", k0, k0 + nk, K);
  KWS_CHECK_ARG(u0 >= 0 && nu > 0 && u0 + nu <= U, "sim_stem: utterance range [%d,%d) outside [0,%d)", u0, u0 + nu, U);
  KWS_CHECK_ARG(C <= G_MAX_C_TOTAL, "sim_stem: C=%d > %d layers", C, G_MAX_C_TOTAL);
  KWS_CHECK_ARG(C <= G_MAX_C || out_mode == KWS_STEM_OUT_NHWC_BF16 || pool,
                "sim_stem: C=%d > %d layers needs the bf16 channels-last output (multi-pass partial sums)", C, G_MAX_C);
  KWS_CHECK_ARG(Dk % 64 == 0 && Dk >= 64, "sim_stem: Dk=%d must be a multiple of 64", Dk);
  KWS_CHECK_ARG(pair_mode == KWS_PAIRS_ALL || pair_mode == KWS_PAIRS_DIAG || pair_mode == KWS_PAIRS_PER_KEYWORD,
                "sim_stem: bad pair_mode %d", pair_mode);
  KWS_CHECK_ARG(pair_mode != KWS_PAIRS_DIAG || (U == K && k0 == u0 && nk == nu),
                "sim_stem: KWS_PAIRS_DIAG needs U == K and equal ranges (got K=%d U=%d)", K, U);
  KWS_CHECK_ARG(pair_mode != KWS_PAIRS_PER_KEYWORD || (long long)K * U < (1ll << 31) / (C > 0 ? C : 1),
                "sim_stem: KWS_PAIRS_PER_KEYWORD bank of K*U=%lld items too large", (long long)K * U);
  KWS_CHECK_ARG(out_mode == KWS_STEM_OUT_NCHW_F32 || out_mode == KWS_STEM_OUT_NHWC_BF16 || pool, "sim_stem: bad out_mode %d",
                out_mode);
  KWS_CHECK_ARG((reinterpret_cast<uintptr_t>(w_fused) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "sim_stem: pointers must be 16-byte aligned");
  CUtensorMap mu;
  {
    // KWS_PAIRS_PER_KEYWORD: the utterance-side bank holds one item per (keyword, utterance)
    const uint64_t u_bank = pair_mode == KWS_PAIRS_PER_KEYWORD ? (uint64_t)K * U : (uint64_t)U;
    const uint64_t dims[3] = {(uint64_t)Dk, (uint64_t)Tu, (uint64_t)C * u_bank};
    const uint64_t strides[2] = {(uint64_t)Dk * 2, (uint64_t)Dk * 2 * (uint64_t)Tu};
    const uint32_t box[3] = {64, (g_fused_whatif & 32) ? 64u : 128u, 1};  // what-if 32: half the utterance tile's L2 reads
    if (int e = make_tensor_map(&mu, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, utt_n, dims, strides, box,
                                CU_TENSOR_MAP_SWIZZLE_128B))
      return e;
  }
  const int Ho = (Tk + 1) / 2, Wo = (Tu + 1) / 2;
  const long long n_pairs = (long long)nk * (pair_mode == KWS_PAIRS_DIAG ? 1 : nu);
  KWS_CHECK_ARG(n_pairs < (1ll << 31), "sim_stem: too many pairs in one launch");
  CUtensorMap mo_lo, mo_hi, mo_lo_f16, mo_hi_f16;  // *_f16: the same tiles typed fp16 (element type of the reduce-add)
  {
    // bf16 channels-last activation [pairs, Ho, Wo, 64]; one box = 32 (lo warp) or 28 (hi warp: slots 32..59)
    // pixels x 64 channels of one output row.  (Encoded for the fp32 NCHW mode as well, where it is not used.)
    // POOL: the full-resolution tiles only exist as the partial sums of a multi-pass job, in the workspace
    void* full = pool ? (workspace ? workspace : out) : out;  // (single-pass POOL: maps encoded but never used)
    const uint64_t dims[4] = {64, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)n_pairs};
    const uint64_t strides[3] = {128, 128ull * Wo, 128ull * Wo * Ho};
    const uint32_t box_lo[4] = {64, 32, 1, 1}, box_hi[4] = {64, G_TILE_OJ - 32, 1, 1};
    if (int e = make_tensor_map(&mo_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, full, dims, strides, box_lo,
                                CU_TENSOR_MAP_SWIZZLE_128B))
      return e;
    if (int e = make_tensor_map(&mo_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, full, dims, strides, box_hi,
                                CU_TENSOR_MAP_SWIZZLE_128B))
      return e;
    if (int e = make_tensor_map(&mo_lo_f16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, full, dims, strides, box_lo,
                                CU_TENSOR_MAP_SWIZZLE_128B))
      return e;
    if (int e = make_tensor_map(&mo_hi_f16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, full, dims, strides, box_hi,
                                CU_TENSOR_MAP_SWIZZLE_128B))
      return e;
  }
  FusedParams p{};
  p.bias = bias;
  p.out = out;
  p.out_mode = out_mode;
  p.K = K, p.U = pair_mode == KWS_PAIRS_PER_KEYWORD ? K * U : U, p.Tk = Tk, p.Tu = Tu, p.nkb = Dk / 64;
  p.per_kw_u = pair_mode == KWS_PAIRS_PER_KEYWORD ? U : 0;
  p.k0 = k0, p.u0 = u0, p.nk = nk, p.nu = nu;
  p.C_total = C;
  p.Ho = Ho;
  p.Wo = Wo;
  p.col_tiles = (p.Wo + G_TILE_OJ - 1) / G_TILE_OJ;
  p.nP = (p.Ho + 1) / 2;
  p.nQ = p.nP + 2;
  p.Hp = (Ho + 1) / 2, p.Wp = (Wo + 1) / 2;
  p.pool_out = pool ? out : nullptr;
  const int col_tiles_full = p.col_tiles, col_tiles_pool = (p.Wp + G_POOL_OJ - 1) / G_POOL_OJ;

  p.diag = pair_mode == KWS_PAIRS_DIAG;
  const int sms = sm_count();
  p.dbg = g_fused_dbg;
  p.whatif = g_fused_whatif;
  p.kwd_len = kwd_len;
  const int n_groups = fused_groups(C);
  const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(w_fused);
  for (int g = 0; g < n_groups; ++g) {
    const int Cg = fused_group_layers(C, g);
    p.w = reinterpret_cast<const uint4*>(wsrc);
    p.C = Cg;
    p.c0 = fused_group_first(C, g);
    p.n_mma = fused_n_mma(Cg);
    p.acc_mode = n_groups == 1 ? 0 : (g == 0 ? 1 : (g == n_groups - 1 ? 3 : 2));
    p.w_bytes = (int)fused_group_bytes(Cg);
    // the pass that writes final values pools (29 pooled columns per item); earlier passes write full-resolution partial sums
    const bool pool_pass = pool && (p.acc_mode == 0 || p.acc_mode == 3);
    p.col_tiles = pool_pass ? col_tiles_pool : col_tiles_full;
    p.num_items = (long long)nk * (p.diag ? 1 : nu) * p.col_tiles;
    long long grid = p.num_items;
    if (grid > sms) grid = sms;
    if (g_fused_grid_limit > 0 && grid > g_fused_grid_limit) grid = g_fused_grid_limit;
    const int rows = n_groups > 1 ? 16 : (g_fused_rows > 0 ? g_fused_rows : (Cg <= 4 ? 48 : (Cg <= 6 ? 32 : 16)));
    KWS_CHECK_ARG(rows == 16 || (rows == 32 && Cg <= 6) || (rows == 48 && Cg <= 4), "sim_stem: bad chunk rows %d", rows);
    p.n_chunks = (p.nQ + rows / 4 - 1) / (rows / 4);
    CUtensorMap mk;
    {
      const uint64_t dims[3] = {(uint64_t)Dk, (uint64_t)Tk, (uint64_t)C * K};
      const uint64_t strides[2] = {(uint64_t)Dk * 2, (uint64_t)Dk * 2 * (uint64_t)Tk};
      const uint32_t box[3] = {64, (uint32_t)rows, 1};
      if (int e = make_tensor_map(&mk, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, kwd_n, dims, strides, box,
                                  CU_TENSOR_MAP_SWIZZLE_128B))
        return e;
    }
    const bool nh = out_mode == KWS_STEM_OUT_NHWC_BF16 || pool;
    void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, FusedParams);
    const bool s12 = Cg == 12 && p.nkb == 1 && g_fused_s12;
    const bool ragged = kwd_len != nullptr;
    KWS_CHECK_ARG(!ragged || nh || p.acc_mode == 0, "sim_stem: internal: multi-pass needs bf16 output");
    if (p.acc_mode != 0) {
      KWS_CHECK_ARG(rows == 16 && nh, "sim_stem: internal: multi-pass needs 16-row chunks, bf16 output");
      p.prefetch = g_multi_prefetch;
      if (g_multi_last_pf1 && p.acc_mode == 3 && Cg <= 8 && !ragged && !pool_pass) p.prefetch = 1;
      if (ragged)
        kern = s12 ? kws_fused_kernel<true, 16, true, 2, true, true> : kws_fused_kernel<true, 16, true, 2, false, true>;
      else
        kern = s12 ? kws_fused_kernel<true, 16, true, 2, true> : kws_fused_kernel<true, 16, true, 2>;
#ifdef KWS_DEBUG_HOOKS
      if (g_multi_nh == 1 && !ragged) kern = kws_fused_kernel<true, 16, true, 1>;
      KWS_CHECK_ARG(!(g_multi_nh == 1 && ragged), "sim_stem: ragged keywords need the 2-trip multi-pass epilogue");
#endif
    } else if (ragged) {  // the fp32 NCHW (parity) output has ragged instances too: score() with the fp32 body
      kern = rows == 48 ? (nh ? kws_fused_kernel<true, 48, false, 2, false, true> : kws_fused_kernel<false, 48, false, 2, false, true>)
             : rows == 32 ? (nh ? kws_fused_kernel<true, 32, false, 2, false, true> : kws_fused_kernel<false, 32, false, 2, false, true>)
                          : (nh ? kws_fused_kernel<true, 16, false, 2, false, true> : kws_fused_kernel<false, 16, false, 2, false, true>);
      if (rows == 16 && nh && s12) kern = kws_fused_kernel<true, 16, false, 2, true, true>;
    } else {
      kern = rows == 48 ? (nh ? kws_fused_kernel<true, 48, false> : kws_fused_kernel<false, 48, false>)
             : rows == 32 ? (nh ? kws_fused_kernel<true, 32, false> : kws_fused_kernel<false, 32, false>)
                          : (nh ? kws_fused_kernel<true, 16, false> : kws_fused_kernel<false, 16, false>);
      if (rows == 16 && nh && s12) kern = kws_fused_kernel<true, 16, false, 2, true>;
    }
    if (pool_pass) {
      if (p.acc_mode == 3) {
        if (ragged)
          kern = s12 ? kws_fused_kernel<true, 16, true, 2, true, true, true> : kws_fused_kernel<true, 16, true, 2, false, true, true>;
        else
          kern = s12 ? kws_fused_kernel<true, 16, true, 2, true, false, true> : kws_fused_kernel<true, 16, true, 2, false, false, true>;
      } else if (ragged) {
        kern = rows == 48 ? kws_fused_kernel<true, 48, false, 2, false, true, true>
               : rows == 32 ? kws_fused_kernel<true, 32, false, 2, false, true, true>
               : s12 ? kws_fused_kernel<true, 16, false, 2, true, true, true> : kws_fused_kernel<true, 16, false, 2, false, true, true>;
      } else {
        kern = rows == 48 ? kws_fused_kernel<true, 48, false, 2, false, false, true>
               : rows == 32 ? kws_fused_kernel<true, 32, false, 2, false, false, true>
               : s12 ? kws_fused_kernel<true, 16, false, 2, true, false, true> : kws_fused_kernel<true, 16, false, 2, false, false, true>;
      }
    }
    const size_t smem = g_smem_bytes(rows, p.n_mma, p.acc_mode != 0 && p.prefetch == 1);
    KWS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.reduce_mid = g_multi_reduce;
    const bool red = p.acc_mode == 2 && p.reduce_mid;
    kern<<<(int)grid, G_THREADS, smem, (cudaStream_t)stream>>>(mu, mk, red ? mo_lo_f16 : mo_lo, red ? mo_hi_f16 : mo_hi, p);
    KWS_CUDA(cudaGetLastError());
    wsrc += fused_group_bytes(Cg);
  }
  return 0;
}
