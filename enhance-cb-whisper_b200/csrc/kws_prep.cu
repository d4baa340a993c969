// Operand preparation kernels (HBM-bound, CUDA cores):
//   - normalize_rows : L variant -- layer selection + L2 normalise + mask + fp16
//   - cast_rows_bf16 : LE/LEF    -- layer selection + bf16 cast, layer-major rows
//   - weight packing (stem BN fold + tap packing, 16-bit casts); the LEF temporal projector is kws_temporal.cu
// One warp owns one embedding row; every global access is a 128-bit (or the
// widest aligned) coalesced vector access.
#include "kws_common.cuh"
#include "../../include/kws_b200.h"

namespace kws {

constexpr int MAX_LAYERS = 64;
struct LayerIdx {
  int32_t v[MAX_LAYERS];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------
// rows: x fp32 [B,Cin,T,D] -> out 16-bit [C, B, T, D]
// NORMALIZE: out = fp16(x * mask / max(||x||, eps)); else out = bf16(x) or saturating fp16(x)
// ---------------------------------------------------------------------------
constexpr int ROW_MAX_V4 = 16;  // D <= 2048

template <bool NORMALIZE, bool BF16>
__global__ void __launch_bounds__(256) rows_kernel(const float* __restrict__ x, int B, int Cin, int T, int D,
                                                   LayerIdx lidx, int C, const float* __restrict__ mask,
                                                   float eps, uint16_t* __restrict__ out) {
  const int warps_per_block = blockDim.x >> 5;
  const long long n_rows = (long long)C * B * T;
  const int lane = threadIdx.x & 31;
  const int nv4 = D >> 2;
  for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n_rows;
       row += (long long)gridDim.x * warps_per_block) {
    const int t = (int)(row % T);
    const int b = (int)((row / T) % B);
    const int c = (int)(row / ((long long)T * B));
    const float4* src = reinterpret_cast<const float4*>(x + (((long long)b * Cin + lidx.v[c]) * T + t) * D);
    float4 v[ROW_MAX_V4];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_MAX_V4; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv4) {
        v[i] = ldg_stream(src + idx);
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
    }
    float scale = 1.f;
    if (NORMALIZE) {
      ss = warp_sum(ss);
      const float m = mask ? mask[((long long)b * C + c) * T + t] : 1.f;
      scale = eps < 0.f ? m : m / fmaxf(sqrtf(ss), eps);  // eps < 0: rows are used as given (mask + cast only)
    }
    uint2* dst = reinterpret_cast<uint2*>(out + row * D);
#pragma unroll
    for (int i = 0; i < ROW_MAX_V4; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv4) {
        uint2 o;
        if (NORMALIZE) {
          o.x = pack_half2(v[i].x * scale, v[i].y * scale);
          o.y = pack_half2(v[i].z * scale, v[i].w * scale);
        } else if (BF16) {
          o.x = pack_bf162(v[i].x, v[i].y);
          o.y = pack_bf162(v[i].z, v[i].w);
        } else {
          o.x = pack_half2_sat(v[i].x, v[i].y);
          o.y = pack_half2_sat(v[i].z, v[i].w);
        }
        dst[idx] = o;
      }
    }
  }
}

static int launch_rows(bool normalize, int dtype16, const float* x, int B, int Cin, int T, int D, const int32_t* layer_idx,
                       int C, const float* mask, float eps, void* out, cudaStream_t st) {
  KWS_CHECK_ARG(x && out && layer_idx, "rows: null pointer");
  KWS_CHECK_ARG(B > 0 && Cin > 0 && T > 0 && C > 0, "rows: non-positive dimension");
  KWS_CHECK_ARG(C <= MAX_LAYERS, "rows: C=%d > %d", C, MAX_LAYERS);
  KWS_CHECK_ARG(D % 8 == 0 && D <= ROW_MAX_V4 * 128, "rows: D=%d must be a multiple of 8 and <= %d", D,
                ROW_MAX_V4 * 128);
  KWS_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "rows: pointers must be 16-byte aligned");
  LayerIdx li;
  for (int i = 0; i < C; ++i) {
    KWS_CHECK_ARG(layer_idx[i] >= 0 && layer_idx[i] < Cin, "rows: layer_idx[%d]=%d out of [0,%d)", i,
                  layer_idx[i], Cin);
    li.v[i] = layer_idx[i];
  }
  const long long n_rows = (long long)C * B * T;
  const int wpb = 8;
  long long blocks = (n_rows + wpb - 1) / wpb;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (normalize)
    rows_kernel<true, false><<<(int)blocks, wpb * 32, 0, st>>>(x, B, Cin, T, D, li, C, mask, eps, (uint16_t*)out);
  else if (dtype16 == KWS_BF16)
    rows_kernel<false, true><<<(int)blocks, wpb * 32, 0, st>>>(x, B, Cin, T, D, li, C, nullptr, eps, (uint16_t*)out);
  else
    rows_kernel<false, false><<<(int)blocks, wpb * 32, 0, st>>>(x, B, Cin, T, D, li, C, nullptr, eps, (uint16_t*)out);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------
__global__ void pack_stem_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ mean,
                                 const float* __restrict__ var, float eps, int C, int G, __half* __restrict__ wp,
                                 float* __restrict__ bias) {
  const int total = G * 49 * 2 * 64 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7, oc = (i >> 3) & 63, chunk = (i >> 9) & 1;
    const int tap = (i >> 10) % 49, g = (i >> 10) / 49;
    const int ch = g * 16 + chunk * 8 + e;
    float v = 0.f;
    if (ch < C) {
      const float s = gamma[oc] / sqrtf(var[oc] + eps);
      v = w[((size_t)oc * C + ch) * 49 + tap] * s;
    }
    wp[i] = __float2half_rn(v);
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64) bias[i] = beta[i] - mean[i] * gamma[i] / sqrtf(var[i] + eps);
}

__global__ void cast16_kernel(const float* __restrict__ s, uint16_t* __restrict__ d, size_t n, int bf16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = s[i];
    d[i] = (uint16_t)((bf16 ? pack_bf162(v, 0.f) : pack_half2_sat(v, 0.f)) & 0xffffu);
  }
}

// ---------------------------------------------------------------------------
// bilinear resize of similarity images (config #4, the original CB-Whisper classifier):
// F.interpolate(mode="bilinear", align_corners=False, antialias=False) semantics, i.e.
// src = scale * (dst + 0.5) - 0.5 clamped at 0, neighbours (i0, min(i0 + 1, n - 1)).
// in  fp32 [K, U, C, Hs, Ws]; rows >= src_h[k] of keyword k are padding and never read
// out fp16 [K, U, C, Ho, pitch16] (stem input) and/or fp32 [K, U, C, Ho, Wo]
// One block row per (image, output row); threads run along the output columns.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ in, const int32_t* __restrict__ src_h,
                                                              int U, int C, int Hs, int Ws, int Ho, int Wo, int pitch16,
                                                              float* __restrict__ out32, __half* __restrict__ out16) {
  const long long img = blockIdx.y;  // (k * U + u) * C + c
  const int k = (int)(img / ((long long)U * C));
  const int h = src_h ? min(max(src_h[k], 1), Hs) : Hs;
  const float sy = (float)h / (float)Ho, sx = (float)Ws / (float)Wo;
  const float* src = in + img * (long long)Hs * Ws;
  for (int i = blockIdx.x; i < Ho; i += gridDim.x) {
    const float fy = fmaxf(sy * ((float)i + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)fy, h - 1), y1 = y0 + (y0 < h - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
    const float* r0 = src + (long long)y0 * Ws;
    const float* r1 = src + (long long)y1 * Ws;
    for (int j = threadIdx.x; j < Wo; j += blockDim.x) {
      const float fx = fmaxf(sx * ((float)j + 0.5f) - 0.5f, 0.f);
      const int x0 = min((int)fx, Ws - 1), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
      const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
      const float v = ly0 * (lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1)) + ly1 * (lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1));
      if (out32) out32[(img * Ho + i) * (long long)Wo + j] = v;
      if (out16) out16[(img * Ho + i) * (long long)pitch16 + j] = __float2half_rn(v);
    }
  }
}

// ---------------------------------------------------------------------------
// Config #4 operand-side resize.  Bilinear resize (align_corners=False, no antialias) is linear and separable,
// and the similarity image is bilinear in its operands, so
//     resize(kwd . utt^T)[i, j] = < sum_r Wy[i, r] kwd[r], sum_x Wx[j, x] utt[x] >
// The width map is applied to the utterance FRAMES before the GEMM (interp_rows_kernel: 1500 -> 750 frames is the
// exact average of neighbours and halves the GEMM), the height map becomes a 64-wide operand of its own
// (resize_weights_kernel) that the fused similarity+stem kernel contracts with the native-resolution similarity
// (src/model/cb_whisper.py:189-210; torchvision resize == F.interpolate(bilinear, align_corners=False)).
// ---------------------------------------------------------------------------
constexpr int IR_MAX_V4 = 10;  // D <= 1280

__device__ __forceinline__ void bilinear_tap(int i, int n_in, int n_out, int& i0, int& i1, float& w0, float& w1) {
  const float s = (float)n_in / (float)n_out;
  const float f = fmaxf(s * ((float)i + 0.5f) - 0.5f, 0.f);
  i0 = min((int)f, n_in - 1);
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  w1 = f - (float)i0;
  w0 = 1.f - w1;
}

// x fp32 [B,Cin,T,D] -> out fp16 [C,B,T_out,D]: out[j] = w0 * n(x[t0]) + w1 * n(x[t1]), n() = L2 normalisation
__global__ void __launch_bounds__(256) interp_rows_kernel(const float* __restrict__ x, int B, int Cin, int T, int D,
                                                          LayerIdx lidx, int C, int T_out, float eps,
                                                          uint16_t* __restrict__ out) {
  const int warps_per_block = blockDim.x >> 5;
  const long long n_rows = (long long)C * B * T_out;
  const int lane = threadIdx.x & 31;
  const int nv4 = D >> 2;
  for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n_rows;
       row += (long long)gridDim.x * warps_per_block) {
    const int j = (int)(row % T_out);
    const int b = (int)((row / T_out) % B);
    const int c = (int)(row / ((long long)T_out * B));
    int t0, t1;
    float w0, w1;
    bilinear_tap(j, T, T_out, t0, t1, w0, w1);
    const float4* base = reinterpret_cast<const float4*>(x + ((long long)b * Cin + lidx.v[c]) * T * D);
    const float4* s0 = base + (long long)t0 * nv4;
    const float4* s1 = base + (long long)t1 * nv4;
    float4 a[IR_MAX_V4], q[IR_MAX_V4];
    float ssa = 0.f, ssq = 0.f;
#pragma unroll
    for (int i = 0; i < IR_MAX_V4; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv4) {
        a[i] = ldg_stream(s0 + idx);
        q[i] = ldg_stream(s1 + idx);
        ssa += a[i].x * a[i].x + a[i].y * a[i].y + a[i].z * a[i].z + a[i].w * a[i].w;
        ssq += q[i].x * q[i].x + q[i].y * q[i].y + q[i].z * q[i].z + q[i].w * q[i].w;
      }
    }
    // eps < 0: frames are used as given (the reference's plain matmul on pre-normalised states, cb_whisper.py:197)
    const float ka = eps < 0.f ? w0 : w0 / fmaxf(sqrtf(warp_sum(ssa)), eps);
    const float kq = eps < 0.f ? w1 : w1 / fmaxf(sqrtf(warp_sum(ssq)), eps);
    uint2* dst = reinterpret_cast<uint2*>(out + row * D);
#pragma unroll
    for (int i = 0; i < IR_MAX_V4; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv4) {
        uint2 o;
        o.x = pack_half2(a[i].x * ka + q[i].x * kq, a[i].y * ka + q[i].y * kq);
        o.y = pack_half2(a[i].z * ka + q[i].z * kq, a[i].w * ka + q[i].w * kq);
        dst[idx] = o;
      }
    }
  }
}

// Height map of the resize as an operand: out fp16 [C,K,Ho,Hp], row i = the two bilinear taps of output row i over
// the src_h[k] valid frames of keyword k (columns >= src_h[k] are zero), the same for every layer c.
__global__ void __launch_bounds__(256) resize_weights_kernel(const int32_t* __restrict__ src_h, int K, int C, int Hp,
                                                             int Ho, __half* __restrict__ out) {
  const long long total = (long long)C * K * Ho * Hp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % Hp);
    long long q = i / Hp;
    const int row = (int)(q % Ho);
    q /= Ho;
    const int k = (int)(q % K);
    const int h = src_h ? min(max(src_h[k], 1), Hp) : Hp;
    int y0, y1;
    float w0, w1;
    bilinear_tap(row, h, Ho, y0, y1, w0, w1);
    float v = 0.f;
    if (r == y0) v += w0;
    if (r == y1) v += w1;
    out[i] = __float2half_rn(v);
  }
}

// ---------------------------------------------------------------------------
// scores + top-k
// ---------------------------------------------------------------------------
__global__ void scores_kernel(const float* __restrict__ logits, const float* __restrict__ hw, size_t n,
                              float thr, float* __restrict__ scores, uint8_t* __restrict__ det) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float2 l = reinterpret_cast<const float2*>(logits)[i];
    const float m = fmaxf(l.x, l.y);
    const float e0 = expf(l.x - m), e1 = expf(l.y - m);
    float s = e1 / (e0 + e1);
    if (hw) s *= hw[i];
    scores[i] = s;
    if (det) det[i] = s >= thr ? 1 : 0;
  }
}

// Per-utterance top-k over the keyword axis (replaces torch.topk, model.py:523), any n_cand, k <= 1024.
// Candidates are ordered by one 64-bit key, (orderable(score) << 32) | (0xffffffff - id): larger key = higher score,
// lower id on equal scores -- so a sharded top-k followed by a merge gives exactly the single-device result.
// A block sorts one SEGMENT of TK_SEG candidates for TK_UG neighbouring utterances in shared memory (bitonic,
// descending) and keeps the first k keys of each; the survivors of all segments (n' = segments * k) go through the
// same kernel again until one segment is left.  Every level reads its input once, TK_UG neighbouring utterances per
// candidate row (the old kernel walked the whole strided column k times: ~10^10 uncoalesced loads at 100 k x 512, k = 200).
constexpr int TK_SEG = 2048;
constexpr int TK_UG = 4;
constexpr int TK_THREADS = 512;

__device__ __forceinline__ uint32_t f32_orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// in: scores/ids (level 0) or keys_in (later levels), n candidates x U; this block: segment blockIdx.x, utterances
// [TK_UG * blockIdx.y, +TK_UG).  out: keys_out [gridDim.x * k, U], or (final) out_scores/out_ids [k, U].
__global__ void __launch_bounds__(TK_THREADS) topk_segment_kernel(const float* __restrict__ scores, const int32_t* __restrict__ ids,
                                                                  const unsigned long long* __restrict__ keys_in, int n, int U,
                                                                  int id_offset, int k, unsigned long long* __restrict__ keys_out,
                                                                  float* __restrict__ os, int32_t* __restrict__ oi) {
  extern __shared__ unsigned long long s_keys[];  // [TK_UG][TK_SEG]
  const int c0 = blockIdx.x * TK_SEG, u0 = blockIdx.y * TK_UG;
  for (int i = threadIdx.x; i < TK_SEG * TK_UG; i += TK_THREADS) {
    const int c = i / TK_UG, uu = i % TK_UG;
    unsigned long long key = 0ull;  // padding: below every real candidate
    if (c0 + c < n && u0 + uu < U) {
      const size_t g = (size_t)(c0 + c) * U + u0 + uu;
      if (keys_in) {
        key = keys_in[g];
      } else {
        const uint32_t id = ids ? (uint32_t)ids[g] : (uint32_t)(id_offset + c0 + c);
        key = ((unsigned long long)f32_orderable(scores[g]) << 32) | (unsigned long long)(0xffffffffu - id);
        if (key == 0ull) key = 1ull;  // (-NaN, id 0xffffffff) cannot occur with int32 ids; keep 0 for padding only
      }
    }
    s_keys[uu * TK_SEG + c] = key;
  }
  __syncthreads();
  for (int size = 2; size <= TK_SEG; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int p = threadIdx.x; p < TK_UG * (TK_SEG / 2); p += TK_THREADS) {
        const int uu = p / (TK_SEG / 2), q = p % (TK_SEG / 2);
        const int i = ((q / stride) * 2 * stride) + (q % stride), j = i + stride;
        unsigned long long* a = s_keys + uu * TK_SEG;
        const unsigned long long x = a[i], y = a[j];
        const bool desc = (i & size) == 0;  // final merge (size == TK_SEG): descending everywhere
        if ((x < y) == desc) a[i] = y, a[j] = x;
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k * TK_UG; i += TK_THREADS) {
    const int r = i / TK_UG, uu = i % TK_UG;
    if (u0 + uu >= U) continue;
    const unsigned long long key = r < TK_SEG ? s_keys[uu * TK_SEG + r] : 0ull;
    if (keys_out) {
      keys_out[((size_t)blockIdx.x * k + r) * U + u0 + uu] = key;
    } else if (key == 0ull) {  // fewer candidates than k
      os[(size_t)r * U + u0 + uu] = -INFINITY;
      oi[(size_t)r * U + u0 + uu] = -1;
    } else {
      os[(size_t)r * U + u0 + uu] = f32_from_orderable((uint32_t)(key >> 32));
      oi[(size_t)r * U + u0 + uu] = (int32_t)(0xffffffffu - (uint32_t)key);
    }
  }
}


// ---------------------------------------------------------------------------
// MaxPool2d(kernel 3, stride 2, padding 1) on a channels-last bf16 activation: the op that follows the
// stem in the reference (HF modeling_resnet.py ResNetEmbeddings.pooler via src/efficient_kws/resnet.py:53).
// HBM-bound: one thread = one pooled pixel x 8 channels (16 bytes); the nine taps of neighbouring pixels
// overlap and are served by L1/L2, so DRAM sees each input byte once.  Out-of-image taps are skipped
// (== the -inf padding of the reference); NaNs propagate like torch's max_pool2d.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2_nan(pa[i], pb[i]);
  return r;
}

__global__ void __launch_bounds__(256) maxpool_nhwc_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                           long long N, int H, int W, int C8, int Hp, int Wp) {
  const long long total = N * Hp * Wp * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C8);
    long long q = i / C8;
    const int pj = (int)(q % Wp);
    q /= Wp;
    const int pi = (int)(q % Hp);
    const long long n = q / Hp;
    const uint4* img = in + n * H * W * C8 + ch;
    const int r0 = 2 * pi, c0 = 2 * pj;  // centre tap: always inside the image
    uint4 m = __ldg(img + ((long long)r0 * W + c0) * C8);
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr) {
      const int r = r0 + dr;
      if (r < 0 || r >= H) continue;
#pragma unroll
      for (int dc = -1; dc <= 1; ++dc) {
        const int c = c0 + dc;
        if ((dr == 0 && dc == 0) || c < 0 || c >= W) continue;
        m = max_bf16x8(m, __ldg(img + ((long long)r * W + c) * C8));
      }
    }
    out[i] = m;
  }
}

}  // namespace kws

using namespace kws;

extern "C" {

int kws_normalize_rows(const float* x, int B, int Cin, int T, int D, const int32_t* layer_idx, int C,
                       const float* mask, float eps, void* out_f16, void* stream) {
  return launch_rows(true, KWS_F16, x, B, Cin, T, D, layer_idx, C, mask, eps, out_f16, (cudaStream_t)stream);
}

int kws_cast_rows16(const float* x, int B, int Cin, int T, int D, const int32_t* layer_idx, int C, int dtype16,
                    void* out16, void* stream) {
  KWS_CHECK_ARG(dtype16 == KWS_F16 || dtype16 == KWS_BF16, "cast_rows16: bad dtype16 %d", dtype16);
  return launch_rows(false, dtype16, x, B, Cin, T, D, layer_idx, C, nullptr, 0.f, out16, (cudaStream_t)stream);
}

int kws_resize_bilinear(const float* feat_f32, const int32_t* src_h, int K, int U, int C, int Hs, int Ws, int Ho, int Wo,
                        float* out_f32, void* out_f16, int pitch16, void* stream) {
  KWS_CHECK_ARG(feat_f32 && (out_f32 || out_f16), "resize: null pointer");
  KWS_CHECK_ARG(K > 0 && U > 0 && C > 0 && Hs > 0 && Ws > 0 && Ho > 0 && Wo > 0, "resize: non-positive dimension");
  KWS_CHECK_ARG(!out_f16 || pitch16 >= Wo, "resize: pitch16=%d < Wo=%d", pitch16, Wo);
  KWS_CHECK_ARG((long long)U * C <= 65535, "resize: U*C=%lld images per keyword exceed one launch", (long long)U * C);
  // gridDim.y <= 65535: launch whole keywords at a time
  const int kb = (int)(65535ll / ((long long)U * C));
  for (int k0 = 0; k0 < K; k0 += kb) {
    const int nk = K - k0 < kb ? K - k0 : kb;
    const long long i0 = (long long)k0 * U * C;
    dim3 grid(Ho < 32 ? Ho : 32, (unsigned)(nk * U * C));
    resize_bilinear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        feat_f32 + i0 * (long long)Hs * Ws, src_h ? src_h + k0 : nullptr, U, C, Hs, Ws, Ho, Wo, pitch16,
        out_f32 ? out_f32 + i0 * (long long)Ho * Wo : nullptr,
        out_f16 ? reinterpret_cast<__half*>(out_f16) + i0 * (long long)Ho * pitch16 : nullptr);
    KWS_CUDA(cudaGetLastError());
  }
  return 0;
}

int kws_interp_rows(const float* x, int B, int Cin, int T, int D, const int32_t* layer_idx, int C, int T_out, float eps,
                    void* out_f16, void* stream) {
  KWS_CHECK_ARG(x && out_f16 && layer_idx, "interp_rows: null pointer");
  KWS_CHECK_ARG(B > 0 && Cin > 0 && T > 0 && C > 0 && T_out > 0, "interp_rows: non-positive dimension");
  KWS_CHECK_ARG(C <= MAX_LAYERS, "interp_rows: C=%d > %d", C, MAX_LAYERS);
  KWS_CHECK_ARG(D % 8 == 0 && D <= IR_MAX_V4 * 128, "interp_rows: D=%d must be a multiple of 8 and <= %d", D,
                IR_MAX_V4 * 128);
  KWS_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_f16) & 15) == 0,
                "interp_rows: pointers must be 16-byte aligned");
  LayerIdx li;
  for (int i = 0; i < C; ++i) {
    KWS_CHECK_ARG(layer_idx[i] >= 0 && layer_idx[i] < Cin, "interp_rows: layer_idx[%d]=%d out of [0,%d)", i,
                  layer_idx[i], Cin);
    li.v[i] = layer_idx[i];
  }
  const long long n_rows = (long long)C * B * T_out;
  const int wpb = 8;
  long long blocks = (n_rows + wpb - 1) / wpb;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  interp_rows_kernel<<<(int)blocks, wpb * 32, 0, (cudaStream_t)stream>>>(x, B, Cin, T, D, li, C, T_out, eps,
                                                                        (uint16_t*)out_f16);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

int kws_resize_row_weights(const int32_t* src_h, int K, int C, int Hp, int Ho, void* out_f16, void* stream) {
  KWS_CHECK_ARG(out_f16, "resize_row_weights: null pointer");
  KWS_CHECK_ARG(K > 0 && C > 0 && Hp > 0 && Ho > 0, "resize_row_weights: non-positive dimension");
  const long long total = (long long)C * K * Ho * Hp;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  resize_weights_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src_h, K, C, Hp, Ho, (__half*)out_f16);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

size_t kws_stem_weight_bytes(int C) { return (size_t)((C + 15) / 16) * 49 * 2 * 64 * 8 * sizeof(uint16_t); }

int kws_pack_stem_weights(const float* conv_w, const float* gamma, const float* beta, const float* mean,
                          const float* var, float eps, int C, void* w_packed, float* bias, void* stream) {
  KWS_CHECK_ARG(conv_w && gamma && beta && mean && var && w_packed && bias, "pack_stem: null pointer");
  KWS_CHECK_ARG(C > 0 && C <= 64, "pack_stem: C=%d out of (0,64]", C);
  const int G = (C + 15) / 16;
  pack_stem_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(conv_w, gamma, beta, mean, var, eps, C, G,
                                                         (__half*)w_packed, bias);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

int kws_cast_f32_to_16(const float* src, void* dst16, size_t n, int dtype16, void* stream) {
  KWS_CHECK_ARG(src && dst16, "cast: null pointer");
  KWS_CHECK_ARG(dtype16 == KWS_F16 || dtype16 == KWS_BF16, "cast: bad dtype16 %d", dtype16);
  if (n == 0) return 0;
  size_t blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  cast16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (uint16_t*)dst16, n, dtype16 == KWS_BF16);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

int kws_maxpool_nhwc(const void* in_bf16, long long N, int H, int W, int C, void* out_bf16, void* stream) {
  KWS_CHECK_ARG(in_bf16 && out_bf16, "maxpool: null pointer");
  KWS_CHECK_ARG(N > 0 && H > 0 && W > 0, "maxpool: non-positive dimension");
  KWS_CHECK_ARG(C > 0 && C % 8 == 0, "maxpool: C=%d must be a multiple of 8 (16-byte channel chunks)", C);
  KWS_CHECK_ARG(((reinterpret_cast<uintptr_t>(in_bf16) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0,
                "maxpool: pointers must be 16-byte aligned");
  const int Hp = (H + 1) / 2, Wp = (W + 1) / 2;
  const long long total = N * Hp * Wp * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 32;  // grid-stride: 8 resident CTAs per SM, 4 waves
  if (blocks > cap) blocks = cap;
  maxpool_nhwc_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(in_bf16), reinterpret_cast<uint4*>(out_bf16), N, H, W, C / 8, Hp, Wp);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

int kws_scores(const float* logits, const float* hotword_mask, size_t n, float threshold, float* scores,
               uint8_t* detections, void* stream) {
  KWS_CHECK_ARG(logits && scores, "scores: null pointer");
  if (n == 0) return 0;
  size_t blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  scores_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, hotword_mask, n, threshold, scores,
                                                              detections);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

size_t kws_topk_workspace_bytes(int n_cand, int U, int k) {
  if (n_cand <= TK_SEG || U <= 0 || k <= 0) return 0;
  // level-1 survivors + level-2 survivors (ping-pong); later levels are smaller and reuse the two halves
  const size_t s0 = (size_t)((n_cand + TK_SEG - 1) / TK_SEG), n1 = s0 * (size_t)k;
  const size_t s1 = (n1 + TK_SEG - 1) / TK_SEG;
  return (n1 + s1 * (size_t)k) * (size_t)U * sizeof(unsigned long long);
}

int kws_topk(const float* scores, const int32_t* ids, int n_cand, int U, int id_offset, int k, float* out_scores,
             int32_t* out_ids, void* workspace, void* stream) {
  KWS_CHECK_ARG(scores && out_scores && out_ids, "topk: null pointer");
  KWS_CHECK_ARG(n_cand > 0 && U > 0 && k > 0 && k <= 1024, "topk: need n_cand>0, U>0, 0<k<=1024");
  KWS_CHECK_ARG(n_cand <= TK_SEG || workspace, "topk: n_cand=%d > %d needs a workspace of kws_topk_workspace_bytes()",
                n_cand, TK_SEG);
  KWS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "topk: workspace must be 8-byte aligned");
  const size_t smem = (size_t)TK_UG * TK_SEG * sizeof(unsigned long long);
  KWS_CUDA(cudaFuncSetAttribute(topk_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned ugroups = (unsigned)((U + TK_UG - 1) / TK_UG);
  KWS_CHECK_ARG(ugroups <= 65535, "topk: U=%d too large", U);
  unsigned long long* buf[2];
  {
    const size_t s0 = (size_t)((n_cand + TK_SEG - 1) / TK_SEG);
    buf[0] = reinterpret_cast<unsigned long long*>(workspace);
    buf[1] = buf[0] ? buf[0] + s0 * (size_t)k * (size_t)U : nullptr;
  }
  const unsigned long long* keys_in = nullptr;
  long long n = n_cand;
  for (int level = 0;; ++level) {
    const long long segs = (n + TK_SEG - 1) / TK_SEG;
    const bool last = segs == 1;
    unsigned long long* keys_out = last ? nullptr : buf[level & 1];
    topk_segment_kernel<<<dim3((unsigned)segs, ugroups), TK_THREADS, smem, (cudaStream_t)stream>>>(
        level == 0 ? scores : nullptr, level == 0 ? ids : nullptr, keys_in, (int)n, U, id_offset, k, keys_out,
        out_scores, out_ids);
    KWS_CUDA(cudaGetLastError());
    if (last) break;
    keys_in = keys_out;
    n = segs * k;
  }
  return 0;
}

}  // extern "C"
