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
      scale = m / fmaxf(sqrtf(ss), eps);
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
    const float ka = w0 / fmaxf(sqrtf(warp_sum(ssa)), eps), kq = w1 / fmaxf(sqrtf(warp_sum(ssq)), eps);
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

// one block per utterance; k rounds of block-wide arg-max over the candidates
// (k is small: the reference uses recall@{1..200}); ties -> lower id first.
__global__ void __launch_bounds__(256) topk_kernel(const float* __restrict__ scores, const int32_t* __restrict__ ids,
                                                   int n, int U, int id_offset, int k, float* __restrict__ os,
                                                   int32_t* __restrict__ oi) {
  extern __shared__ unsigned long long s_taken[];  // bitmap of taken candidates
  __shared__ float s_best[8];
  __shared__ int s_bid[8], s_bpos[8];
  const int u = blockIdx.x, tid = threadIdx.x;
  const int words = (n + 63) / 64;
  for (int i = tid; i < words; i += blockDim.x) s_taken[i] = 0ull;
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    float best = -INFINITY;
    int bid = 0x7fffffff, bpos = -1;
    for (int c = tid; c < n; c += blockDim.x) {
      if ((s_taken[c >> 6] >> (c & 63)) & 1ull) continue;
      const float v = scores[(size_t)c * U + u];
      const int id = ids ? ids[(size_t)c * U + u] : id_offset + c;
      if (bpos < 0 || v > best || (v == best && id < bid)) best = v, bid = id, bpos = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oid = __shfl_xor_sync(0xffffffffu, bid, o);
      const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
      if (op >= 0 && (bpos < 0 || ov > best || (ov == best && oid < bid))) best = ov, bid = oid, bpos = op;
    }
    if ((tid & 31) == 0) s_best[tid >> 5] = best, s_bid[tid >> 5] = bid, s_bpos[tid >> 5] = bpos;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        if (s_bpos[w] >= 0 && (bpos < 0 || s_best[w] > best || (s_best[w] == best && s_bid[w] < bid)))
          best = s_best[w], bid = s_bid[w], bpos = s_bpos[w];
      }
      if (bpos >= 0) {
        s_taken[bpos >> 6] |= 1ull << (bpos & 63);
        os[(size_t)r * U + u] = best;
        oi[(size_t)r * U + u] = bid;
      } else {  // fewer candidates than k
        os[(size_t)r * U + u] = -INFINITY;
        oi[(size_t)r * U + u] = -1;
      }
    }
    __syncthreads();
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

int kws_topk(const float* scores, const int32_t* ids, int n_cand, int U, int id_offset, int k, float* out_scores,
             int32_t* out_ids, void* stream) {
  KWS_CHECK_ARG(scores && out_scores && out_ids, "topk: null pointer");
  KWS_CHECK_ARG(n_cand > 0 && U > 0 && k > 0 && k <= 1024, "topk: need n_cand>0, U>0, 0<k<=1024");
  const size_t smem = (size_t)((n_cand + 63) / 64) * sizeof(unsigned long long);
  KWS_CHECK_ARG(smem <= 48 * 1024, "topk: n_cand=%d too large for one pass (max %d)", n_cand, 48 * 1024 * 8);
  topk_kernel<<<U, 256, smem, (cudaStream_t)stream>>>(scores, ids, n_cand, U, id_offset, k, out_scores, out_ids);
  KWS_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
