/*
 * kws_b200.h -- C ABI of the B200-native efficient_kws scoring path.
 *
 * The reference (Priberam/Enhance-CB-Whisper) has no native boundary: its hot
 * path is the body of one nn.Module.forward (src/efficient_kws/model.py:129-221)
 * plus the HuggingFace ResNet stem it calls (src/efficient_kws/resnet.py:38,53).
 * Each entry point below replaces the stock PyTorch ops of one stage of that
 * forward; the "replaces" note gives the reference lines.  The Python mirror of
 * the reference interface (enhance-cb-whisper_b200/model.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a maintainer would add.
 *
 * Conventions
 *  - every function returns int: 0 ok, <0 bad argument (see kws_last_error()),
 *    >0 a cudaError_t raised while launching;
 *  - all data pointers are DEVICE pointers owned by the caller, 16-byte aligned;
 *    nothing is allocated, freed or synchronised inside the library;
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued on it;
 *  - re-entrant from several host threads on different streams: the only global
 *    state is a one-time driver entry-point lookup (std::call_once) and cached
 *    device attributes (atomics).  No environment variable is read and no setter
 *    is exported; the development hooks (kws_debug_*, KWS_FUSED_* variables) exist
 *    only in the separate -DKWS_DEBUG_HOOKS flavour libkws_b200_dbg.so;
 *  - kernels are launched on the CURRENT device (cudaGetDevice): callers make the
 *    device of their pointers current first (the Python front end does);
 *  - layouts are row-major with the last index contiguous.
 *
 * "Layer-major" operand layout: compressed keyword / utterance embeddings are
 * kept as fp16 [C, B, T', Dk] (layer, item, frame, dim), L2-normalised over Dk
 * with the 0/1 frame mask folded in as a row scale, so the similarity of a
 * (keyword, utterance) pair in layer c is a plain K-major GEMM of two slabs.
 */
#ifndef KWS_B200_H_
#define KWS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KWS_ABI_VERSION 8

/* 16-bit operand formats (same encoding as the tcgen05 kind::f16 descriptor) */
#define KWS_F16 0  /* IEEE half: 10-bit mantissa; for L2-normalised data and sane weights */
#define KWS_BF16 1 /* bfloat16: 7-bit mantissa, fp32 range                               */

/* kws_mlp out_mode */
#define KWS_MLP_OUT_NORM_F16 0 /* LE : fp16 [C,R,P], row-normalised * mask          */
#define KWS_MLP_OUT_RAW_F32 1  /* fp32 [C,R,P], un-normalised (parity / inspection)     */
#define KWS_MLP_OUT_RAW_16 2   /* LEF: dtype16 [C,R,P], un-normalised (feeds kws_temporal) */

/* kws_sim pair_mode */
#define KWS_PAIRS_ALL 0
#define KWS_PAIRS_DIAG 1
#define KWS_PAIRS_PER_KEYWORD 2 /* kws_sim_stem only: utt_n holds one item per (keyword, utterance): [C, K*U, Tu, Dk] */

/* kws_stem out_mode */
#define KWS_STEM_OUT_NCHW_F32 0  /* fp32 [pairs,64,Ho,Wo]   (parity with the reference) */
#define KWS_STEM_OUT_NHWC_BF16 1 /* bf16 [pairs,Ho,Wo,64]   (channels_last hand-off)    */
#define KWS_STEM_OUT_POOL_NHWC_BF16 2 /* bf16 [pairs,ceil(Ho/2),ceil(Wo/2),64]: stem + MaxPool2d(3,2,1), kws_sim_stem_pool only */

int kws_abi_version(void);
/* thread-local description of the last non-zero return value */
const char* kws_last_error(void);
/* number of SMs of the current device (grid sizing; 148 on B200) */
int kws_sm_count(void);

/* ---- once per checkpoint ------------------------------------------------ */

/* Fold BatchNorm2d (eval) into the stem convolution and pack it for the
 * tap-decomposed implicit GEMM.  Replaces nothing at run time; prepares
 * model.feature_extractor.embedder.embedder.{convolution,normalization}
 * (HF modeling_resnet.py:39-54 via src/efficient_kws/resnet.py:38).
 *   conv_w fp32 [64,C,7,7]; gamma,beta,mean,var fp32 [64]
 *   w_packed fp16 [G][49][2][64][8], G = ceil(C/16); channel = 16 g + 8 chunk + e
 *   bias fp32 [64] = beta - mean * gamma / sqrt(var + eps)                     */
int kws_pack_stem_weights(const float* conv_w, const float* gamma, const float* beta, const float* mean,
                          const float* var, float eps, int C, void* w_packed, float* bias, void* stream);
size_t kws_stem_weight_bytes(int C);

/* The same fold, packed for the fused similarity+stem kernel (kws_sim_stem*): per kernel row the seven
 * taps x C channels are laid out as 2 (C <= 8) or 3 (C <= 12) N=128 B-operands, two taps per MMA and
 * channels 8..11 of two neighbouring pixels per 16-byte chunk (layout documented in csrc/kws_fused.cu).
 *   w_fused fp16, kws_stem_fused_weight_bytes(C) bytes (0 if the fused kernel does not cover C)        */
int kws_pack_stem_fused(const float* conv_w, const float* gamma, const float* beta, const float* mean,
                        const float* var, float eps, int C, void* w_fused, float* bias, void* stream);
size_t kws_stem_fused_weight_bytes(int C);

/* Fold BatchNorm1d (eval) into the LEF temporal Conv1d and pack it for kws_temporal.
 * time_projector[i] = Conv1d(P,P,3,pad 1) -> BatchNorm1d -> MaxPool1d(3,2,1)
 * (src/efficient_kws/model.py:107-124).
 *   conv_w fp32 [C,P,P,3], conv_b/gamma/beta/mean/var fp32 [C,P]
 *   w16_packed fp16|bf16 [C,3,P/8,P(out),8(in)] (B operand of each tap, K-chunk-major: the kernel's
 *   shared-memory image), b_folded fp32 [C,P].  P % 8 == 0.                     */
int kws_fold_temporal_weights(const float* conv_w, const float* conv_b, const float* gamma, const float* beta,
                              const float* mean, const float* var, float eps, int C, int P, int dtype16,
                              void* w16_packed, float* b_folded, void* stream);

/* fp32 -> fp16/bf16, saturating (projector Linear weights, src/efficient_kws/model.py:92-104) */
int kws_cast_f32_to_16(const float* src, void* dst16, size_t n, int dtype16, void* stream);

/* ---- per batch of keywords or utterances --------------------------------- */

/* L variant operand preparation: sparse layer selection + L2 normalisation +
 * mask folding + fp16 cast.  Replaces the norm/clamp/divide of sim_matrix
 * (model.py:214-216), the layer slice x[-n_layers:] (dataset.py:570-573) and
 * the two mask multiplies (model.py:187-191).
 *   x fp32 [B,Cin,T,D]; layer_idx HOST int32 [C] (indices into Cin)
 *   mask fp32 [B,C,T] or NULL; out fp16 [C,B,T,D]
 *   eps < 0: no normalisation (rows used as given: selection + mask + saturating-free fp16 cast) -- config #4,
 *   whose caller normalises without eps beforehand (src/model/cb_whisper.py:106) and then does a plain matmul  */
int kws_normalize_rows(const float* x, int B, int Cin, int T, int D, const int32_t* layer_idx, int C,
                       const float* mask, float eps, void* out_f16, void* stream);

/* LE/LEF input staging: layer selection + 16-bit cast into layer-major rows.
 *   x fp32 [B,Cin,T,D] -> out fp16|bf16 [C, B*T, D]                            */
int kws_cast_rows16(const float* x, int B, int Cin, int T, int D, const int32_t* layer_idx, int C, int dtype16,
                    void* out16, void* stream);

/* Per-layer MLP projector[i] = Linear(D,H) -> ReLU -> Linear(H,P), H = D/2
 * (model.py:92-104, applied :146-150), as two tcgen05 GEMMs with fused
 * bias/ReLU and bias/normalise/mask epilogues.
 *   x [C,R,D] (R = B*T rows per layer), w1 [C,H,D], w2 [C,P,H], hidden workspace [C,R,H]: all of
 *   the 16-bit type dtype16 (KWS_F16 | KWS_BF16); b1 fp32 [C,H], b2 fp32 [C,P]
 *   mask fp32 [B,C,T] or NULL (only used by KWS_MLP_OUT_NORM_F16)
 *   out: see out_mode.  Requires D % 64 == 0, H % 64 == 0, P % 16 == 0, P <= 256 */
int kws_mlp(const void* x16, int C, int B, int T, int D, int H, int P, int dtype16, const void* w1_16,
            const float* b1, const void* w2_16, const float* b2, void* hidden16, const float* mask, float eps,
            int out_mode, void* out, void* stream);

/* The same projector as ONE kernel, fed by the raw fp32 embeddings: layer selection (dataset.py:570-573), fp32 -> 16-bit
 * cast, Linear(D,H) + ReLU, Linear(H,P) and the normalise * mask epilogue (model.py:92-104, :146-150, :214-216,
 * :187-191) fused; the hidden activation stays in tensor / shared memory and x is read once (H <= 384; twice from L2
 * for larger H, which run as two hidden passes).  Replaces kws_cast_rows16 + kws_mlp.
 *   x fp32 [B,Cin,T,D]; layer_idx HOST int32 [C]; w1 [C,H,D], w2 [C,P,H] of the 16-bit type dtype16; b1 fp32 [C,H],
 *   b2 fp32 [C,P]; mask fp32 [B,C,T] or NULL (KWS_MLP_OUT_NORM_F16 only); out [C, B*T, P] as kws_mlp's out_mode.
 *   kws_mlp_fused_supported(D,H,P) == 1: D % 64 == 0, H % 64 == 0 and H = n * NH with NH <= 384 a multiple of 64
 *   (n <= 4), P % 16 == 0, P <= 64, and the stages fit 227 KB of shared memory; otherwise use the two-call form. */
int kws_mlp_fused(const float* x, int B, int Cin, int T, int D, const int32_t* layer_idx, int C, int H, int P,
                  int dtype16, const void* w1_16, const float* b1, const void* w2_16, const float* b2, const float* mask,
                  float eps, int out_mode, void* out, void* stream);
int kws_mlp_fused_supported(int D, int H, int P);

/* LEF temporal projector (BN folded) + MaxPool1d(3,2,1) + L2 normalisation +
 * mask folding (model.py:107-124, :152-166, :214-216, :187-191) as a tcgen05 implicit GEMM over the
 * frame axis (three taps = one shared-memory tile read at three row offsets).
 *   proj16 fp16|bf16 [C,B,T,P] (kws_mlp, KWS_MLP_OUT_RAW_16) -> out fp16 [C,B,T2,P], T2 = ceil(T/2)
 *   w16_packed / b_folded from kws_fold_temporal_weights (same dtype16)
 *   mask fp32 [B,C,T2] (pooled resolution) or NULL.  P % 32 == 0, 32 <= P <= 128. */
int kws_temporal(const void* proj16, int C, int B, int T, int P, int dtype16, const void* w16_packed,
                 const float* b_folded, const float* mask, float eps, void* out_f16, void* stream);

/* ---- per (keyword, utterance) pair ---------------------------------------- */

/* Layer-wise cosine-similarity 'images' of all K x U pairs: batched
 * tcgen05/TMEM GEMM fed by TMA (replaces torch.bmm + permute + stack + mask,
 * model.py:174-191, :217).
 *   kwd_n fp16 [C,K,Tk,Dk], utt_n fp16 [C,U,Tu,Dk] (prepared operands)
 *   feat_f32 fp32 [K,U,C,Tk,Tu] or NULL   (== KWSOutput.features per utterance)
 *   feat_f16 fp16 [K,U,C,Tk,pitch16] or NULL (stem input), pitch16 >= Tu
 *   pair_mode KWS_PAIRS_ALL: every keyword against every utterance (K*U pairs);
 *             KWS_PAIRS_DIAG: U == K, keyword k against utterance k only (the
 *             training-style batch of model.py:171-173; outputs are [K,C,Tk,*])
 *   Requires Dk % 64 == 0, Tk <= 256.                                          */
int kws_sim(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk, int pair_mode,
            float* feat_f32, void* feat_f16, int pitch16, void* stream);

/* ResNet stem Conv2d(C,64,7,stride 2,pad 3,bias=False) + BatchNorm2d + ReLU as
 * a tap-decomposed tcgen05 implicit GEMM (HF modeling_resnet.py:39-54 via
 * src/efficient_kws/resnet.py:38,53).
 *   feat_f16 fp16 [pairs,C,Tk,pitch16]; w_packed/bias from kws_pack_stem_weights
 *   out: [pairs,64,Ho,Wo] fp32 or [pairs,Ho,Wo,64] bf16, Ho=ceil(Tk/2), Wo=ceil(Tu/2)
 *   C <= 16 runs in one launch.  C > 16 runs one launch per 16-channel group and chains the
 *   partial sums through `workspace` (fp32, kws_stem_workspace_bytes(); may be NULL for C <= 16). */
int kws_stem(const void* feat_f16, int pairs, int C, int Tk, int Tu, int pitch16, const void* w_packed,
             const float* bias, int out_mode, void* out, void* workspace, void* stream);
size_t kws_stem_workspace_bytes(int pairs, int C, int Tk, int Tu);

/* Fused similarity + stem: kws_sim followed by kws_stem without the [pairs,C,Tk,Tu] tensor ever
 * reaching HBM -- similarity tiles go TMEM -> fp16 shared-memory ring -> stem MMAs in one kernel
 * (replaces model.py:174-191,:217 and HF modeling_resnet.py:39-54 via resnet.py:38,53 together).
 *   kwd_n fp16 [C,K,Tk,Dk], utt_n fp16 [C,U,Tu,Dk] (prepared operands, mask folded)
 *   w_fused/bias from kws_pack_stem_fused; pair_mode / out_mode / out as kws_sim / kws_stem
 *   Requires Dk % 64 == 0.  C <= 12 runs as one launch in either output mode
 *   (kws_sim_stem_supported() == 1); 12 < C <= 64 runs one launch per group of 12 layers, the passes chaining
 *   fp16 partial sums through `out` itself, bf16 channels-last output only (== 2); otherwise (== 0) use
 *   kws_sim + kws_stem.                                                                     */
int kws_sim_stem(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk, int pair_mode,
                 const void* w_fused, const float* bias, int out_mode, void* out, void* stream);
int kws_sim_stem_supported(int C, int Tk, int Tu, int Dk);
/* Same kernel over a sub-block of the pair grid: keywords [k0, k0+nk) x utterances [u0, u0+nu) of the
 * resident operand banks (K, U stay the bank sizes).  out holds nk*nu pairs, pair = (k-k0)*nu + (u-u0)
 * (KWS_PAIRS_DIAG: nk pairs, ranges must coincide).  This is how the batched replacement of the
 * test_step group loop (model.py:769-780) streams a K x U job through a bounded activation buffer
 * without re-packing operand slabs. */
int kws_sim_stem_range(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk,
                       int pair_mode, int k0, int nk, int u0, int nu, const void* w_fused, const float* bias,
                       int out_mode, void* out, void* stream);

/* The same with a keyword LENGTH TABLE (ragged keyword bank: real keywords are 10-60 of the 150 padded frames,
 * src/efficient_kws/dataset.py:784-819).  kwd_len DEVICE int32 [K]: frames >= kwd_len[k] of keyword k are zero rows of
 * kwd_n (the 0/1 frame mask folded in by the compression kernels guarantees it).  Output rows whose receptive field
 * starts at or beyond kwd_len[k] are exactly relu(bias): the kernel skips their similarity and stem MMAs and fills
 * them with the constant tile.  The output is bit-identical to kws_sim_stem_range; kwd_len == NULL is that call.
 * Not with KWS_PAIRS_PER_KEYWORD. */
int kws_sim_stem_ragged(const void* kwd_n, const void* utt_n, const int32_t* kwd_len, int C, int K, int U, int Tk, int Tu,
                        int Dk, int pair_mode, int k0, int nk, int u0, int nu, const void* w_fused, const float* bias,
                        int out_mode, void* out, void* stream);

/* Similarity + stem + the max-pool that follows it, in one kernel: what ResNetEmbeddings hands to the
 * encoder, `pooler(embedder(x))` = MaxPool2d(kernel 3, stride 2, padding 1) of relu(BN(conv7x7/2)) -- HF
 * modeling_resnet.py:57-78 (embedder :67, pooler :76-77), reached through src/efficient_kws/resnet.py:53.
 * out: bf16 channels-last [pairs, ceil(Ho/2), ceil(Wo/2), 64], 1.8 instead of 7.2 MB per pair at 150x1500; the stem
 * activation itself never reaches HBM.  Bit-identical to kws_maxpool_nhwc applied to kws_sim_stem_ragged's bf16 output
 * (the maximum commutes with the monotone bf16 rounding).  Arguments as kws_sim_stem_ragged (kwd_len may be NULL).
 * More than 12 layers run as channel-group passes whose fp16 partial sums live in `workspace`
 * (kws_sim_stem_pool_workspace_bytes(C, pairs of this call, Tk, Tu) bytes, 16-byte aligned; NULL for C <= 12). */
size_t kws_sim_stem_pool_workspace_bytes(int C, long long pairs, int Tk, int Tu);
int kws_sim_stem_pool(const void* kwd_n, const void* utt_n, const int32_t* kwd_len, int C, int K, int U, int Tk, int Tu,
                      int Dk, int pair_mode, int k0, int nk, int u0, int nu, const void* w_fused, const float* bias,
                      void* out, void* workspace, void* stream);

/* Config #4 (original CB-Whisper classifier): bilinear resize of the layer-wise similarity images of
 * ragged keywords to the classifier's fixed input size (replaces torchvision resize(..., antialias=False)
 * == F.interpolate(bilinear, align_corners=False), src/model/cb_whisper.py:208; dataset twin
 * src/data/dataset.py:312-317).
 *   feat_f32 fp32 [K,U,C,Hs,Ws] from kws_sim (keywords zero-padded to Hs frames)
 *   src_h DEVICE int32 [K]: valid frames of each keyword (NULL: all Hs)
 *   out_f32 fp32 [K,U,C,Ho,Wo] or NULL; out_f16 fp16 [K,U,C,Ho,pitch16] or NULL (kws_stem input)
 *   Requires U*C <= 65535 (whole keywords per launch).                                           */
int kws_resize_bilinear(const float* feat_f32, const int32_t* src_h, int K, int U, int C, int Hs, int Ws, int Ho, int Wo,
                        float* out_f32, void* out_f16, int pitch16, void* stream);

/* Config #4, operand-side resize (src/model/cb_whisper.py:189-210; torchvision resize(antialias=False) ==
 * F.interpolate(bilinear, align_corners=False)).  The bilinear resize is linear and separable and the similarity
 * image is bilinear in its operands, so resize(kwd . utt^T) = (Wy kwd) . (Wx utt)^T:
 *   kws_interp_rows        applies Wx to the utterance frames: x fp32 [B,Cin,T,D] -> out fp16 [C,B,T_out,D], every
 *                          source frame L2-normalised first (like kws_normalize_rows; eps < 0: frames used as
 *                          given, like the reference's plain matmul), D % 8 == 0, D <= 1280
 *   kws_sim_operand        native-resolution similarity as a K-major fp16 operand: kwd_n [C,K,Tk,Dk] (Tk % 16 == 0,
 *                          zero-padded frames), utt_n [C,U,Tu,Dk] -> out fp16 [C, K*U, Tu, Tk], item = k*U + u
 *   kws_resize_row_weights Wy as an operand: out fp16 [C,K,Ho,Hp]; row i holds the two bilinear taps of output row i
 *                          over the src_h[k] valid frames of keyword k (src_h DEVICE int32 [K] or NULL = Hp)
 * and kws_sim_stem(kwd_n = Wy, utt_n = similarity operand, Dk = Hp = 64, KWS_PAIRS_PER_KEYWORD) contracts them and
 * applies the stem: the resized [pairs,C,Ho,Wo] image never exists. */
int kws_interp_rows(const float* x, int B, int Cin, int T, int D, const int32_t* layer_idx, int C, int T_out, float eps,
                    void* out_f16, void* stream);
int kws_sim_operand(const void* kwd_n, const void* utt_n, int C, int K, int U, int Tk, int Tu, int Dk, void* out_f16,
                    void* stream);
int kws_resize_row_weights(const int32_t* src_h, int K, int C, int Hp, int Ho, void* out_f16, void* stream);

/* MaxPool2d(3, stride 2, padding 1) on the channels-last bf16 stem activation: the first op of the ResNet
 * body (HF modeling_resnet.py ResNetEmbeddings.pooler, reached through src/efficient_kws/resnet.py:53).
 * Bit-identical to torch's max_pool2d (max is exact; NaNs propagate).
 *   in  bf16 [N,H,W,C], out bf16 [N,ceil(H/2),ceil(W/2),C]; C % 8 == 0                              */
int kws_maxpool_nhwc(const void* in_bf16, long long N, int H, int W, int C, void* out_bf16, void* stream);

/* ---- scores ---------------------------------------------------------------- */

/* Detection score softmax(logits)[:,1] * hotword_mask (model.py:783-795) and
 * thresholded detections (score >= threshold) for n pairs.
 *   logits fp32 [n,2]; hotword_mask fp32 [n] or NULL; scores fp32 [n];
 *   detections uint8 [n] or NULL                                               */
int kws_scores(const float* logits, const float* hotword_mask, size_t n, float threshold, float* scores,
               uint8_t* detections, void* stream);

/* Per-utterance top-k over the keyword axis with deterministic tie-break
 * (lower global keyword index first); used for the local top-k of a keyword
 * shard and for merging the all-gathered candidates (replaces torch.topk,
 * model.py:523).  Segmented bitonic selection on 64-bit (score, id) keys: every
 * level reads its input once (coalesced over neighbouring utterances), any n_cand.
 *   scores fp32 [n_cand, U] (candidate-major), ids int32 [n_cand, U] or NULL
 *   (NULL: candidate c has global id id_offset + c)
 *   out_scores fp32 [k, U], out_ids int32 [k, U]; k <= 1024; rows beyond the number of
 *   candidates hold (-inf, -1)
 *   workspace: kws_topk_workspace_bytes(n_cand, U, k) bytes, 8-byte aligned (0 bytes / NULL for n_cand <= 2048) */
int kws_topk(const float* scores, const int32_t* ids, int n_cand, int U, int id_offset, int k, float* out_scores,
             int32_t* out_ids, void* workspace, void* stream);
size_t kws_topk_workspace_bytes(int n_cand, int U, int k);

#ifdef __cplusplus
}
#endif
#endif /* KWS_B200_H_ */
