"""Config #4: the original CB-Whisper keyword spotter (12-channel ResNet-50 on
bilinearly resized cosine-similarity images), B200 path.

Mirrors the two reference pieces that make up this variant of the hot path:

* ``CBWhisper._calculate_cosine_similarity_matrices_`` (src/model/cb_whisper.py:189-210;
  dataset twin src/data/dataset.py:311-317): per keyword
  ``matmul(kwd_hs [C,Tk_i,D], utt_hs^T [S,C,D,Tu]) -> [S,C,Tk_i,Tu]`` on pre-normalised hidden
  states, then ``torchvision resize(..., (150, 750), antialias=False)``;
* the consumer ``KWSModel.forward(input_features [G,12,150,750])`` of src/model/model.py:78-93
  (ResNet stem + body + head) and the ``argmax(logits) == 1`` detection rule of
  src/model/cb_whisper.py:128.

Here the ragged keywords are zero-padded into one resident fp16 bank.  Two paths:

* ``similarity_images`` (parity with the reference's intermediate): all keyword x segment similarities through
  the tcgen05 GEMM (``kws_sim``), one resize kernel (``kws_resize_bilinear``) -> the [K,S,C,150,750] images.
* ``CBWKeywordSpotterB200.logits`` (the scoring path): the resized image is never built.  Bilinear resize is
  linear and separable and the similarity is bilinear in its operands, so
  ``resize(kwd . utt^T) = (Wy kwd) . (Wx utt)^T``: the width map is applied to the utterance frames
  (``kws_interp_rows``: 1500 -> 750 frames halves the GEMM), the native-resolution similarity is written as a
  64-wide fp16 operand (``kws_sim_operand``), the height map is an operand of its own
  (``kws_resize_row_weights``), and the fused similarity+stem kernel (``kws_sim_stem``, Dk = 64,
  ``KWS_PAIRS_PER_KEYWORD``) contracts the two and applies the stem.  Keywords longer than 64 frames fall back to
  images + ``kws_stem``.

The ResNet body and head are library code (``body.py`` / the unmodified modules).  CUDA only, inference only.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from .body import FusedBody
from .model import run_body

FUSED_MAX_FRAMES = 64  # keyword frames the fused path holds as the contraction dimension of the height map


NO_NORM = -1.0  # eps < 0 for kws_normalize_rows / kws_interp_rows: hidden states are used as given


def pack_keywords(kwd_list: Sequence[torch.Tensor], device: torch.device, multiple: int = 16):
    """Ragged pre-normalised keywords [C, Tk_i, D] -> (fp16 operand bank [C, K, Tkp, D], int32 lengths [K]).
    Rows beyond a keyword's length are zero (they produce zero similarity and are never sampled by the
    resize).  Like the reference (plain ``torch.matmul`` on whatever it is given, cb_whisper.py:197) the states
    are NOT re-normalised: the caller normalises them once (cb_whisper.py:106, src/utils.py:195).  Operands are
    fp16: states far outside [-1, 1] lose precision, unit vectors keep the similarity within ~3e-4."""
    if len(kwd_list) == 0:
        raise ops.KWSError("empty keyword list")
    C, _, D = kwd_list[0].shape
    lens = [int(k.shape[1]) for k in kwd_list]
    if min(lens) < 1:
        raise ops.KWSError("keywords need at least one frame")
    Tkp = (max(lens) + multiple - 1) // multiple * multiple
    bank = torch.zeros((len(kwd_list), C, Tkp, D), dtype=torch.float32, device=device)
    for i, k in enumerate(kwd_list):
        if k.shape[0] != C or k.shape[2] != D:
            raise ops.KWSError(f"keyword {i} has shape {tuple(k.shape)}, expected [{C}, T, {D}]")
        bank[i, :, : lens[i]] = k.to(device=device, dtype=torch.float32)
    # layer-major fp16 layout, values as given (eps < 0: no normalisation)
    kwd_n = ops.normalize_rows(bank, list(range(C)), None, eps=NO_NORM)
    return kwd_n, torch.tensor(lens, dtype=torch.int32, device=device)


def similarity_images(kwd_list: Sequence[torch.Tensor], utt_hs: torch.Tensor, size: Tuple[int, int] = (150, 750),
                      want_f32: bool = True, want_f16: bool = False):
    """Replacement of ``_calculate_cosine_similarity_matrices_``.

    kwd_list: K ragged keywords [C, Tk_i, D]; utt_hs: [S, C, Tu, D] (both L2-normalised over D).
    Returns (fp32 [K, S, C, size0, size1] | None, fp16 [K, S, C, size0, pitch] | None).  (The reference
    returns a per-segment list of [K, C, size0, size1]; index the result with ``[:, s]``.)"""
    if not utt_hs.is_cuda:
        raise ops.KWSError("utt_hs must be a CUDA tensor (the kws_b200 path has no CPU fallback)")
    dev = utt_hs.device
    S, C, Tu, D = utt_hs.shape
    kwd_n, lens = pack_keywords(kwd_list, dev)
    utt_n = ops.normalize_rows(utt_hs.float().contiguous(), list(range(C)), None, eps=NO_NORM)
    K = kwd_n.shape[1]
    # bound the fp32 intermediate [kb, S, C, Tkp, Tu]
    per_kw = S * C * kwd_n.shape[2] * Tu * 4
    kb = max(1, min(K, (2 << 30) // max(per_kw, 1), 65535 // (S * C)))
    o32, o16 = [], []
    for k0 in range(0, K, kb):
        k1 = min(K, k0 + kb)
        f32, _ = ops.sim(kwd_n[:, k0:k1].contiguous(), utt_n, want_f32=True, want_f16=False)
        a, b = ops.resize_bilinear(f32, lens[k0:k1].contiguous(), size, want_f32=want_f32, want_f16=want_f16)
        o32.append(a)
        o16.append(b)
    cat = lambda xs: (xs[0] if len(xs) == 1 else torch.cat(xs)) if xs[0] is not None else None
    return cat(o32), cat(o16)


class CBWKeywordSpotterB200:
    """12-channel ResNet classifier on resized similarity images: the B200 twin of the original
    ``KWSModel`` (src/model/model.py:17-93) as driven by ``CBWhisper.keyword_spotting``
    (src/model/cb_whisper.py:110-128).  ``resnet`` is a module with the reference wrapper's attribute names
    (``feature_extractor``, ``classifier``), e.g. ``enhance_cb_whisper_b200.Resnet(12, 2)`` or the
    reference's own ``model.resnet.Resnet`` carrying a trained checkpoint."""

    def __init__(self, resnet: torch.nn.Module, size: Tuple[int, int] = (150, 750), body_dtype: str = "float32"):
        self.resnet = resnet.eval()
        self.size = tuple(size)
        self.body_dtype = body_dtype
        self.fused_pool = True  # bf16 body: the stem's max-pool runs inside the fused kernel (kws_sim_stem_pool)
        self._packed = None
        self._packed_fused = None
        self._lowp = None

    def _weights(self, device):
        if self._packed is None or self._packed[0].device != device:
            emb = self.resnet.feature_extractor.embedder.embedder
            self._packed = ops.pack_stem_weights(emb.convolution.weight.to(device), emb.normalization.weight.to(device),
                                                 emb.normalization.bias.to(device),
                                                 emb.normalization.running_mean.to(device),
                                                 emb.normalization.running_var.to(device))
        return self._packed

    def _weights_fused(self, device):
        if self._packed_fused is None or self._packed_fused[0].device != device:
            emb = self.resnet.feature_extractor.embedder.embedder
            self._packed_fused = ops.pack_stem_fused(emb.convolution.weight.to(device), emb.normalization.weight.to(device),
                                                     emb.normalization.bias.to(device),
                                                     emb.normalization.running_mean.to(device),
                                                     emb.normalization.running_var.to(device))
        return self._packed_fused

    def _body(self, st, pooled: bool = False):
        if self.body_dtype == "float32":
            return run_body(self.resnet, st)
        if self._lowp is None:  # BatchNorms folded + cuDNN fused convolutions (body.py), max-pool in libkws_b200
            self._lowp = FusedBody(self.resnet, torch.bfloat16)
        return self._lowp(st if pooled else ops.maxpool_nhwc(st), pooled=True)

    @torch.no_grad()
    def stem_fused(self, kwd_n: torch.Tensor, lens: torch.Tensor, utt_i: torch.Tensor, out_mode: int,
                   max_pairs: int = 256, consume=None):
        """Scoring path up to the stem activation, resized image never built.  kwd_n fp16 [C,K,64,D] (pack_keywords),
        lens int32 [K], utt_i fp16 [C,S,Wi,D] (ops.interp_rows to the image width).  Calls
        ``consume(k0, k1, stem_activation [(k1-k0)*S, 64, Ho, Wo])`` per block of keywords (``out_mode``
        STEM_OUT_POOL_NHWC_BF16: the max-pooled activation [(k1-k0)*S, 64, ceil(Ho/2), ceil(Wo/2)])."""
        C, K, Tkp, _ = kwd_n.shape
        S = utt_i.shape[1]
        wf, bias = self._weights_fused(kwd_n.device)
        kb = max(1, min(K, max_pairs // max(S, 1)))
        for k0 in range(0, K, kb):
            k1 = min(K, k0 + kb)
            s_op = ops.sim_operand(kwd_n[:, k0:k1].contiguous() if (k0, k1) != (0, K) else kwd_n, utt_i)
            wy = ops.resize_row_weights(lens[k0:k1].contiguous(), k1 - k0, C, Tkp, self.size[0])
            if out_mode == ops.STEM_OUT_POOL_NHWC_BF16:  # + the max-pool behind the stem, in the same kernel
                st = ops.sim_stem_pool(wy, s_op, wf, bias, per_keyword=True)
            else:
                st = ops.sim_stem(wy, s_op, wf, bias, out_mode, per_keyword=True)
            if consume is not None:
                consume(k0, k1, st)

    def fused_ok(self, kwd_list, utt_hs) -> bool:
        C = utt_hs.shape[1]
        return (max(int(k.shape[1]) for k in kwd_list) <= FUSED_MAX_FRAMES and utt_hs.shape[3] % 64 == 0
                and utt_hs.shape[3] <= 1280 and bool(ops.sim_stem_supported(C, self.size[0], self.size[1], 64,
                                                                            ops.STEM_OUT_NHWC_BF16 if self.body_dtype != "float32"
                                                                            else ops.STEM_OUT_NCHW_F32)))

    @torch.no_grad()
    def logits(self, kwd_list: Sequence[torch.Tensor], utt_hs: torch.Tensor, max_pairs: int = 128,
               fused: Optional[bool] = None) -> torch.Tensor:
        """-> logits fp32 [K, S, 2] for every keyword x segment.  ``fused``: None = use the fused path when the
        shapes allow (keywords <= 64 frames), False = images + un-fused stem, True = insist."""
        dev = utt_hs.device
        if not utt_hs.is_cuda:
            raise ops.KWSError("utt_hs must be a CUDA tensor (the kws_b200 path has no CPU fallback)")
        lowp = self.body_dtype != "float32"
        out_mode = ops.STEM_OUT_NHWC_BF16 if lowp else ops.STEM_OUT_NCHW_F32
        use_fused = self.fused_ok(kwd_list, utt_hs) if fused is None else fused
        if use_fused:
            if not self.fused_ok(kwd_list, utt_hs):
                raise ops.KWSError("fused config-#4 path needs keywords <= 64 frames, D % 64 == 0, D <= 1280")
            S, C = utt_hs.shape[0], utt_hs.shape[1]
            kwd_n, lens = pack_keywords(kwd_list, dev, multiple=FUSED_MAX_FRAMES)
            utt_i = ops.interp_rows(utt_hs.float().contiguous(), list(range(C)), self.size[1], eps=NO_NORM)
            K = kwd_n.shape[1]
            out = torch.empty((K, S, 2), dtype=torch.float32, device=dev)

            pooled = lowp and self.fused_pool
            def consume(k0, k1, st):
                out[k0:k1] = self._body(st, pooled).float().view(k1 - k0, S, 2)

            self.stem_fused(kwd_n, lens, utt_i, ops.STEM_OUT_POOL_NHWC_BF16 if pooled else out_mode, max_pairs, consume)
            return out
        _, f16 = similarity_images(kwd_list, utt_hs, self.size, want_f32=False, want_f16=True)
        K, S = f16.shape[:2]
        wp, bias = self._weights(dev)
        flat = f16.view(K * S, *f16.shape[2:])
        out = torch.empty((K * S, 2), dtype=torch.float32, device=dev)
        for p0 in range(0, K * S, max_pairs):
            p1 = min(K * S, p0 + max_pairs)
            st = ops.stem(flat[p0:p1], self.size[1], wp, bias, out_mode)
            out[p0:p1] = self._body(st).float()
        return out.view(K, S, 2)

    @torch.no_grad()
    def keyword_spotting(self, utt_hs: Optional[torch.Tensor], kw_groups, n_segments: Optional[int] = None,
                         max_pairs: int = 128) -> List[List[str]]:
        """The ``oracle == 'kws'`` branch of ``CBWhisper.keyword_spotting`` (src/model/cb_whisper.py:96-131) with
        the similarity / resize / classifier on the B200 path.

        utt_hs: [S, C, Tu, D] normalised encoder hidden states of the S segments (cb_whisper.py:100-106), or None
        when feature extraction failed (every segment then retrieves nothing, :113-116).  kw_groups: iterable of
        keyword-database groups, each a mapping with ``'hidden_states'`` (list of [C, Tk_i, D]) and ``'keywords'``
        (list of str) -- what ``kw_database.group(idx)`` returns.  Returns, per segment, the de-duplicated
        retrieved keywords (``argmax(logits) == 1``, :126-131); order within a segment is unspecified, as in the
        reference (``list(set(...))``)."""
        S = utt_hs.shape[0] if utt_hs is not None else int(n_segments or 1)
        keywords: List[List[str]] = [[] for _ in range(S)]
        if utt_hs is None:
            return keywords
        for group in kw_groups:
            hs = group["hidden_states"]
            if len(hs) == 0:
                continue
            hit = self.logits(hs, utt_hs, max_pairs=max_pairs).argmax(dim=-1) == 1  # [K,S]
            hit = hit.cpu()
            for s in range(S):
                keywords[s] += [group["keywords"][i] for i in torch.nonzero(hit[:, s]).flatten().tolist()]
        return [list(set(k)) for k in keywords]

    @torch.no_grad()
    def detect(self, kwd_list: Sequence[torch.Tensor], utt_hs: torch.Tensor) -> List[List[int]]:
        """Per segment, the indices of the keywords with ``argmax(logits) == 1`` (cb_whisper.py:128)."""
        lg = self.logits(kwd_list, utt_hs)
        hit = lg.argmax(dim=-1) == 1  # [K,S]
        return [torch.nonzero(hit[:, s]).flatten().tolist() for s in range(hit.shape[1])]
