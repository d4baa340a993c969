"""Host-side driver of the B200 hot path: weight packing and the batched
compress -> similarity -> stem pipeline over K keywords x U utterances.

This is the batched replacement of the per-group Python loop in the reference's
``test_step`` (src/efficient_kws/model.py:748-802): keywords are compressed
once into a resident fp16 bank, every utterance is compressed once, and all
K x U pairs stream through the similarity GEMM and the stem in pair chunks.
All arithmetic is in libkws_b200.so (see ``ops``); nothing here computes on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Iterator, Mapping, Optional, Sequence, Tuple

import torch

from . import ops

STEM_KEY = "model.feature_extractor.embedder.embedder."


@dataclass
class PackedWeights:
    """Device-resident, kernel-ready weights of one checkpoint (inference only:
    BatchNorm layers are folded with their running statistics)."""

    variant: str  # "L" | "LE" | "LEF"
    C: int
    D: int
    P: int
    mlp_dtype16: int = ops.F16  # operand type of the projector GEMMs (ops.F16 | ops.BF16)
    w1: Optional[torch.Tensor] = None  # 16-bit [C,H,D]
    b1: Optional[torch.Tensor] = None  # fp32 [C,H]
    w2: Optional[torch.Tensor] = None  # 16-bit [C,P,H]
    b2: Optional[torch.Tensor] = None  # fp32 [C,P]
    wt: Optional[torch.Tensor] = None  # 16-bit [C,3,P/8,P,8]  temporal conv, BN folded, packed for kws_temporal
    bt: Optional[torch.Tensor] = None  # fp32 [C,P]
    stem_w: Optional[torch.Tensor] = None  # fp16 [G,49,2,64,8]
    stem_b: Optional[torch.Tensor] = None  # fp32 [64]
    stem_wf: Optional[torch.Tensor] = None  # fp16, fused-kernel packing (C <= 12), else None

    @property
    def Dk(self) -> int:
        return self.D if self.variant == "L" else self.P


def pack_weights(sd: Mapping[str, torch.Tensor], variant: str, C: int, D: int, P: int,
                 device: torch.device, mlp_dtype16: int = ops.F16) -> PackedWeights:
    """Pack a reference-keyed state_dict (projector.{i}.{0,2}.*, time_projector.{i}.{0,1}.*,
    model.feature_extractor.embedder.embedder.*) for the kernels.

    ``mlp_dtype16``: fp16 (default) keeps the similarity of the LE/LEF variants within the 2e-3
    parity bound -- the reference's inputs are L2-normalised (src/utils.py:195) so |x| <= 1 and
    |hidden| <= ||W1_j|| + |b1_j|, far inside fp16 range; conversions saturate instead of
    overflowing.  bf16 is range-safe for arbitrary activations at ~3e-3 similarity error."""
    def dev(k):
        return sd[k].detach().to(device=device, dtype=torch.float32)

    pw = PackedWeights(variant=variant, C=C, D=D, P=P, mlp_dtype16=mlp_dtype16)
    if variant in ("LE", "LEF"):
        pw.w1 = ops.cast16(torch.stack([dev(f"projector.{i}.0.weight") for i in range(C)]), mlp_dtype16)
        pw.b1 = torch.stack([dev(f"projector.{i}.0.bias") for i in range(C)]).contiguous()
        pw.w2 = ops.cast16(torch.stack([dev(f"projector.{i}.2.weight") for i in range(C)]), mlp_dtype16)
        pw.b2 = torch.stack([dev(f"projector.{i}.2.bias") for i in range(C)]).contiguous()
    if variant == "LEF":
        st = lambda name: torch.stack([dev(f"time_projector.{i}.{name}") for i in range(C)])
        pw.wt, pw.bt = ops.fold_temporal_weights(st("0.weight"), st("0.bias"), st("1.weight"), st("1.bias"),
                                                 st("1.running_mean"), st("1.running_var"), dtype16=mlp_dtype16)
    if STEM_KEY + "convolution.weight" in sd:
        pw.stem_w, pw.stem_b = ops.pack_stem_weights(
            dev(STEM_KEY + "convolution.weight"), dev(STEM_KEY + "normalization.weight"),
            dev(STEM_KEY + "normalization.bias"), dev(STEM_KEY + "normalization.running_mean"),
            dev(STEM_KEY + "normalization.running_var"))
        if ops._lib.load().kws_stem_fused_weight_bytes(C):
            pw.stem_wf, _ = ops.pack_stem_fused(
                dev(STEM_KEY + "convolution.weight"), dev(STEM_KEY + "normalization.weight"),
                dev(STEM_KEY + "normalization.bias"), dev(STEM_KEY + "normalization.running_mean"),
                dev(STEM_KEY + "normalization.running_var"))
    return pw


class KWSEngine:
    """compress / similarity / stem over batches, with bounded workspaces."""

    def __init__(self, weights: PackedWeights, workspace_bytes: int = 8 << 30, fused_mlp: bool = True):
        self.w = weights
        self.workspace_bytes = int(workspace_bytes)
        self.fused_mlp = bool(fused_mlp)  # kws_mlp_fused where the shape allows; False: kws_cast_rows16 + kws_mlp

    # -- stage 1: per-layer compression -> normalised fp16 operands [C,B,T',Dk] ------
    def out_frames(self, T: int) -> int:
        return (T + 1) // 2 if self.w.variant == "LEF" else T

    def compress(self, x: torch.Tensor, mask: Optional[torch.Tensor],
                 layer_idx: Optional[Sequence[int]] = None) -> torch.Tensor:
        """x fp32 [B,Cin,T,D]; mask fp32 [B,C,T'] at the resolution the similarity sees
        (T for L/LE, ceil(T/2) for LEF) or None -> fp16 [C,B,T',Dk]."""
        w = self.w
        B, Cin, T, D = x.shape
        if D != w.D:
            raise ops.KWSError(f"embedding dim {D} != model embedding_dim {w.D}")
        if layer_idx is None:
            layer_idx = list(range(w.C))
        if len(layer_idx) != w.C:
            raise ops.KWSError(f"need {w.C} layer indices, got {len(layer_idx)}")
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        if mask is not None:
            want = (B, w.C, self.out_frames(T))
            if tuple(mask.shape) != want:
                # e.g. the [B,12,T] masks of the eval datasets without the [-n_layers:] slice, or a full-resolution
                # mask for LEF: the kernels index mask[(b*C + c)*T' + t], so a wrong shape must not get through
                raise ops.KWSError(f"mask must be [B,C,T']={want} for the {w.variant} variant, got {tuple(mask.shape)}")
            mask = mask.to(torch.float32).contiguous()
        if w.variant == "L":
            return ops.normalize_rows(x, layer_idx, mask)
        T2 = self.out_frames(T)
        H = w.w1.shape[1]
        if self.fused_mlp and ops.mlp_fused_supported(D, H, w.P):
            # one kernel from the raw fp32 rows to the compressed operands: no 16-bit copy of x, no hidden tensor
            if w.variant == "LE":
                return ops.mlp_fused(x, layer_idx, w.w1, w.b1, w.w2, w.b2, mask, ops.MLP_OUT_NORM_F16)
            proj = ops.mlp_fused(x, layer_idx, w.w1, w.b1, w.w2, w.b2, None, ops.MLP_OUT_RAW_16)
            return ops.temporal(proj, w.wt, w.bt, mask)
        out = torch.empty((w.C, B, T2, w.P), dtype=torch.float16, device=x.device)
        per_item = w.C * T * (D + H) * 2 + w.C * T * w.P * 2
        step = max(1, min(B, self.workspace_bytes // max(per_item, 1)))
        for b0 in range(0, B, step):
            b1 = min(B, b0 + step)
            xb = ops.cast_rows16(x[b0:b1], layer_idx, w.mlp_dtype16)
            mb = mask[b0:b1].contiguous() if mask is not None else None
            if w.variant == "LE":
                o = ops.mlp(xb, b1 - b0, T, w.w1, w.b1, w.w2, w.b2, mb, ops.MLP_OUT_NORM_F16)
            else:
                proj = ops.mlp(xb, b1 - b0, T, w.w1, w.b1, w.w2, w.b2, None, ops.MLP_OUT_RAW_16)
                o = ops.temporal(proj, w.wt, w.bt, mb)
            if step >= B:
                return o
            out[:, b0:b1] = o
        return out

    @staticmethod
    def keyword_lengths(mask: torch.Tensor) -> torch.Tensor:
        """mask [K,C,T'] (0/1 per frame, any layer) -> int32 [K]: 1 + index of the last frame that is unmasked in any
        layer (0 if none).  Frames at or beyond it are zero rows of the compressed bank whatever the variant, because
        the mask is folded in as a row scale."""
        live = (mask != 0).any(dim=1)  # [K,T']
        T = live.shape[1]
        idx = torch.arange(1, T + 1, device=mask.device, dtype=torch.int32)
        return (live.to(torch.int32) * idx).amax(dim=1).to(torch.int32).contiguous()

    # -- stage 2+3 over pair chunks -----------------------------------------------------
    def pair_chunks(self, K: int, U: int, Tk: int, Tu: int, max_pairs: int) -> Iterator[Tuple[int, int, int, int]]:
        """Tile the K x U pair grid into (k0,k1,u0,u1) blocks of at most max_pairs pairs.  Blocks are
        keyword-major (a run of keywords against a few utterances) so that the utterance tiles, the larger
        operand, stay L2-resident while the keyword rows stream."""
        max_pairs = max(1, max_pairs)
        ub = max(1, min(U, max_pairs // max(1, min(K, max_pairs))))
        kb = max(1, min(K, max_pairs // ub))
        for u0 in range(0, U, ub):
            for k0 in range(0, K, kb):
                yield k0, min(K, k0 + kb), u0, min(U, u0 + ub)

    def fused(self, Tk: int, Tu: int, out_mode: int = ops.STEM_OUT_NHWC_BF16) -> bool:
        """True when the fused similarity+stem kernel covers this model and output mode
        (kws_sim_stem_supported; more than 12 layers run as channel-group passes, bf16 output only)."""
        return self.w.stem_wf is not None and ops.sim_stem_supported(self.w.C, Tk, Tu, self.w.Dk, out_mode)

    def hot_path(self, kwd_n: torch.Tensor, utt_n: torch.Tensor, out_mode: int, max_pairs: int = 1024,
                 consume: Optional[Callable] = None, bufs: Optional[dict] = None,
                 launch_events: Optional[list] = None, kwd_len: Optional[torch.Tensor] = None):
        """Similarity + stem for all pairs, chunked.  ``consume(k0,k1,u0,u1,stem_out)`` receives the
        stem activation of each chunk ([pairs,64,Ho,Wo], pair = (k-k0)*(u1-u0) + (u-u0));
        intermediate buffers are reused across chunks (``bufs``).  The fused kernel is used whenever it
        supports the shape (the similarity tensor then never exists in HBM); otherwise kws_sim + kws_stem.
        ``launch_events``: if a list, a (start, end) CUDA-event pair around every pair-kernel launch is
        appended (bench.py reads per-launch durations from them).  ``kwd_len``: int32 [K] valid frames of every
        keyword at the similarity's resolution (``keyword_lengths``): the fused kernel then skips the rows beyond a
        keyword (bit-identical output); ignored by the un-fused path.
        ``out_mode`` STEM_OUT_POOL_NHWC_BF16: ``consume`` receives the MAX-POOLED stem activation
        ([pairs,64,ceil(Ho/2),ceil(Wo/2)] bf16 channels_last, what ResNetEmbeddings hands to the encoder) from the fused
        similarity + stem + pool kernel; the stem activation itself never reaches HBM (un-fused shapes: kws_sim ->
        kws_stem -> kws_maxpool_nhwc)."""
        Cc, K, Tk, Dk = kwd_n.shape
        _, U, Tu, _ = utt_n.shape
        bufs = bufs if bufs is not None else {}
        fused = self.fused(Tk, Tu, out_mode)
        Ho, Wo = (Tk + 1) // 2, (Tu + 1) // 2
        f32 = out_mode == ops.STEM_OUT_NCHW_F32
        pool = out_mode == ops.STEM_OUT_POOL_NHWC_BF16
        if pool and fused:
            Ho, Wo = (Ho + 1) // 2, (Wo + 1) // 2  # the buffer holds the pooled activation
        n = 0
        for k0, k1, u0, u1 in self.pair_chunks(K, U, Tk, Tu, max_pairs):
            np_ = (k1 - k0) * (u1 - u0)
            keyo = ("stem", max_pairs, Tk, Tu, out_mode)
            if keyo not in bufs or bufs[keyo].numel() < np_ * 64 * Ho * Wo:
                cap = max(np_, min(max_pairs, K * U))
                bufs[keyo] = torch.empty(cap * 64 * Ho * Wo, dtype=torch.float32 if f32 else torch.bfloat16,
                                         device=kwd_n.device)
            if launch_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            if fused and pool:
                need = ops.sim_stem_pool_workspace_bytes(Cc, np_, Tk, Tu)
                if need and ("pool_ws" not in bufs or bufs["pool_ws"].numel() < need):
                    bufs["pool_ws"] = torch.empty(need, dtype=torch.uint8, device=kwd_n.device)
                st = ops.sim_stem_pool(kwd_n, utt_n, self.w.stem_wf, self.w.stem_b, out=bufs[keyo], k_range=(k0, k1),
                                       u_range=(u0, u1), kwd_len=kwd_len, workspace=bufs.get("pool_ws"))
            elif fused:
                st = ops.sim_stem(kwd_n, utt_n, self.w.stem_wf, self.w.stem_b, out_mode, out=bufs[keyo],
                                  k_range=(k0, k1), u_range=(u0, u1), kwd_len=kwd_len)
            else:
                kk = kwd_n[:, k0:k1].contiguous() if (k0, k1) != (0, K) else kwd_n
                uu = utt_n[:, u0:u1].contiguous() if (u0, u1) != (0, U) else utt_n
                key16 = ("f16", np_, Cc, Tk, Tu)
                if key16 not in bufs:
                    bufs[key16] = torch.empty((k1 - k0, u1 - u0, Cc, Tk, ops.pitch_for(Tu)), dtype=torch.float16,
                                              device=kwd_n.device)
                _, f16 = ops.sim(kk, uu, want_f32=False, want_f16=True, out_f16=bufs[key16])
                shape = (np_, 64, Ho, Wo) if f32 else (np_, Ho, Wo, 64)
                st = ops.stem(f16, Tu, self.w.stem_w, self.w.stem_b, ops.STEM_OUT_NHWC_BF16 if pool else out_mode,
                              out=bufs[keyo][: np_ * 64 * Ho * Wo].view(shape))
                if pool:
                    st = ops.maxpool_nhwc(st)
            if launch_events is not None:
                ev[1].record()
                launch_events.append(ev + (np_,))
            if consume is not None:
                consume(k0, k1, u0, u1, st)
            n += np_
        return n
