"""Torch-tensor front end of the C ABI: allocates outputs with the caching
allocator, checks shapes/devices, passes raw device pointers and the current
CUDA stream to libkws_b200.so.  PyTorch is plumbing here (memory + streams);
all arithmetic happens in the CUDA library.  CPU tensors are rejected -- there
is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (BF16, F16, MLP_OUT_NORM_F16, MLP_OUT_RAW_16, MLP_OUT_RAW_F32, PAIRS_ALL, PAIRS_DIAG, PAIRS_PER_KEYWORD,
                   STEM_OUT_NCHW_F32,
                   STEM_OUT_NHWC_BF16, STEM_OUT_POOL_NHWC_BF16, KWSError)

TORCH16 = {F16: torch.float16, BF16: torch.bfloat16}

SIM_EPS = 1e-6  # reference src/efficient_kws/model.py:210
BN_EPS = 1e-5


def _cuda(t: Optional[torch.Tensor], name: str, dtype=None) -> int:
    if t is None:
        return 0
    if not t.is_cuda:
        raise KWSError(f"{name} must be a CUDA tensor (the kws_b200 path has no CPU fallback)")
    if not t.is_contiguous():
        raise KWSError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise KWSError(f"{name} must be {dtype}, got {t.dtype}")
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _guard(fn):
    """Run an op on the device of its first CUDA tensor argument: the library launches on the CURRENT device and
    ``_stream()`` is that device's current stream, so tensors living on another GPU would otherwise be handed to
    kernels on the wrong device/stream."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        if dev is None and isinstance(kwargs.get("device"), (torch.device, str)):
            d = torch.device(kwargs["device"])
            dev = d if d.type == "cuda" else None
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapped


# kernels launched through this module since import (bench.py reports the count of the timed region)
LAUNCHES = 0
_KERNELS_PER_CALL = {"kws_mlp": 2}


def check(rc: int, what: str, launches: Optional[int] = None) -> None:  # _lib.check + launch accounting
    global LAUNCHES
    _lib.check(rc, what)
    LAUNCHES += _KERNELS_PER_CALL.get(what, 1) if launches is None else launches


def _layers(layer_idx: Sequence[int]):
    arr = (C.c_int32 * len(layer_idx))(*[int(i) for i in layer_idx])
    return arr


def sm_count() -> int:
    return _lib.load().kws_sm_count()


# ---- once per checkpoint -----------------------------------------------------
@_guard
def pack_stem_weights(conv_w, gamma, beta, mean, var, eps: float = BN_EPS) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (w_packed fp16 [G,49,2,64,8], bias fp32 [64])"""
    lib = _lib.load()
    Cc = conv_w.shape[1]
    if tuple(conv_w.shape) != (64, Cc, 7, 7):
        raise KWSError(f"stem conv weight must be [64,C,7,7], got {tuple(conv_w.shape)}")
    args = [t.detach().float().contiguous() for t in (conv_w, gamma, beta, mean, var)]
    G = (Cc + 15) // 16
    wp = torch.empty((G, 49, 2, 64, 8), dtype=torch.float16, device=conv_w.device)
    bias = torch.empty(64, dtype=torch.float32, device=conv_w.device)
    check(lib.kws_pack_stem_weights(*[_cuda(a, "stem weight", torch.float32) for a in args], eps, Cc,
                                    _cuda(wp, "w_packed"), _cuda(bias, "bias"), _stream()), "kws_pack_stem_weights")
    return wp, bias


@_guard
def pack_stem_fused(conv_w, gamma, beta, mean, var, eps: float = BN_EPS) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stem conv + BN folded and packed for the fused similarity+stem kernel
    -> (w_fused fp16 [kws_stem_fused_weight_bytes(C) / 2], bias fp32 [64]); C <= 12."""
    lib = _lib.load()
    Cc = conv_w.shape[1]
    if tuple(conv_w.shape) != (64, Cc, 7, 7):
        raise KWSError(f"stem conv weight must be [64,C,7,7], got {tuple(conv_w.shape)}")
    nbytes = lib.kws_stem_fused_weight_bytes(Cc)
    if nbytes == 0:
        raise KWSError(f"the fused similarity+stem kernel does not cover C={Cc}")
    args = [t.detach().float().contiguous() for t in (conv_w, gamma, beta, mean, var)]
    wf = torch.empty(nbytes // 2, dtype=torch.float16, device=conv_w.device)
    bias = torch.empty(64, dtype=torch.float32, device=conv_w.device)
    check(lib.kws_pack_stem_fused(*[_cuda(a, "stem weight", torch.float32) for a in args], eps, Cc,
                                  _cuda(wf, "w_fused"), _cuda(bias, "bias"), _stream()), "kws_pack_stem_fused")
    return wf, bias


@_guard
def fold_temporal_weights(conv_w, conv_b, gamma, beta, mean, var, eps: float = BN_EPS, dtype16: int = F16):
    """conv_w [C,P,P,3] ... -> (w16 packed fp16|bf16 [C,3,P/8,P,8], b_folded fp32 [C,P])"""
    lib = _lib.load()
    Cc, P = conv_w.shape[0], conv_w.shape[1]
    args = [t.detach().float().contiguous() for t in (conv_w, conv_b, gamma, beta, mean, var)]
    wf = torch.empty((Cc, 3, P // 8, P, 8), dtype=TORCH16[dtype16], device=conv_w.device)
    bf = torch.empty((Cc, P), dtype=torch.float32, device=conv_w.device)
    check(lib.kws_fold_temporal_weights(*[_cuda(a, "temporal weight", torch.float32) for a in args], eps, Cc, P,
                                        dtype16, _cuda(wf, "wf"), _cuda(bf, "bf"), _stream()),
          "kws_fold_temporal_weights")
    return wf, bf


@_guard
def cast16(src: torch.Tensor, dtype16: int = F16) -> torch.Tensor:
    """fp32 -> fp16 (saturating) | bf16"""
    lib = _lib.load()
    src = src.detach().float().contiguous()
    dst = torch.empty(src.shape, dtype=TORCH16[dtype16], device=src.device)
    check(lib.kws_cast_f32_to_16(_cuda(src, "src", torch.float32), _cuda(dst, "dst"), src.numel(), dtype16,
                                 _stream()), "kws_cast_f32_to_16")
    return dst


# ---- per batch of keywords / utterances ----------------------------------------
@_guard
def normalize_rows(x: torch.Tensor, layer_idx: Sequence[int], mask: Optional[torch.Tensor],
                   eps: float = SIM_EPS) -> torch.Tensor:
    """x fp32 [B,Cin,T,D] -> fp16 [C,B,T,D] (selected layers, L2-normalised, mask folded)."""
    lib = _lib.load()
    B, Cin, T, D = x.shape
    Cc = len(layer_idx)
    if mask is not None and tuple(mask.shape) != (B, Cc, T):
        raise KWSError(f"mask must be [B,C,T]=({B},{Cc},{T}), got {tuple(mask.shape)}")
    out = torch.empty((Cc, B, T, D), dtype=torch.float16, device=x.device)
    check(lib.kws_normalize_rows(_cuda(x, "x", torch.float32), B, Cin, T, D, _layers(layer_idx), Cc,
                                 _cuda(mask, "mask", torch.float32), eps, _cuda(out, "out"), _stream()),
          "kws_normalize_rows")
    return out


@_guard
def cast_rows16(x: torch.Tensor, layer_idx: Sequence[int], dtype16: int = F16) -> torch.Tensor:
    """x fp32 [B,Cin,T,D] -> fp16|bf16 [C, B*T, D] (selected layers, layer-major rows)."""
    lib = _lib.load()
    B, Cin, T, D = x.shape
    Cc = len(layer_idx)
    out = torch.empty((Cc, B * T, D), dtype=TORCH16[dtype16], device=x.device)
    check(lib.kws_cast_rows16(_cuda(x, "x", torch.float32), B, Cin, T, D, _layers(layer_idx), Cc, dtype16,
                              _cuda(out, "out"), _stream()), "kws_cast_rows16")
    return out


@_guard
def mlp(x16: torch.Tensor, B: int, T: int, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
        b2: torch.Tensor, mask: Optional[torch.Tensor], out_mode: int, eps: float = SIM_EPS) -> torch.Tensor:
    """x [C,B*T,D]; w1 [C,H,D]; w2 [C,P,H] all fp16 or all bf16; b1 fp32 [C,H]; b2 fp32 [C,P]
    -> fp16 [C,B,T,P] (normalised), fp32 [C,B,T,P] (raw) or the operands' 16-bit type (raw, feeds temporal)."""
    lib = _lib.load()
    x_bf16 = x16
    dt = x16.dtype
    if dt not in (torch.float16, torch.bfloat16) or w1.dtype != dt or w2.dtype != dt:
        raise KWSError(f"x/w1/w2 must share one 16-bit dtype, got {x16.dtype}/{w1.dtype}/{w2.dtype}")
    dtype16 = F16 if dt == torch.float16 else BF16
    Cc, R, D = x_bf16.shape
    H, P = w1.shape[1], w2.shape[1]
    if R != B * T:
        raise KWSError(f"x rows {R} != B*T = {B * T}")
    if tuple(w1.shape) != (Cc, H, D) or tuple(w2.shape) != (Cc, P, H):
        raise KWSError("projector weight shapes do not match x")
    if mask is not None and tuple(mask.shape) != (B, Cc, T):
        # the epilogue indexes mask[(b*C + c)*T + t]: any other shape would be mis-indexed or read out of bounds
        # (the reference raises a broadcast error, model.py:187-191)
        raise KWSError(f"mask must be [B,C,T]=({B},{Cc},{T}), got {tuple(mask.shape)}")
    hidden = torch.empty((Cc, R, H), dtype=dt, device=x_bf16.device)
    out_dt = {MLP_OUT_NORM_F16: torch.float16, MLP_OUT_RAW_F32: torch.float32, MLP_OUT_RAW_16: dt}[out_mode]
    out = torch.empty((Cc, B, T, P), dtype=out_dt, device=x_bf16.device)
    check(lib.kws_mlp(_cuda(x_bf16, "x", dt), Cc, B, T, D, H, P, dtype16, _cuda(w1, "w1", dt),
                      _cuda(b1, "b1", torch.float32), _cuda(w2, "w2", dt),
                      _cuda(b2, "b2", torch.float32), _cuda(hidden, "hidden"),
                      _cuda(mask, "mask", torch.float32), eps, out_mode, _cuda(out, "out"), _stream()), "kws_mlp")
    return out


def mlp_fused_supported(D: int, H: int, P: int) -> bool:
    return bool(_lib.load().kws_mlp_fused_supported(int(D), int(H), int(P)))


@_guard
def mlp_fused(x: torch.Tensor, layer_idx: Sequence[int], w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
              b2: torch.Tensor, mask: Optional[torch.Tensor], out_mode: int, eps: float = SIM_EPS) -> torch.Tensor:
    """The per-layer projector as one kernel on the raw embeddings: x fp32 [B,Cin,T,D] (+ layer selection) ->
    fp16 [C,B,T,P] (normalised * mask), fp32 (raw) or the weights' 16-bit type (raw, feeds temporal).  w1 [C,H,D],
    w2 [C,P,H] fp16 or bf16.  The hidden activation never reaches HBM (kws_mlp_fused)."""
    lib = _lib.load()
    B, Cin, T, D = x.shape
    Cc = len(layer_idx)
    dt = w1.dtype
    if dt not in (torch.float16, torch.bfloat16) or w2.dtype != dt:
        raise KWSError(f"w1/w2 must share one 16-bit dtype, got {w1.dtype}/{w2.dtype}")
    H, P = w1.shape[1], w2.shape[1]
    if tuple(w1.shape) != (Cc, H, D) or tuple(w2.shape) != (Cc, P, H):
        raise KWSError("projector weight shapes do not match x / layer_idx")
    if mask is not None and tuple(mask.shape) != (B, Cc, T):
        raise KWSError(f"mask must be [B,C,T]=({B},{Cc},{T}), got {tuple(mask.shape)}")
    out_dt = {MLP_OUT_NORM_F16: torch.float16, MLP_OUT_RAW_F32: torch.float32, MLP_OUT_RAW_16: dt}[out_mode]
    out = torch.empty((Cc, B, T, P), dtype=out_dt, device=x.device)
    check(lib.kws_mlp_fused(_cuda(x, "x", torch.float32), B, Cin, T, D, _layers(layer_idx), Cc, H, P,
                            F16 if dt == torch.float16 else BF16, _cuda(w1, "w1", dt), _cuda(b1, "b1", torch.float32),
                            _cuda(w2, "w2", dt), _cuda(b2, "b2", torch.float32), _cuda(mask, "mask", torch.float32),
                            eps, out_mode, _cuda(out, "out"), _stream()), "kws_mlp_fused")
    return out


@_guard
def temporal(proj16: torch.Tensor, w16: torch.Tensor, bf: torch.Tensor, mask: Optional[torch.Tensor],
             eps: float = SIM_EPS) -> torch.Tensor:
    """proj fp16|bf16 [C,B,T,P] (mlp(..., MLP_OUT_RAW_16)) -> fp16 [C,B,ceil(T/2),P];
    w16/bf from fold_temporal_weights (same 16-bit type); mask fp32 [B,C,ceil(T/2)] or None."""
    lib = _lib.load()
    Cc, B, T, P = proj16.shape
    T2 = (T + 1) // 2
    dt = proj16.dtype
    if dt not in (torch.float16, torch.bfloat16) or w16.dtype != dt:
        raise KWSError(f"proj/w16 must share one 16-bit dtype, got {proj16.dtype}/{w16.dtype}")
    if tuple(w16.shape) != (Cc, 3, P // 8, P, 8):
        raise KWSError(f"w16 must be [C,3,P/8,P,8]=({Cc},3,{P // 8},{P},8), got {tuple(w16.shape)}")
    if mask is not None and tuple(mask.shape) != (B, Cc, T2):
        raise KWSError(f"LEF mask must be at pooled resolution [B,C,ceil(T/2)]=({B},{Cc},{T2}), "
                       f"got {tuple(mask.shape)}")
    out = torch.empty((Cc, B, T2, P), dtype=torch.float16, device=proj16.device)
    check(lib.kws_temporal(_cuda(proj16, "proj", dt), Cc, B, T, P, F16 if dt == torch.float16 else BF16,
                           _cuda(w16, "w16", dt), _cuda(bf, "bf", torch.float32),
                           _cuda(mask, "mask", torch.float32), eps, _cuda(out, "out"), _stream()), "kws_temporal")
    return out


# ---- per pair --------------------------------------------------------------------
def pitch_for(Tu: int) -> int:
    return (Tu + 7) // 8 * 8


@_guard
def sim(kwd_n: torch.Tensor, utt_n: torch.Tensor, want_f32: bool, want_f16: bool, diag: bool = False,
        out_f32: Optional[torch.Tensor] = None, out_f16: Optional[torch.Tensor] = None):
    """kwd_n fp16 [C,K,Tk,Dk], utt_n fp16 [C,U,Tu,Dk] ->
    (features fp32 [K,U,C,Tk,Tu] | None, features fp16 [K,U,C,Tk,pitch] | None); DIAG drops the U axis."""
    lib = _lib.load()
    Cc, K, Tk, Dk = kwd_n.shape
    Cu, U, Tu, Dku = utt_n.shape
    if Cc != Cu or Dk != Dku:
        raise KWSError(f"operand mismatch: kwd {tuple(kwd_n.shape)} vs utt {tuple(utt_n.shape)}")
    lead = (K,) if diag else (K, U)
    pitch = pitch_for(Tu)
    f32 = f16 = None
    if want_f32:
        f32 = out_f32 if out_f32 is not None else torch.empty(lead + (Cc, Tk, Tu), dtype=torch.float32,
                                                              device=kwd_n.device)
    if want_f16:
        # pitch padding is never read by the stem (it bounds-checks against Tu)
        f16 = out_f16 if out_f16 is not None else torch.empty(lead + (Cc, Tk, pitch), dtype=torch.float16,
                                                              device=kwd_n.device)
    check(lib.kws_sim(_cuda(kwd_n, "kwd_n", torch.float16), _cuda(utt_n, "utt_n", torch.float16), Cc, K, U, Tk, Tu,
                      Dk, PAIRS_DIAG if diag else PAIRS_ALL, _cuda(f32, "feat_f32", torch.float32),
                      _cuda(f16, "feat_f16", torch.float16), pitch, _stream()), "kws_sim")
    return f32, f16


@_guard
def stem(feat_f16: torch.Tensor, Tu: int, w_packed: torch.Tensor, bias: torch.Tensor, out_mode: int,
         out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """feat_f16 fp16 [..., C, Tk, pitch] -> NCHW fp32 [N,64,Ho,Wo] or channels_last bf16 (logical
    shape [N,64,Ho,Wo], physical [N,Ho,Wo,64])."""
    lib = _lib.load()
    Cc, Tk, pitch = feat_f16.shape[-3:]
    pairs = feat_f16.numel() // (Cc * Tk * pitch)
    Ho, Wo = (Tk + 1) // 2, (Tu + 1) // 2
    if out is None:
        if out_mode == STEM_OUT_NCHW_F32:
            out = torch.empty((pairs, 64, Ho, Wo), dtype=torch.float32, device=feat_f16.device)
        else:
            out = torch.empty((pairs, Ho, Wo, 64), dtype=torch.bfloat16, device=feat_f16.device)
    ws_bytes = lib.kws_stem_workspace_bytes(pairs, Cc, Tk, Tu)
    if ws_bytes and (workspace is None or workspace.numel() * workspace.element_size() < ws_bytes):
        workspace = torch.empty(ws_bytes // 4, dtype=torch.float32, device=feat_f16.device)
    check(lib.kws_stem(_cuda(feat_f16, "feat_f16", torch.float16), pairs, Cc, Tk, Tu, pitch,
                       _cuda(w_packed, "w_packed", torch.float16), _cuda(bias, "bias", torch.float32), out_mode,
                       _cuda(out, "out"), _cuda(workspace, "workspace") if ws_bytes else 0, _stream()), "kws_stem",
          launches=(Cc + 15) // 16)
    if out_mode == STEM_OUT_NHWC_BF16:
        return out.permute(0, 3, 1, 2)  # channels_last view
    return out


def sim_stem_supported(Cc: int, Tk: int, Tu: int, Dk: int, out_mode: int = STEM_OUT_NHWC_BF16) -> bool:
    """kws_sim_stem_supported: 1 = both output modes, 2 = bf16 channels-last only (C > 12, multi-pass), 0 = no."""
    r = _lib.load().kws_sim_stem_supported(Cc, Tk, Tu, Dk)
    return r == 1 or (r == 2 and out_mode in (STEM_OUT_NHWC_BF16, STEM_OUT_POOL_NHWC_BF16))


@_guard
def sim_stem(kwd_n: torch.Tensor, utt_n: torch.Tensor, w_fused: torch.Tensor, bias: torch.Tensor, out_mode: int,
             diag: bool = False, out: Optional[torch.Tensor] = None, k_range: Optional[Tuple[int, int]] = None,
             u_range: Optional[Tuple[int, int]] = None, per_keyword: bool = False,
             kwd_len: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused similarity + stem (w_fused from pack_stem_fused).  kwd_n fp16 [C,K,Tk,Dk], utt_n fp16 [C,U,Tu,Dk] -> stem activation of the
    pairs of keywords k_range=(k0,k1) x utterances u_range=(u0,u1) (default: all), pair = (k-k0)*(u1-u0) + (u-u0)
    (DIAG: pair = k-k0): NCHW fp32 [N,64,Ho,Wo] or channels_last bf16.  ``out`` may be a larger reused
    buffer (its first N pairs are written).  ``per_keyword``: utt_n is [C, K*U, Tu, Dk], one utterance-side operand
    per (keyword, utterance) (config #4: the native-resolution similarity, contracted with the resize's height map).
    ``kwd_len``: int32 [K] valid frames per keyword (frames beyond are zero rows of kwd_n): output rows beyond a
    keyword are filled with relu(bias) without similarity / stem work -- bit-identical output (kws_sim_stem_ragged)."""
    lib = _lib.load()
    Cc, K, Tk, Dk = kwd_n.shape
    Cu, U, Tu, Dku = utt_n.shape
    if Cc != Cu or Dk != Dku:
        raise KWSError(f"operand mismatch: kwd {tuple(kwd_n.shape)} vs utt {tuple(utt_n.shape)}")
    if per_keyword:
        if diag or U % K != 0:
            raise KWSError(f"per_keyword needs utt_n [C, K*U, Tu, Dk] (got {U} items for {K} keywords) and diag=False")
        U //= K
    k0, k1 = k_range if k_range is not None else (0, K)
    u0, u1 = u_range if u_range is not None else (0, U)
    pairs = (k1 - k0) if diag else (k1 - k0) * (u1 - u0)
    Ho, Wo = (Tk + 1) // 2, (Tu + 1) // 2
    f32 = out_mode == STEM_OUT_NCHW_F32
    if out is None:
        out = torch.empty((pairs, 64, Ho, Wo) if f32 else (pairs, Ho, Wo, 64),
                          dtype=torch.float32 if f32 else torch.bfloat16, device=kwd_n.device)
    else:
        want = torch.float32 if f32 else torch.bfloat16
        if out.dtype != want or out.numel() < pairs * 64 * Ho * Wo:
            raise KWSError(f"out buffer too small / wrong dtype for {pairs} pairs")
    if kwd_len is not None and (kwd_len.dtype != torch.int32 or kwd_len.numel() != K):
        raise KWSError(f"kwd_len must be int32 [K={K}]")
    check(lib.kws_sim_stem_ragged(_cuda(kwd_n, "kwd_n", torch.float16), _cuda(utt_n, "utt_n", torch.float16),
                                 _cuda(kwd_len, "kwd_len", torch.int32), Cc, K,
                                 U, Tk, Tu, Dk,
                                 PAIRS_PER_KEYWORD if per_keyword else (PAIRS_DIAG if diag else PAIRS_ALL), k0, k1 - k0,
                                 u0, u1 - u0,
                                 _cuda(w_fused, "w_fused", torch.float16), _cuda(bias, "bias", torch.float32),
                                 out_mode, _cuda(out, "out"), _stream()), "kws_sim_stem", launches=(Cc + 11) // 12)
    shape = (pairs, 64, Ho, Wo) if f32 else (pairs, Ho, Wo, 64)
    view = out.view(-1)[: pairs * 64 * Ho * Wo].view(shape)
    if out_mode == STEM_OUT_NHWC_BF16:
        return view.permute(0, 3, 1, 2)
    return view


def sim_stem_pool_workspace_bytes(Cc: int, pairs: int, Tk: int, Tu: int) -> int:
    """Bytes of partial-sum workspace sim_stem_pool needs for ``pairs`` pairs per call (0 up to 12 layers)."""
    return int(_lib.load().kws_sim_stem_pool_workspace_bytes(int(Cc), int(pairs), int(Tk), int(Tu)))


@_guard
def sim_stem_pool(kwd_n: torch.Tensor, utt_n: torch.Tensor, w_fused: torch.Tensor, bias: torch.Tensor,
                  diag: bool = False, out: Optional[torch.Tensor] = None, k_range: Optional[Tuple[int, int]] = None,
                  u_range: Optional[Tuple[int, int]] = None, kwd_len: Optional[torch.Tensor] = None,
                  workspace: Optional[torch.Tensor] = None, per_keyword: bool = False) -> torch.Tensor:
    """Fused similarity + stem + MaxPool2d(3,2,1) (kws_sim_stem_pool): the activation ResNetEmbeddings hands to the
    encoder, bf16 channels_last [N,64,ceil(Ho/2),ceil(Wo/2)]; arguments as sim_stem.  More than 12 layers need a
    ``workspace`` for the partial sums of the channel-group passes (allocated here when not given).  ``per_keyword``:
    utt_n is [C, K*U, Tu, Dk], one utterance-side operand per (keyword, utterance) (config #4)."""
    lib = _lib.load()
    Cc, K, Tk, Dk = kwd_n.shape
    Cu, U, Tu, Dku = utt_n.shape
    if Cc != Cu or Dk != Dku:
        raise KWSError(f"operand mismatch: kwd {tuple(kwd_n.shape)} vs utt {tuple(utt_n.shape)}")
    if per_keyword:
        if diag or kwd_len is not None or U % K != 0:
            raise KWSError("per_keyword needs utt_n [C, K*U, Tu, Dk], diag=False and no keyword length table")
        U //= K
    k0, k1 = k_range if k_range is not None else (0, K)
    u0, u1 = u_range if u_range is not None else (0, U)
    pairs = (k1 - k0) if diag else (k1 - k0) * (u1 - u0)
    Ho, Wo = (Tk + 1) // 2, (Tu + 1) // 2
    Hp, Wp = (Ho + 1) // 2, (Wo + 1) // 2
    if out is None:
        out = torch.empty((pairs, Hp, Wp, 64), dtype=torch.bfloat16, device=kwd_n.device)
    elif out.dtype != torch.bfloat16 or out.numel() < pairs * 64 * Hp * Wp:
        raise KWSError(f"out buffer too small / wrong dtype for {pairs} pooled pairs")
    if kwd_len is not None and (kwd_len.dtype != torch.int32 or kwd_len.numel() != K):
        raise KWSError(f"kwd_len must be int32 [K={K}]")
    need = lib.kws_sim_stem_pool_workspace_bytes(Cc, pairs, Tk, Tu)
    if need and (workspace is None or workspace.numel() * workspace.element_size() < need):
        workspace = torch.empty(need, dtype=torch.uint8, device=kwd_n.device)
    check(lib.kws_sim_stem_pool(_cuda(kwd_n, "kwd_n", torch.float16), _cuda(utt_n, "utt_n", torch.float16),
                                _cuda(kwd_len, "kwd_len", torch.int32), Cc, K, U, Tk, Tu, Dk,
                                PAIRS_PER_KEYWORD if per_keyword else (PAIRS_DIAG if diag else PAIRS_ALL), k0, k1 - k0,
                                u0, u1 - u0,
                                _cuda(w_fused, "w_fused", torch.float16), _cuda(bias, "bias", torch.float32),
                                _cuda(out, "out"), _cuda(workspace, "workspace") if need else None, _stream()),
          "kws_sim_stem_pool", launches=(Cc + 11) // 12)
    return out.view(-1)[: pairs * 64 * Hp * Wp].view(pairs, Hp, Wp, 64).permute(0, 3, 1, 2)


@_guard
def interp_rows(x: torch.Tensor, layer_idx: Sequence[int], T_out: int, eps: float = SIM_EPS) -> torch.Tensor:
    """x fp32 [B,Cin,T,D] -> fp16 [C,B,T_out,D]: bilinear (align_corners=False) resampling of the L2-normalised
    frames along T (config #4: the width map of the image resize applied to the utterance operand)."""
    lib = _lib.load()
    B, Cin, T, D = x.shape
    Cc = len(layer_idx)
    out = torch.empty((Cc, B, T_out, D), dtype=torch.float16, device=x.device)
    check(lib.kws_interp_rows(_cuda(x, "x", torch.float32), B, Cin, T, D, _layers(layer_idx), Cc, int(T_out), eps,
                              _cuda(out, "out"), _stream()), "kws_interp_rows")
    return out


@_guard
def sim_operand(kwd_n: torch.Tensor, utt_n: torch.Tensor) -> torch.Tensor:
    """kwd_n fp16 [C,K,Tk,Dk] (Tk % 16 == 0), utt_n fp16 [C,U,Tu,Dk] -> similarity as a K-major fp16 operand
    [C, K*U, Tu, Tk] (item = k*U + u)."""
    lib = _lib.load()
    Cc, K, Tk, Dk = kwd_n.shape
    Cu, U, Tu, Dku = utt_n.shape
    if Cc != Cu or Dk != Dku:
        raise KWSError(f"operand mismatch: kwd {tuple(kwd_n.shape)} vs utt {tuple(utt_n.shape)}")
    out = torch.empty((Cc, K * U, Tu, Tk), dtype=torch.float16, device=kwd_n.device)
    check(lib.kws_sim_operand(_cuda(kwd_n, "kwd_n", torch.float16), _cuda(utt_n, "utt_n", torch.float16), Cc, K, U, Tk,
                              Tu, Dk, _cuda(out, "out"), _stream()), "kws_sim_operand")
    return out


@_guard
def resize_row_weights(src_h: Optional[torch.Tensor], K: int, Cc: int, Hp: int, Ho: int, device=None) -> torch.Tensor:
    """Height map of the bilinear resize as an operand: fp16 [C,K,Ho,Hp] (src_h int32 [K] valid frames, or None)."""
    lib = _lib.load()
    if src_h is not None and (src_h.dtype != torch.int32 or src_h.numel() != K):
        raise KWSError("src_h must be int32 [K]")
    dev = src_h.device if src_h is not None else device
    out = torch.empty((Cc, K, Ho, Hp), dtype=torch.float16, device=dev)
    check(lib.kws_resize_row_weights(_cuda(src_h, "src_h", torch.int32), K, Cc, Hp, Ho, _cuda(out, "out"), _stream()),
          "kws_resize_row_weights")
    return out


@_guard
def resize_bilinear(feat_f32: torch.Tensor, src_h: Optional[torch.Tensor], size: Tuple[int, int],
                    want_f32: bool = False, want_f16: bool = True):
    """feat fp32 [K,U,C,Hs,Ws] (+ src_h int32 [K]: valid rows per keyword) -> bilinear (align_corners=False)
    resize to ``size``: (fp32 [K,U,C,Ho,Wo] | None, fp16 [K,U,C,Ho,pitch] | None)."""
    lib = _lib.load()
    K, U, Cc, Hs, Ws = feat_f32.shape
    Ho, Wo = size
    pitch = pitch_for(Wo)
    o32 = torch.empty((K, U, Cc, Ho, Wo), dtype=torch.float32, device=feat_f32.device) if want_f32 else None
    o16 = torch.empty((K, U, Cc, Ho, pitch), dtype=torch.float16, device=feat_f32.device) if want_f16 else None
    if src_h is not None and (src_h.dtype != torch.int32 or src_h.numel() != K):
        raise KWSError("src_h must be int32 [K]")
    check(lib.kws_resize_bilinear(_cuda(feat_f32, "feat_f32", torch.float32), _cuda(src_h, "src_h", torch.int32), K, U,
                                  Cc, Hs, Ws, Ho, Wo, _cuda(o32, "out_f32"), _cuda(o16, "out_f16"), pitch, _stream()),
          "kws_resize_bilinear", launches=(K + max(1, 65535 // (U * Cc)) - 1) // max(1, 65535 // (U * Cc)))
    return o32, o16


@_guard
def maxpool_nhwc(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """MaxPool2d(3, 2, 1) of a channels-last bf16 activation (logical [N,C,H,W], physical [N,H,W,C]) ->
    channels-last bf16 [N,C,ceil(H/2),ceil(W/2)]; bit-identical to F.max_pool2d."""
    lib = _lib.load()
    if x.dim() != 4:
        raise KWSError(f"maxpool_nhwc expects [N,C,H,W], got {tuple(x.shape)}")
    N, Cc, H, W = x.shape
    phys = x.permute(0, 2, 3, 1)
    if not phys.is_contiguous():
        raise KWSError("maxpool_nhwc needs a channels_last (NHWC-contiguous) tensor")
    Hp, Wp = (H + 1) // 2, (W + 1) // 2
    if out is None:
        out = torch.empty((N, Hp, Wp, Cc), dtype=torch.bfloat16, device=x.device)
    elif out.dtype != torch.bfloat16 or out.numel() < N * Hp * Wp * Cc:
        raise KWSError("maxpool_nhwc: out buffer too small / wrong dtype")
    check(lib.kws_maxpool_nhwc(_cuda(phys, "x", torch.bfloat16), N, H, W, Cc, _cuda(out, "out"), _stream()),
          "kws_maxpool_nhwc")
    return out.view(-1)[: N * Hp * Wp * Cc].view(N, Hp, Wp, Cc).permute(0, 3, 1, 2)


# ---- scores ------------------------------------------------------------------------
@_guard
def scores(logits: torch.Tensor, hotword_mask: Optional[torch.Tensor], threshold: float):
    """logits fp32 [n,2] -> (scores fp32 [n], detections uint8 [n])"""
    lib = _lib.load()
    n = logits.shape[0]
    logits = logits.float().contiguous()
    sc = torch.empty(n, dtype=torch.float32, device=logits.device)
    det = torch.empty(n, dtype=torch.uint8, device=logits.device)
    if hotword_mask is not None:
        hotword_mask = hotword_mask.float().contiguous()
    check(lib.kws_scores(_cuda(logits, "logits", torch.float32), _cuda(hotword_mask, "hotword_mask", torch.float32),
                         n, float(threshold), _cuda(sc, "scores"), _cuda(det, "det"), _stream()), "kws_scores")
    return sc, det


@_guard
def topk(scores_cu: torch.Tensor, k: int, ids: Optional[torch.Tensor] = None, id_offset: int = 0):
    """scores fp32 [n_cand,U] (+ ids int32 [n_cand,U]) -> (top scores [k,U], top ids int32 [k,U]);
    ties broken by the lower id."""
    lib = _lib.load()
    n, U = scores_cu.shape
    os_ = torch.empty((k, U), dtype=torch.float32, device=scores_cu.device)
    oi = torch.empty((k, U), dtype=torch.int32, device=scores_cu.device)
    ws_bytes = lib.kws_topk_workspace_bytes(n, U, int(k))
    ws = torch.empty(ws_bytes // 8, dtype=torch.int64, device=scores_cu.device) if ws_bytes else None
    levels, m = 1, n
    while m > 2048:  # one launch per selection level (segments of 2048 candidates keep k survivors each)
        m = (m + 2047) // 2048 * int(k)
        levels += 1
    check(lib.kws_topk(_cuda(scores_cu, "scores", torch.float32), _cuda(ids, "ids", torch.int32), n, U,
                       int(id_offset), int(k), _cuda(os_, "out_scores"), _cuda(oi, "out_ids"), _cuda(ws, "workspace"),
                       _stream()), "kws_topk", launches=levels)
    return os_, oi
