"""Keyword-bank loader: the reference's on-disk ``*.bin`` hidden states -> the
resident compressed fp16 operand bank of the B200 path (SURVEY.md section 8f
rank 2).

On-disk format (src/utils.py:186-201): one ``torch.save``d fp32 tensor
``[12, T, D]`` per keyword / utterance -- Whisper encoder ``hidden_states[10:22]``
trimmed to ``T = ceil(mel_frames / 2)`` frames and L2-normalised over ``D`` --
named ``str(idx).zfill(n) + ".bin"`` (src/efficient_kws/dataset.py:696-738);
missing indices are *ghost* keywords, scored as all-zero tensors whose score is
multiplied by a 0 ``hotword_mask`` afterwards (dataset.py:730-738, model.py:783-789).

The reference pads every keyword to ``features_size[0]`` frames and re-sends the
whole padded bank with every DataLoader item.  Here the ragged tensors are padded
on the fly, a few hundred at a time, pushed through the model's compression
kernels and only the compressed operands ``[C, K, Tk', Dk]`` stay resident
(cfg5: 100 k keywords x 32 layers = 1.2 TB padded fp32, 30.7 GB compressed).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

from . import ops


def read_bin(path: str) -> torch.Tensor:
    """One ``*.bin`` file -> CPU fp32 tensor [layers, T, D] (written by src/utils.py:199-201)."""
    with open(path, "rb") as f:
        t = torch.load(f, map_location="cpu", weights_only=True)
    if not isinstance(t, torch.Tensor) or t.dim() != 3:
        raise ops.KWSError(f"{path}: expected a [layers, T, D] tensor, got {type(t).__name__} "
                           f"{tuple(t.shape) if isinstance(t, torch.Tensor) else ''}")
    return t.detach().float()


def iter_bin_dir(folder: str, n_items: Optional[int] = None) -> Iterator[Optional[torch.Tensor]]:
    """Keyword hidden states of a ``keywords-hs/<kw_type>`` folder in index order; ``None`` for ghost
    (missing) indices (dataset.py:700-729).  ``n_items`` = number of keywords (len(keywords.txt)); default:
    the highest index present + 1."""
    present = {}
    for name in os.listdir(folder):
        stem, ext = os.path.splitext(name)
        if ext == ".bin" and stem.isdigit():
            present[int(stem)] = name
    if n_items is None:
        n_items = max(present) + 1 if present else 0
    for idx in range(n_items):
        name = present.get(idx)
        yield read_bin(os.path.join(folder, name)) if name is not None else None


def pad_item(hs: torch.Tensor, n_frames: int, n_layers: int) -> Tuple[torch.Tensor, int]:
    """Layer selection ``x[-n_layers:]`` (dataset.py:815-816), truncation / zero padding of the frame axis to
    ``n_frames`` (dataset.py:784-814).  -> ([n_layers, n_frames, D], valid length)."""
    hs = hs[-n_layers:]
    T = min(hs.shape[1], n_frames)
    out = hs.new_zeros((hs.shape[0], n_frames, hs.shape[2]))
    out[:, :T] = hs[:, :T]
    return out, T


@dataclass
class KeywordBank:
    """Resident compressed keyword operands of one model checkpoint."""

    kwd_n: torch.Tensor  # fp16 [C, K, Tk', Dk]: L2-normalised, frame mask folded
    lengths: torch.Tensor  # int32 [K] valid frames before compression (0 for ghosts)
    hotword_mask: torch.Tensor  # fp32 [K]: 0 for ghost keywords, 1 otherwise
    lef: bool = False  # kwd_n is at pooled resolution ceil(T/2) (LEF variant)

    @property
    def K(self) -> int:
        return self.kwd_n.shape[1]

    def shard(self, lo: int, hi: int) -> "KeywordBank":
        return KeywordBank(self.kwd_n[:, lo:hi].contiguous(), self.lengths[lo:hi], self.hotword_mask[lo:hi], self.lef)

    def compressed_lengths(self) -> torch.Tensor:
        """int32 [K]: valid frames at the resolution of ``kwd_n`` (LEF halves the frame axis: mask[..., ::2] keeps
        ceil(len / 2) frames) -- the length table of kws_sim_stem_ragged."""
        Tc = self.kwd_n.shape[2]
        lens = self.lengths.to(torch.int32)
        if self.lef:
            lens = (lens + 1) // 2
        return torch.clamp(lens, max=Tc).to(torch.int32).contiguous()


@torch.no_grad()
def build_keyword_bank(model, items: Iterable[Optional[torch.Tensor]], n_frames: int, device=None,
                       chunk: int = 256) -> KeywordBank:
    """Stream ragged keyword hidden states ([layers, T_i, D] CPU tensors, ``None`` = ghost) through the model's
    compression kernels, ``chunk`` keywords at a time.  ``model`` is a ``KWSModelB200`` (its variant decides
    L / LE / LEF; LEF masks are taken at pooled resolution ``mask[..., ::2]``)."""
    device = torch.device(device) if device is not None else next(model.parameters()).device
    eng = model.prepare(device)
    C = model.hparams.n_layers
    lef = model.variant == "LEF"
    outs: List[torch.Tensor] = []
    lens: List[int] = []
    hot: List[float] = []
    buf: List[torch.Tensor] = []
    blens: List[int] = []
    D = None

    def flush():
        if not buf:
            return
        x = torch.stack(buf).to(device, non_blocking=True)  # [b, C, n_frames, D]
        ln = torch.tensor(blens, device=device)
        m = (torch.arange(n_frames, device=device)[None] < ln[:, None]).float()  # [b, n_frames]
        if lef:
            m = m[:, ::2]
        m = m[:, None, :].expand(-1, C, -1).contiguous()
        outs.append(eng.compress(x, m, list(range(C))))
        buf.clear()
        blens.clear()

    for hs in items:
        if hs is None:
            if D is None:
                # shape is needed for the all-zero stand-in; defer until a real item was seen
                buf.append(None)
                blens.append(0)
            else:
                buf.append(torch.zeros((C, n_frames, D)))
                blens.append(0)
            lens.append(0)
            hot.append(0.0)
        else:
            if hs.shape[0] < C:
                raise ops.KWSError(f"item has {hs.shape[0]} layers, model selects the last {C}")
            if D is None:
                D = hs.shape[2]
                for i, b in enumerate(buf):
                    if b is None:
                        buf[i] = torch.zeros((C, n_frames, D))
            padded, T = pad_item(hs, n_frames, C)
            buf.append(padded)
            blens.append(T)
            lens.append(T)
            hot.append(1.0)
        if len(buf) >= chunk and D is not None:
            flush()
    if D is None:
        raise ops.KWSError("keyword bank has no real (non-ghost) item")
    flush()
    kwd_n = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
    return KeywordBank(kwd_n.contiguous(), torch.tensor(lens, dtype=torch.int32, device=device),
                       torch.tensor(hot, dtype=torch.float32, device=device), lef)


@torch.no_grad()
def score_bank(model, bank: KeywordBank, utt_features: torch.Tensor, utt_mask: torch.Tensor, max_pairs: int = 256,
               threshold: Optional[float] = None):
    """All keywords of the bank x the given utterances ([U, layers, Tu, D] fp32 + mask [U, C, Tu']) ->
    (scores [K, U], detections uint8 [K, U], logits [K, U, 2]); ghosts score 0 (model.py:783-789)."""
    eng = model.prepare(bank.kwd_n.device)
    C = model.hparams.n_layers
    utt = utt_features[:, -C:].contiguous() if utt_features.shape[1] != C else utt_features
    utt_n = eng.compress(utt.to(bank.kwd_n.device), utt_mask.to(bank.kwd_n.device), list(range(C)))
    return model.score_compressed(bank.kwd_n, utt_n, bank.hotword_mask, max_pairs, threshold,
                                  kwd_len=bank.compressed_lengths())
