"""Inference form of the (third-party) ResNet body that follows the stem.

The reference runs HuggingFace's ``ResNetModel`` unfused (src/efficient_kws/resnet.py:42-58 ->
modeling_resnet.py: max-pool, residual stages, adaptive avg-pool) and a ``Linear`` head.  Everything in it is
library code (cuDNN / cuBLAS) and stays library code here; this module only removes the elementwise passes
between the library calls: every BatchNorm is folded into the convolution before it (running statistics,
eps of the module) and each convolution is issued as ONE cuDNN fused op,

    conv + bias + ReLU                 torch.cudnn_convolution_relu
    conv + residual + bias + ReLU      torch.cudnn_convolution_add_relu   (last conv of a residual layer)
    conv                               F.conv2d                            (projection shortcut; its folded bias is
                                                                            added to the bias of the residual
                                                                            layer's last conv, so no bias pass)

on channels-last 16-bit activations.  Not part of the hot path of SURVEY.md section 8; it bounds the end-to-end
number, which is why it is worth running well.  CUDA only.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = conv.weight.detach().float() * s.view(-1, 1, 1, 1)
    b = bn.bias.detach().float() - bn.running_mean.detach().float() * s
    if conv.bias is not None:
        b = b + conv.bias.detach().float() * s
    return w.to(dtype).contiguous(memory_format=torch.channels_last), b.to(dtype).contiguous()


class _Conv:
    __slots__ = ("w", "b", "stride", "padding")

    def __init__(self, layer, dtype):  # ResNetConvLayer | ResNetShortCut
        conv = layer.convolution
        self.w, self.b = _fold(conv, layer.normalization, dtype)
        self.stride, self.padding = tuple(conv.stride), tuple(conv.padding)

    def relu(self, x):
        return torch.cudnn_convolution_relu(x, self.w, self.b, self.stride, self.padding, (1, 1), 1)

    def add_relu(self, x, z):
        return torch.cudnn_convolution_add_relu(x, self.w, z, 1.0, self.b, self.stride, self.padding, (1, 1), 1)

    def plain(self, x):  # bias-free: the caller has moved self.b into the consumer's bias
        return F.conv2d(x, self.w, None, self.stride, self.padding)


class FusedBody:
    """``FusedBody(resnet, dtype)(stem_activation) -> logits fp32 [N, num_labels]``.

    ``resnet``: the wrapper of src/efficient_kws/resnet.py (``feature_extractor`` = HF ResNetModel,
    ``classifier`` = Flatten + Linear), bottleneck or basic layers.  ``stem_activation``: [N,64,Ho,Wo] in
    ``dtype``, channels-last (what kws_sim_stem writes).  ``pooled=True``: the input is already max-pooled."""

    def __init__(self, resnet: nn.Module, dtype: torch.dtype = torch.bfloat16):
        fe = resnet.feature_extractor
        self.dtype = dtype
        self.pool = fe.embedder.pooler
        self.stages: List[List[Tuple[Optional[_Conv], List[_Conv]]]] = []
        for stage in fe.encoder.stages:
            layers = []
            for layer in stage.layers:
                sc = None if isinstance(layer.shortcut, nn.Identity) else _Conv(layer.shortcut, dtype)
                convs = [_Conv(cl, dtype) for cl in layer.layer]
                if sc is not None:
                    # relu(conv3(h) + b3 + (shortcut(x) + bs)) == relu(conv3(h) + (b3 + bs) + shortcut_nobias(x)):
                    # torch's conv2d-with-bias is a convolution plus an elementwise pass over the widest tensor of the block
                    _, b_sc = _fold(layer.shortcut.convolution, layer.shortcut.normalization, torch.float32)
                    _, b_last = _fold(layer.layer[-1].convolution, layer.layer[-1].normalization, torch.float32)
                    convs[-1].b = (b_last + b_sc).to(dtype).contiguous()
                layers.append((sc, convs))
            self.stages.append(layers)
        lin = resnet.classifier[1]
        self.lin_w = lin.weight.detach().float()
        self.lin_b = lin.bias.detach().float()

    def _stage(self, x, layers):
        for sc, convs in layers:
            res = x if sc is None else sc.plain(x)
            h = x
            for c in convs[:-1]:
                h = c.relu(h)
            x = convs[-1].add_relu(h, res)
        return x

    @torch.no_grad()
    def __call__(self, stem_activation: torch.Tensor, pooled: bool = False) -> torch.Tensor:
        x = stem_activation
        if x.device.type != "cuda":
            raise RuntimeError("FusedBody runs on CUDA only (cuDNN fused convolutions)")
        if x.dtype != self.dtype:
            x = x.to(self.dtype)
        x = x.contiguous(memory_format=torch.channels_last)
        if not pooled:
            x = self.pool(x)
        for layers in self.stages:
            x = self._stage(x, layers)
        feat = x.float().mean(dim=(2, 3))  # AdaptiveAvgPool2d((1,1)) + flatten
        return F.linear(feat, self.lin_w, self.lin_b)

    # development aid: time of one part (0 = max-pool, 1..4 = stages) on a given input batch
    @torch.no_grad()
    def stage_times(self, stem_activation: torch.Tensor, part: int, iters: int = 3) -> float:
        x = self.pool(stem_activation.contiguous(memory_format=torch.channels_last))
        if part == 0:
            fn, arg = self.pool, stem_activation
        else:
            for layers in self.stages[: part - 1]:
                x = self._stage(x, layers)
            fn, arg = (lambda t: self._stage(t, self.stages[part - 1])), x
        fn(arg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn(arg)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
