"""Probe: how fast can the (third-party, out-of-scope) ResNet body behind the stem run on this box?

The body (max-pool, 4 residual stages, avg-pool, Linear; HF modeling_resnet.py via src/efficient_kws/resnet.py:42-58)
bounds the end-to-end number.  Variants timed on one batch of bf16 channels-last stem activations:
  eager   : the unmodified HF modules in bf16 channels_last (what model._body runs today)
  bench   : same with torch.backends.cudnn.benchmark
  fused   : BatchNorms folded into the convolutions, every conv issued as cuDNN's fused conv+bias+ReLU /
            conv+residual+bias+ReLU (torch.cudnn_convolution_relu / cudnn_convolution_add_relu)
Prints ms per batch, pairs/s and the max |logit| difference against the fp32 modules.
"""
import argparse
import importlib
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("enhance_cb_whisper_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=250)
    ap.add_argument("--h", type=int, default=75)
    ap.add_argument("--w", type=int, default=750)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--dtype", default="bfloat16")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = getattr(torch, args.dtype)
    from enhance_cb_whisper_b200.model import Resnet, run_body
    from enhance_cb_whisper_b200.body import FusedBody

    torch.manual_seed(0)
    net = Resnet(12, 2).eval()
    g = torch.Generator().manual_seed(1)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.1 * torch.randn(m.bias.shape, generator=g)
            m.running_mean = 0.1 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var = 0.5 + torch.rand(m.running_var.shape, generator=g)
    net = net.to(dev)
    x = torch.rand((args.pairs, 64, args.h, args.w), device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)

    def timeit(fn, name):
        for _ in range(2):
            y = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            y = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        print(f"{name:10s} {ms:9.2f} ms/batch  {args.pairs / ms * 1e3:10.0f} pairs/s", flush=True)
        return y

    with torch.no_grad():
        small = x[:8]
        ref = run_body(net, small.float())
        import copy
        lowp = copy.deepcopy(net).to(dtype=dt, memory_format=torch.channels_last)
        lowp.classifier.float()
        y = timeit(lambda: run_body(lowp, x), "eager")
        print("  eager vs fp32 max|d|", (run_body(lowp, small) - ref).abs().max().item())
        fb = FusedBody(net, dt)
        y2 = timeit(lambda: fb(x), "fused")
        print("  fused vs fp32 max|d|", (fb(small) - ref).abs().max().item(), " vs eager", (y2 - y).abs().max().item())
        for st in range(5):
            t = fb.stage_times(x, st)
            print(f"  fused part {st}: {t:8.2f} ms")
        torch.backends.cudnn.benchmark = True
        timeit(lambda: run_body(lowp, x), "eager+bm")
        timeit(lambda: fb(x), "fused+bm")


if __name__ == "__main__":
    main()
