"""Small fixed workload that launches every hot-path kernel a few times (for ncu and first
timings; not a benchmark -- bench.py is).  Shapes are cfg2-like slabs (LE, C=12, D=768).

    python tools/prof_kernels.py [--pairs-k 16] [--utts 2] [--iters 3] [--only stem|sim|mlp|rows|temporal|maxpool|fused]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enhance_cb_whisper_b200 import ops  # noqa: E402


def timeit(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs-k", type=int, default=74)
    ap.add_argument("--utts", type=int, default=2)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--C", type=int, default=12)
    ap.add_argument("--D", type=int, default=768)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    Cc, K, U, Tk, Tu, D, P = a.C, a.pairs_k, a.utts, 150, 1500, a.D, 64
    H = D // 2
    pairs = K * U
    only = set(filter(None, a.only.split(",")))
    want = lambda s: not only or s in only

    def unit(*shape):
        x = torch.randn(*shape, generator=g, device=dev)
        return x / x.norm(dim=-1, keepdim=True)

    if want("rows") or want("mlp"):
        xk = unit(K, Cc, Tk, D)
        if want("rows"):
            t = timeit(lambda: ops.cast_rows16(xk, list(range(Cc)), ops.F16), a.iters)
            print(f"cast_rows16 {K}x{Cc}x{Tk}x{D}: {t:.3f} ms -> {xk.numel() * 6 / t / 1e6:.0f} GB/s")
            mk = torch.ones(K, Cc, Tk, device=dev)
            t = timeit(lambda: ops.normalize_rows(xk, list(range(Cc)), mk), a.iters)
            print(f"normalize_rows: {t:.3f} ms -> {xk.numel() * 6 / t / 1e6:.0f} GB/s")
        if want("mlp"):
            x16 = ops.cast_rows16(xk, list(range(Cc)), ops.F16)
            w1 = ops.cast16(torch.randn(Cc, H, D, generator=g, device=dev) / D ** 0.5)
            w2 = ops.cast16(torch.randn(Cc, P, H, generator=g, device=dev) / H ** 0.5)
            b1 = torch.zeros(Cc, H, device=dev)
            b2 = torch.zeros(Cc, P, device=dev)
            t = timeit(lambda: ops.mlp(x16, K, Tk, w1, b1, w2, b2, None, ops.MLP_OUT_NORM_F16), a.iters)
            fl = 2.0 * Cc * K * Tk * (D * H + H * P)
            print(f"mlp {Cc}x{K * Tk}x{D}: {t:.3f} ms -> {fl / t / 1e9:.1f} TFLOP/s")
    if want("temporal"):
        proj = torch.randn(Cc, K, Tk, P, generator=g, device=dev).half()
        cw = torch.randn(Cc, P, P, 3, generator=g, device=dev) / (3 * P) ** 0.5
        one, zero = torch.ones(Cc, P, device=dev), torch.zeros(Cc, P, device=dev)
        wf, bf = ops.fold_temporal_weights(cw, zero, one, zero, zero, one)
        t = timeit(lambda: ops.temporal(proj, wf, bf, None), a.iters)
        byt = proj.numel() * 2 + proj.numel() // 2 * 2
        print(f"temporal {Cc}x{K}x{Tk}x{P}: {t:.3f} ms -> {byt / t / 1e6:.0f} GB/s, "
              f"{2.0 * 3 * P * P * proj.numel() / P / t / 1e9:.1f} TFLOP/s")
    if want("maxpool"):
        act = torch.rand(pairs, 75, 750, 64, generator=g, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
        t = timeit(lambda: ops.maxpool_nhwc(act), a.iters)
        byt = act.numel() * 2 * 1.25
        print(f"maxpool {pairs}x75x750x64: {t:.3f} ms -> {byt / t / 1e6:.0f} GB/s")
        del act
    if only and only <= {"rows", "mlp", "temporal", "maxpool"}:
        return
    kn = unit(Cc, K, Tk, P).half()
    un = unit(Cc, U, Tu, P).half()
    pitch = ops.pitch_for(Tu)
    f16 = torch.empty(K, U, Cc, Tk, pitch, dtype=torch.float16, device=dev)
    if want("sim") or want("stem"):
        t = timeit(lambda: ops.sim(kn, un, False, True, out_f16=f16), a.iters)
        print(f"sim {pairs} pairs C={Cc}: {t:.3f} ms -> {pairs / t * 1e3:.0f} pairs/s, "
              f"{2.0 * Cc * Tk * Tu * P * pairs / t / 1e9:.1f} TFLOP/s, out {f16.numel() * 2 / t / 1e6:.0f} GB/s")
    if want("stem"):
        conv_w = torch.randn(64, Cc, 7, 7, generator=g, device=dev) * 0.05
        one, zero = torch.ones(64, device=dev), torch.zeros(64, device=dev)
        wp, bias = ops.pack_stem_weights(conv_w, one, zero, zero, one)
        out = torch.empty(pairs, 75, 750, 64, dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: ops.stem(f16, Tu, wp, bias, ops.STEM_OUT_NHWC_BF16, out=out), a.iters)
        fl = 2.0 * 64 * 49 * Cc * 75 * 750 * pairs
        print(f"stem {pairs} pairs C={Cc}: {t:.3f} ms -> {pairs / t * 1e3:.0f} pairs/s, {fl / t / 1e9:.1f} TFLOP/s (algorithmic)")
    if want("fused"):
        conv_w = torch.randn(64, Cc, 7, 7, generator=g, device=dev) * 0.05
        one, zero = torch.ones(64, device=dev), torch.zeros(64, device=dev)
        wp, bias = ops.pack_stem_fused(conv_w, one, zero, zero, one)
        out = torch.empty(pairs, 75, 750, 64, dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NHWC_BF16, out=out), a.iters)
        fl = (2.0 * 64 * 49 * Cc * 75 * 750 + 2.0 * Cc * Tk * Tu * P) * pairs
        print(f"fused sim+stem {pairs} pairs C={Cc}: {t:.3f} ms -> {pairs / t * 1e3:.0f} pairs/s, {fl / t / 1e9:.1f} TFLOP/s (algorithmic)")
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
