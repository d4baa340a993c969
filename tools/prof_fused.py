"""One launch set of a shipped fused-kernel instance (for `ncu --set full` and quick timings; not a benchmark).

    python tools/prof_fused.py cfg2|cfg1|cfg3|mlp2|mlp3 [--pairs-k 74] [--utts 2] [--iters 2]

cfg2: C=12, Dk=64, 150x1500  -> kws_fused_kernel<1,16,0,2,1>            (LE, the bench's headline instance)
cfg1: C=4,  Dk=384, 150x1500 -> kws_fused_kernel<1,48,0,2,0>            (L variant)
cfg3: C=32, Dk=64, 75x750    -> kws_fused_kernel<1,16,1,2,1> x2 + <1,16,1,2,0>  (LEF, multi-pass 12+12+8)
mlp2 / mlp3: the fused per-layer projector (kws_mlp_fused) at D=768 / D=1280 on [K,12|32,150,D] keyword rows.
148-pair launches: the bounded mbarrier waits trap under `--set full` on launches longer than ~10 ms.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enhance_cb_whisper_b200 import ops  # noqa: E402

SHAPES = {"cfg2": (12, 64, 150, 1500), "cfg1": (4, 384, 150, 1500), "cfg3": (32, 64, 75, 750)}


def _lib_ws(C, pairs, Tk, Tu):
    from enhance_cb_whisper_b200 import _lib

    return int(_lib.load().kws_sim_stem_pool_workspace_bytes(C, pairs, Tk, Tu))


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
    ap.add_argument("what", choices=sorted(SHAPES) + ["mlp2", "mlp3"])
    ap.add_argument("--pairs-k", type=int, default=74)
    ap.add_argument("--utts", type=int, default=2)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--ab-pairs", action="store_true", help="mlp*: A/B of the CTA-pair W1 multicast (debug-hooks library)")
    ap.add_argument("--pool-only", action="store_true", help="run only kws_sim_stem_pool (ncu capture of the POOL instance)")
    ap.add_argument("--pool", action="store_true", help="also time the fused stem+max-pool kernel against stem + kws_maxpool_nhwc")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1)
    if a.what.startswith("mlp"):
        Cc, D, P, T, K = (12, 768, 64, 150, a.pairs_k) if a.what == "mlp2" else (32, 1280, 64, 150, a.pairs_k)
        H = D // 2
        x = unit(K, Cc, T, D)
        w1 = ops.cast16(torch.randn(Cc, H, D, generator=g, device=dev) / D ** 0.5)
        w2 = ops.cast16(torch.randn(Cc, P, H, generator=g, device=dev) / H ** 0.5)
        b1 = torch.randn(Cc, H, generator=g, device=dev) * 0.1
        b2 = torch.randn(Cc, P, generator=g, device=dev) * 0.1
        mask = torch.ones(K, Cc, T, device=dev)
        fl = 2.0 * Cc * K * T * (D * H + H * P)
        by = x.numel() * 4 + Cc * K * T * P * 2
        modes = [None]
        if a.ab_pairs:  # same-run A/B of the CTA-pair W1 multicast (needs tools/experiments/mlp_cta_pairs_*.patch applied
            # and the debug flavour: KWS_B200_LIB=..._dbg.so)
            from enhance_cb_whisper_b200 import _lib
            modes = [0, 1, 0, 1]
        ref = None
        for md in modes:
            if md is not None:
                _lib.load().kws_debug_set_mlp_pairs(md)
            t = timeit(lambda: ops.mlp_fused(x, list(range(Cc)), w1, b1, w2, b2, mask, ops.MLP_OUT_NORM_F16), a.iters)
            out = ops.mlp_fused(x, list(range(Cc)), w1, b1, w2, b2, mask, ops.MLP_OUT_NORM_F16)
            same = "" if ref is None else f" bit-identical to the first mode: {bool(torch.equal(out, ref))}"
            ref = out if ref is None else ref
            print(f"mlp_fused {K}x{Cc}x{T}x{D}{'' if md is None else f' pairs={md}'}: {t:.3f} ms -> {fl / t / 1e9:.1f} TFLOP/s "
                  f"({fl / t / 1e9 / 1364.9:.3f} of sustained), {by / t / 1e6:.0f} GB/s ({by / t / 1e6 / 6554.2:.3f} of HBM peak){same}")
        return
    Cc, Dk, Tk, Tu = SHAPES[a.what]
    K, U = a.pairs_k, a.utts
    Ho, Wo = (Tk + 1) // 2, (Tu + 1) // 2
    kn, un = unit(Cc, K, Tk, Dk).half(), unit(Cc, U, Tu, Dk).half()
    one, zero = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    wp, bias = ops.pack_stem_fused(torch.randn(64, Cc, 7, 7, generator=g, device=dev) * 0.05, one, zero, zero, one)
    if a.pool_only:  # just the fused stem + max-pool kernel (for an ncu capture of the POOL instance)
        Hp, Wp = (Ho + 1) // 2, (Wo + 1) // 2
        pooled = torch.empty(K * U, Hp, Wp, 64, dtype=torch.bfloat16, device=dev)
        ws = torch.empty(max(1, _lib_ws(Cc, K * U, Tk, Tu)), dtype=torch.uint8, device=dev)
        tp = timeit(lambda: ops.sim_stem_pool(kn, un, wp, bias, out=pooled, workspace=ws), a.iters)
        print(f"fused+pool {a.what}, {K * U} pairs: {tp:.3f} ms -> {K * U / tp * 1e3:.0f} pairs/s")
        return
    out = torch.empty(K * U, Ho, Wo, 64, dtype=torch.bfloat16, device=dev)
    t = timeit(lambda: ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NHWC_BF16, out=out), a.iters)
    fl = (2.0 * 64 * 49 * Cc * Ho * Wo + 2.0 * Cc * Tk * Tu * Dk) * K * U
    print(f"fused {a.what} C={Cc} Dk={Dk} {Tk}x{Tu}, {K * U} pairs: {t:.3f} ms -> {K * U / t * 1e3:.0f} pairs/s, "
          f"{fl / t / 1e9:.1f} TFLOP/s algorithmic ({fl / t / 1e9 / 1364.9:.3f} of sustained)")
    if a.pool:  # A/B of SURVEY 8f row 3: stem + separate max-pool kernel vs the fused stem+pool kernel
        Hp, Wp = (Ho + 1) // 2, (Wo + 1) // 2
        pooled = torch.empty(K * U, Hp, Wp, 64, dtype=torch.bfloat16, device=dev)
        tm = timeit(lambda: ops.maxpool_nhwc(out.permute(0, 3, 1, 2), out=pooled), a.iters)
        ws = torch.empty(max(1, _lib_ws(Cc, K * U, Tk, Tu)), dtype=torch.uint8, device=dev)
        tp = timeit(lambda: ops.sim_stem_pool(kn, un, wp, bias, out=pooled, workspace=ws), a.iters)
        print(f"  stem {t:.3f} ms + kws_maxpool_nhwc {tm:.3f} ms = {t + tm:.3f} ms ({K * U / (t + tm) * 1e3:.0f} pairs/s to the "
              f"pooled activation)  vs  kws_sim_stem_pool {tp:.3f} ms ({K * U / tp * 1e3:.0f} pairs/s): x{(t + tm) / tp:.3f}")


if __name__ == "__main__":
    main()
