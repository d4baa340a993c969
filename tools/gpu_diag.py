"""Kernel-by-kernel diagnostic on a real B200 (run under gpurun, not a pytest):

    python tests/gpu_diag.py [--quick] > gpurun_out/diag.log

Each stage is compared against torch arithmetic on the SAME quantised operands
(isolates kernel bugs from rounding) and, end to end, against the CPU oracle.
On a mismatch it prints sub-blocks and runs one-hot probes that localise
descriptor / swizzle / indexing errors.  Also prints first timings.
"""
from __future__ import annotations

import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import enhance_cb_whisper_b200 as kb  # noqa: E402
from enhance_cb_whisper_b200 import ops  # noqa: E402
from oracle import kws_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
FAILS = []


def report(name, got, exp, tol):
    got, exp = got.float(), exp.float()
    err = (got - exp).abs()
    mx = err.max().item() if err.numel() else 0.0
    bad = int((err > tol).sum().item())
    nan = int(torch.isnan(got).sum().item())
    status = "OK " if (mx <= tol and nan == 0) else "FAIL"
    print(f"[{status}] {name}: max|err|={mx:.3e} tol={tol:.1e} bad={bad}/{err.numel()} nan={nan} "
          f"|exp|max={exp.abs().max().item():.3e}", flush=True)
    if status == "FAIL":
        FAILS.append(name)
        idx = (err > tol).nonzero()[:6].tolist()
        print("   first bad idx:", idx)
        for i in idx[:3]:
            print("    got", got[tuple(i)].item(), "exp", exp[tuple(i)].item())
    return status == "OK "


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def unit_rows(*shape, gen):
    x = torch.randn(*shape, generator=gen, device=dev)
    return x / x.norm(dim=-1, keepdim=True)


def test_prep(gen):
    print("== prep: normalize_rows / cast_rows ==")
    B, Cin, T, D = 3, 5, 37, 384
    x = torch.randn(B, Cin, T, D, generator=gen, device=dev)
    x[1, :, 30:] = 0
    lidx = [4, 1, 2]
    mask = (torch.rand(B, 3, T, generator=gen, device=dev) > 0.2).float()
    out = ops.normalize_rows(x, lidx, mask)
    xs = x[:, lidx].permute(1, 0, 2, 3)
    exp = xs / xs.norm(dim=-1, keepdim=True).clamp(min=1e-6) * mask.permute(1, 0, 2)[..., None]
    report("normalize_rows", out, exp, 1e-3)
    ob = ops.cast_rows16(x, lidx, ops.BF16)
    report("cast_rows16 bf16", ob.view(3, B, T, D), xs.bfloat16(), 0.0)
    oh = ops.cast_rows16(x, lidx, ops.F16)
    report("cast_rows16 f16", oh.view(3, B, T, D), xs.half(), 0.0)


def gemm_probe(Dk, Tk, Tu):
    """one-hot probes: A[j,k]=(k==k0), B[i,k]=(i+1)*(k==k0)  ->  D[j,i] = i+1"""
    for k0 in (0, 7, 8, 15, 16, 31, 32, 63, Dk - 1):
        a = torch.zeros(1, 1, Tu, Dk, device=dev, dtype=torch.float16)
        b = torch.zeros(1, 1, Tk, Dk, device=dev, dtype=torch.float16)
        a[..., k0] = 1
        b[0, 0, :, k0] = torch.arange(1, Tk + 1, device=dev, dtype=torch.float16)
        f32, _ = ops.sim(b, a, True, False)
        exp = torch.arange(1, Tk + 1, device=dev, dtype=torch.float32)[:, None].expand(Tk, Tu)
        e = (f32[0, 0, 0] - exp).abs().max().item()
        print(f"   probe k0={k0}: max err {e:.3e}; row0[:8]={f32[0,0,0,:8,0].tolist()}")


def test_sim(gen, shapes):
    print("== sim GEMM (tcgen05) vs torch.matmul on the same fp16 operands ==")
    for (Cc, K, U, Tk, Tu, Dk) in shapes:
        kn = unit_rows(Cc, K, Tk, Dk, gen=gen).half()
        un = unit_rows(Cc, U, Tu, Dk, gen=gen).half()
        f32, f16 = ops.sim(kn, un, True, True)
        exp = torch.einsum("ckid,cujd->kucij", kn.float(), un.float())
        ok = report(f"sim f32 C{Cc} K{K} U{U} Tk{Tk} Tu{Tu} Dk{Dk}", f32, exp, 2e-4)
        report("sim f16", f16[..., :Tu], exp, 1e-3)
        if not ok:
            print("   got[0,0,0,:4,:6]\n", f32[0, 0, 0, :4, :6], "\n   exp\n", exp[0, 0, 0, :4, :6])
            gemm_probe(Dk, Tk, Tu)
    # diagonal pairing
    Cc, K, Tk, Tu, Dk = 2, 3, 20, 70, 64
    kn = unit_rows(Cc, K, Tk, Dk, gen=gen).half()
    un = unit_rows(Cc, K, Tu, Dk, gen=gen).half()
    f32, _ = ops.sim(kn, un, True, False, diag=True)
    exp = torch.einsum("ckid,ckjd->kcij", kn.float(), un.float())
    report("sim diag", f32, exp, 2e-4)


def test_mlp(gen):
    print("== MLP (two tcgen05 GEMMs) ==")
    for (Cc, B, T, D, P, d16) in [(2, 3, 50, 128, 64, ops.BF16), (3, 2, 150, 768, 64, ops.F16),
                                  (1, 1, 300, 256, 32, ops.F16)]:
        H = D // 2
        tdt = ops.TORCH16[d16]
        x = unit_rows(B, Cc, T, D, gen=gen)
        w1 = torch.randn(Cc, H, D, generator=gen, device=dev) * (4.0 / D ** 0.5)
        b1 = torch.randn(Cc, H, generator=gen, device=dev) * 0.05
        w2 = torch.randn(Cc, P, H, generator=gen, device=dev) / H ** 0.5
        b2 = torch.randn(Cc, P, generator=gen, device=dev) * 0.05
        mask = (torch.rand(B, Cc, T, generator=gen, device=dev) > 0.1).float()
        xb = ops.cast_rows16(x, list(range(Cc)), d16)
        w1b, w2b = ops.cast16(w1, d16), ops.cast16(w2, d16)
        raw = ops.mlp(xb, B, T, w1b, b1, w2b, b2, None, ops.MLP_OUT_RAW_F32)
        # same-quantisation expectation: bf16 x, bf16 W1, fp32 acc, bf16 hidden, bf16 W2
        xq = xb.float().view(Cc, B, T, D)
        h = torch.relu(torch.einsum("cbtd,chd->cbth", xq, w1b.float()) + b1[:, None, None, :]).to(tdt).float()
        exp = torch.einsum("cbth,cph->cbtp", h, w2b.float()) + b2[:, None, None, :]
        ok = report(f"mlp raw C{Cc} B{B} T{T} D{D} P{P} {tdt}", raw, exp, 2e-3)
        nrm = ops.mlp(xb, B, T, w1b, b1, w2b, b2, mask, ops.MLP_OUT_NORM_F16)
        expn = exp / exp.norm(dim=-1, keepdim=True).clamp(min=1e-6) * mask.permute(1, 0, 2)[..., None]
        report("mlp norm", nrm, expn, 2e-3)
        # against the true fp32 MLP
        hf = torch.relu(torch.einsum("bctd,chd->cbth", x, w1) + b1[:, None, None, :])
        ef = torch.einsum("cbth,cph->cbtp", hf, w2) + b2[:, None, None, :]
        print(f"     vs fp32 MLP: max|err| {(raw - ef).abs().max().item():.3e} (|out| max {ef.abs().max().item():.2f})")
        if not ok:
            print("   got\n", raw[0, 0, :3, :6], "\n   exp\n", exp[0, 0, :3, :6])


def test_temporal(gen):
    print("== temporal (conv1d+BN+maxpool+normalise) ==")
    for (Cc, B, T, P) in [(2, 3, 23, 64), (3, 2, 150, 64), (1, 1, 71, 32)]:
        sd = O.make_weights("LEF", Cc, 128, P, seed=5)
        proj = torch.randn(Cc, B, T, P, generator=gen, device=dev).half()
        st = lambda n: torch.stack([sd[f"time_projector.{i}.{n}"] for i in range(Cc)]).to(dev)
        wf, bf = ops.fold_temporal_weights(st("0.weight"), st("0.bias"), st("1.weight"), st("1.bias"),
                                           st("1.running_mean"), st("1.running_var"))
        T2 = (T + 1) // 2
        mask = (torch.rand(B, Cc, T2, generator=gen, device=dev) > 0.1).float()
        out = ops.temporal(proj, wf, bf, mask)
        exp = torch.stack([O.project_time(proj[i].float().cpu(), sd, i) for i in range(Cc)]).to(dev)
        exp = exp / exp.norm(dim=-1, keepdim=True).clamp(min=1e-6) * mask.permute(1, 0, 2)[..., None]
        report(f"temporal C{Cc} B{B} T{T} P{P}", out, exp, 2e-3)


def stem_expect(f16, Tu, sd, quant=True):
    """conv on the same fp16 features with fp16-rounded folded weights (fp64 accumulate)."""
    w = sd[O.STEM + "convolution.weight"].double()
    g, b = sd[O.STEM + "normalization.weight"].double(), sd[O.STEM + "normalization.bias"].double()
    m, v = sd[O.STEM + "normalization.running_mean"].double(), sd[O.STEM + "normalization.running_var"].double()
    s = g / torch.sqrt(v + 1e-5)
    wq = (w * s[:, None, None, None])
    if quant:
        wq = wq.float().half().double()
    x = f16[..., :Tu].double().flatten(0, -4)
    y = torch.nn.functional.conv2d(x, wq.to(x.device), (b - m * s).to(x.device), stride=2, padding=3)
    return torch.relu(y).float()


def test_stem(gen, shapes):
    print("== stem (tap-decomposed tcgen05 implicit GEMM) ==")
    for (N, Cc, Tk, Tu) in shapes:
        sd = O.make_weights("L", Cc, 64, seed=9)
        sdd = {k: v.to(dev) for k, v in sd.items()}
        pitch = ops.pitch_for(Tu)
        f16 = (torch.rand(N, Cc, Tk, pitch, generator=gen, device=dev) * 2 - 1).half()
        wp, bias = ops.pack_stem_weights(sdd[O.STEM + "convolution.weight"], sdd[O.STEM + "normalization.weight"],
                                         sdd[O.STEM + "normalization.bias"],
                                         sdd[O.STEM + "normalization.running_mean"],
                                         sdd[O.STEM + "normalization.running_var"])
        out = ops.stem(f16, Tu, wp, bias, ops.STEM_OUT_NCHW_F32)
        exp = stem_expect(f16, Tu, sdd)
        ok = report(f"stem nchw N{N} C{Cc} Tk{Tk} Tu{Tu}", out, exp, 2e-4)
        o2 = ops.stem(f16, Tu, wp, bias, ops.STEM_OUT_NHWC_BF16)
        report("stem nhwc bf16", o2, exp, 2e-2)
        if not ok:
            print("   got[0,0,:3,:6]\n", out[0, 0, :3, :6], "\n   exp\n", exp[0, 0, :3, :6])
            # single-tap probes: weight one-hot at (oc=0,c=0,di,dj)
            for (di, dj) in [(3, 3), (0, 0), (3, 4), (4, 3), (6, 6)]:
                w = torch.zeros(64, Cc, 7, 7, device=dev)
                w[:, 0, di, dj] = torch.arange(1, 65, device=dev).float()
                one, zero = torch.ones(64, device=dev), torch.zeros(64, device=dev)
                wp1, b1 = ops.pack_stem_weights(w, one, zero, zero, one - 1e-5)
                o = ops.stem(f16, Tu, wp1, b1, ops.STEM_OUT_NCHW_F32)
                e = torch.relu(torch.nn.functional.conv2d(f16[..., :Tu].float(), w, None, stride=2, padding=3))
                print(f"   tap probe (di={di},dj={dj}): max err {(o - e).abs().max().item():.3e}")


def test_e2e():
    print("== end to end vs committed golden fixtures (reference forward) ==")
    from oracle.make_golden import CASES, build_body, load_case

    for name in CASES:
        meta, ins, sd, outs = load_case(name)
        v = meta["variant"]
        m = kb.KWSModelB200(n_layers=meta["C"], embedding_dim=meta["D"], proj_mlp_units=meta["P"],
                            learn_features=v != "L", proj_mlp=v != "L", frames_conv=v == "LEF",
                            resnet_version=meta["resnet_version"], features_size=(meta["Tk"], meta["Tu"]))
        fe, head = build_body(meta["C"], meta["resnet_version"], meta["body_seed"])
        full = dict(m.state_dict())
        full.update({"model.feature_extractor." + k: t for k, t in fe.state_dict().items()})
        full.update({"model.classifier." + k: t for k, t in head.state_dict().items()})
        full.update(sd)
        m.load_state_dict(full)
        m = m.to(dev).eval()
        kwd, utt = ins["kwd"].to(dev), ins["utt"].to(dev)
        km, um = ins["kwd_mask"].to(dev), ins["utt_mask"].to(dev)
        if v == "LEF":
            km, um = km[..., ::2].contiguous(), um[..., ::2].contiguous()
        for u in range(meta["U"]):
            r = m(kwd_features=kwd, utt_features=utt[u:u + 1], kwd_mask=km, utt_mask=um[u:u + 1])
            report(f"{name} u{u} features", r.features, outs["features"][:, u].to(dev), 2e-3)
            report(f"{name} u{u} logits", r.logits, outs["logits"][:, u].to(dev), 2e-3)


def bench_sizes(gen):
    print("== first timings (cfg2-like slabs) ==")
    Cc, K, U, Tk, Tu, Dk = 12, 64, 4, 150, 1500, 64
    kn = unit_rows(Cc, K, Tk, Dk, gen=gen).half()
    un = unit_rows(Cc, U, Tu, Dk, gen=gen).half()
    pitch = ops.pitch_for(Tu)
    f16 = torch.empty(K, U, Cc, Tk, pitch, dtype=torch.float16, device=dev)
    t = timeit(lambda: ops.sim(kn, un, False, True, out_f16=f16))
    pairs = K * U
    print(f"sim  LE  {pairs} pairs: {t:.3f} ms -> {pairs / t * 1e3:.0f} pairs/s, "
          f"{2 * Cc * Tk * Tu * Dk * pairs / t / 1e9:.1f} TFLOP/s, out {f16.numel() * 2 / t / 1e6:.0f} GB/s")
    sd = {k: v.to(dev) for k, v in O.make_weights("L", Cc, 64, seed=9).items()}
    wp, bias = ops.pack_stem_weights(sd[O.STEM + "convolution.weight"], sd[O.STEM + "normalization.weight"],
                                     sd[O.STEM + "normalization.bias"], sd[O.STEM + "normalization.running_mean"],
                                     sd[O.STEM + "normalization.running_var"])
    out = torch.empty(pairs, 75, 750, 64, dtype=torch.bfloat16, device=dev)
    t = timeit(lambda: ops.stem(f16, Tu, wp, bias, ops.STEM_OUT_NHWC_BF16, out=out))
    fl = 2 * 64 * 49 * Cc * 75 * 750 * pairs
    print(f"stem LE  {pairs} pairs: {t:.3f} ms -> {pairs / t * 1e3:.0f} pairs/s, {fl / t / 1e9:.1f} TFLOP/s (algorithmic)")
    # L variant GEMM, cfg1 shape
    Cc, K, U, Dk = 4, 32, 2, 384
    kn = unit_rows(Cc, K, Tk, Dk, gen=gen).half()
    un = unit_rows(Cc, U, Tu, Dk, gen=gen).half()
    f16 = torch.empty(K, U, Cc, Tk, pitch, dtype=torch.float16, device=dev)
    t = timeit(lambda: ops.sim(kn, un, False, True, out_f16=f16))
    print(f"sim  L   {K * U} pairs D=384: {t:.3f} ms -> {2 * Cc * Tk * Tu * Dk * K * U / t / 1e9:.1f} TFLOP/s")


STAGES = ["prep", "sim", "mlp", "temporal", "stem", "e2e", "bench"]


def run_stage(name):
    gen = torch.Generator(device=dev).manual_seed(1234)
    print("device:", torch.cuda.get_device_name(0), "SMs", ops.sm_count())
    if name == "prep":
        test_prep(gen)
    elif name == "sim":
        test_sim(gen, [(1, 1, 1, 16, 128, 64), (2, 3, 2, 22, 70, 64), (2, 2, 2, 150, 300, 128),
                       (1, 2, 1, 150, 1500, 384)])
    elif name == "mlp":
        test_mlp(gen)
    elif name == "temporal":
        test_temporal(gen)
    elif name == "stem":
        test_stem(gen, [(1, 3, 8, 40), (2, 3, 22, 70), (2, 12, 23, 301), (1, 16, 150, 1500)])
    elif name == "e2e":
        test_e2e()
    elif name == "bench":
        bench_sizes(gen)
    torch.cuda.synchronize()
    print(f"stage {name}: failures: {FAILS}")
    return 1 if FAILS else 0


def main():
    """Each stage runs in its own process (a trapped kernel kills the CUDA context) under a timeout."""
    import subprocess

    if "--stage" in sys.argv:
        sys.exit(run_stage(sys.argv[sys.argv.index("--stage") + 1]))
    stages = [s for s in STAGES if not ("--quick" in sys.argv and s == "bench")]
    rcs = {}
    for st in stages:
        print(f"######## stage {st}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], timeout=240)
            rcs[st] = r.returncode
        except subprocess.TimeoutExpired:
            rcs[st] = "timeout"
        print(f"######## stage {st} rc={rcs[st]}", flush=True)
    print("SUMMARY", rcs)
    sys.exit(0 if all(v == 0 for v in rcs.values()) else 1)


if __name__ == "__main__":
    main()
