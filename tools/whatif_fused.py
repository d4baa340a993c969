"""Development aid: what bounds the fused kernel?  Times one shape with parts of the roles switched off
(kws_debug_set_fused_whatif, debug-hooks flavour only; the results of those runs are WRONG by construction).

    python enhance-cb-whisper_b200/build.py --debug-hooks
    python tools/whatif_fused.py cfg2|cfg1 [K U]
bits: 1 epilogue releases the accumulator at once and does nothing else | 2 no half-b shift (mailbox, shuffles) |
      4 no TMA store | 8 stem issues kernel rows 0 and 3 only | 16 similarity issues one of four k-steps per stage
"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("KWS_B200_LIB", os.path.join(ROOT, "enhance-cb-whisper_b200", "libkws_b200_dbg.so"))
from enhance_cb_whisper_b200 import ops, _lib

SHAPES = {"cfg2": (12, 64, 150, 1500), "cfg1": (4, 384, 150, 1500), "lef12": (12, 64, 75, 750), "cfg3": (32, 64, 75, 750)}
what = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
Cc, Dk, Tk, Tu = SHAPES[what]
K, U = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (592, 2)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1)
kn, un = unit(Cc, K, Tk, Dk).half(), unit(Cc, U, Tu, Dk).half()
one, zero = torch.ones(64, device=dev), torch.zeros(64, device=dev)
wp, bias = ops.pack_stem_fused(torch.randn(64, Cc, 7, 7, generator=g, device=dev) * 0.05, one, zero, zero, one)
out = torch.empty(K * U, (Tk + 1) // 2, (Tu + 1) // 2, 64, dtype=torch.bfloat16, device=dev)
lib = _lib.load()
run = lambda: ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NHWC_BF16, out=out)
ITERS = int(os.environ.get("KWS_WHATIF_ITERS", "3"))  # >= 300: sustained (power-capped) regime instead of a burst
combos = [0, 8, 16, 32, 1, 2, 4, 0] if ITERS > 50 else [0, 32, 32 | 1, 32 | 8 | 16, 32 | 1 | 8 | 16, 1 | 8 | 16, 0] if os.environ.get('KWS_WHATIF_L2') else [0, 1, 2, 4, 2 | 4, 8, 16, 8 | 16, 1 | 8, 1 | 16, 1 | 8 | 16, 0]
if Cc > 12:
    combos = [0, 8, 16, 32, 2, 0] if ITERS > 50 else [0, 32, 2, 32 | 2, 0] if os.environ.get('KWS_WHATIF_L2') else [0, 2, 8, 16, 8 | 16, 0]  # multi-pass: the epilogue's partial-sum protocol must stay intact
for bits in combos:
    lib.kws_debug_set_fused_whatif(bits)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(ITERS):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / ITERS
    print(f"{what} whatif={bits:2d}: {ms:.3f} ms -> {K * U / ms * 1e3:.0f} pairs/s", flush=True)
lib.kws_debug_set_fused_whatif(0)
if Cc > 12 and Cc % 12 != 0 and Cc % 12 <= 8:  # last pass of <= 8 layers: second staging set + one-step-ahead partial-sum loads
    for on in (0, 1, 0, 1):
        lib.kws_debug_set_fused_last_pf1(on)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{what} last-pass prefetch into a second staging set = {on}: {ms:.3f} ms -> {K * U / ms * 1e3:.0f} pairs/s", flush=True)
    lib.kws_debug_set_fused_last_pf1(0)
