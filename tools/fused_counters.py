"""Development aid: per-CTA cycle breakdown of the fused kernel's MMA issuer (waits vs issue)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the hooks live in the development flavour only: KWS_FUSED_TIMERS=1 python enhance-cb-whisper_b200/build.py --debug-hooks
os.environ.setdefault("KWS_B200_LIB", os.path.join(ROOT, "enhance-cb-whisper_b200", "libkws_b200_dbg.so"))
from enhance_cb_whisper_b200 import ops, _lib

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
# usage: fused_counters.py [C Tk Tu K U [Dk]]   (default: the cfg2 shape; C > 12 reports the LAST channel-group pass)
Cc, Tk, Tu, K, U = (int(a) for a in sys.argv[1:6]) if len(sys.argv) >= 6 else (12, 150, 1500, 74, 2)
P = int(sys.argv[6]) if len(sys.argv) >= 7 else 64
Ho, Wo = (Tk + 1) // 2, (Tu + 1) // 2
unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1)
kn, un = unit(Cc, K, Tk, P).half(), unit(Cc, U, Tu, P).half()
wp, bias = ops.pack_stem_fused(torch.randn(64, Cc, 7, 7, generator=g, device=dev) * 0.05, torch.ones(64, device=dev),
                                 torch.zeros(64, device=dev), torch.zeros(64, device=dev), torch.ones(64, device=dev))
out = torch.empty(K * U, Ho, Wo, 64, dtype=torch.bfloat16, device=dev)
ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NHWC_BF16, out=out)
buf = torch.zeros(148, 32, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.kws_debug_set_fused_counters.argtypes = [ctypes.c_void_p]
lib.kws_debug_set_fused_counters(buf.data_ptr())
ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NHWC_BF16, out=out)
torch.cuda.synchronize()
lib.kws_debug_set_fused_counters(None)
b = buf.double().mean(0).tolist()
items = K * U * ((Wo + 59) // 60) / 148
steps = (Ho + 1) // 2
print(f"C={Cc} {Tk}x{Tu} K={K} U={U}: {steps} steps/item")
print(f"per item (mean over CTAs, {items:.1f} items/CTA): total {b[0]/items:.0f} cyc | deadline-sim wait {b[1]/items:.0f} | "
      f"aempty wait {b[2]/items:.0f} | qfull wait {b[3]/items:.0f} | issue {b[4]/items:.0f}")
print(f"  epilogue: afull wait {b[5]/items:.0f} | until acc released {b[6]/items:.0f} | whole step body {b[7]/items:.0f}")
print(f"  converter: sfull wait {b[8]/items:.0f} | tmem ld {b[9]/items:.0f} | qempty wait {b[10]/items:.0f} | pack+store {b[11]/items:.0f}")
print(f"  multi-pass epilogue: partial-sum load wait {b[12]/items:.0f} | LDS of the partial tile {b[13]/items:.0f}")
print("  epilogue phases per item: " + " | ".join(f"{n} {b[16+i]/items:.0f}" for i, n in enumerate(
    ["g0 ld+wait", "g0 mbox+bar", "g0 shfl+math+sts", "g1 ld+wait", "g1 mbox+bar", "g1 shfl+math+sts", "fence+tma"])))
