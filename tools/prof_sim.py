"""Stand-alone similarity GEMM (kws_sim) at the L-variant shapes (Dk = D): timing + ncu target.
    python tools/prof_sim.py [--D 1280] [--C 3] [--K 50] [--U 4]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enhance_cb_whisper_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=1280)
ap.add_argument("--C", type=int, default=3)
ap.add_argument("--K", type=int, default=50)
ap.add_argument("--U", type=int, default=4)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
Tk, Tu = 150, 1500
unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1)
kn, un = unit(a.C, a.K, Tk, a.D).half(), unit(a.C, a.U, Tu, a.D).half()
f16 = torch.empty(a.K, a.U, a.C, Tk, ops.pitch_for(Tu), dtype=torch.float16, device=dev)
ops.sim(kn, un, False, True, out_f16=f16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(a.iters):
    ops.sim(kn, un, False, True, out_f16=f16)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
fl = 2.0 * a.C * Tk * Tu * a.D * a.K * a.U
print(f"kws_sim L D={a.D} C={a.C} {a.K}x{a.U} pairs: {ms:.3f} ms -> {fl / ms / 1e9:.1f} TFLOP/s (algorithmic, Tk=150 of a 160-wide tile), "
      f"{a.K * a.U / ms * 1e3:.0f} pairs/s, fp16 out {f16.numel() * 2 / ms / 1e6:.0f} GB/s")
