"""Development aid: event trace of CTA 0 of the fused kernel (build with KWS_FUSED_TIMERS=1)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the hooks live in the development flavour only: KWS_FUSED_TIMERS=1 python enhance-cb-whisper_b200/build.py --debug-hooks
os.environ.setdefault("KWS_B200_LIB", os.path.join(ROOT, "enhance-cb-whisper_b200", "libkws_b200_dbg.so"))
from enhance_cb_whisper_b200 import ops, _lib

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
# usage: fused_trace.py [grid_limit [C Dk]]
Cc, K, U, Tk, Tu, P = 12, 74, 2, 150, 1500, 64
if len(sys.argv) > 3:
    Cc, P = int(sys.argv[2]), int(sys.argv[3])
unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1)
kn, un = unit(Cc, K, Tk, P).half(), unit(Cc, U, Tu, P).half()
wp, bias = ops.pack_stem_fused(torch.randn(64, Cc, 7, 7, generator=g, device=dev) * 0.05, torch.ones(64, device=dev),
                               torch.zeros(64, device=dev), torch.zeros(64, device=dev), torch.ones(64, device=dev))
out = torch.empty(K * U, 75, 750, 64, dtype=torch.bfloat16, device=dev)
ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NHWC_BF16, out=out)
TS = 640
buf = torch.zeros(148 * 32 + 4 * TS * 8, dtype=torch.int64, device=dev)
lib = _lib.load()
if len(sys.argv) > 1 and int(sys.argv[1]) > 0:
    lib.kws_debug_set_fused_grid_limit(int(sys.argv[1]))
    print('grid limit', sys.argv[1])
lib.kws_debug_set_fused_counters.argtypes = [ctypes.c_void_p]
lib.kws_debug_set_fused_counters(buf.data_ptr())
ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NHWC_BF16, out=out)
torch.cuda.synchronize()
lib.kws_debug_set_fused_counters(None)
tr = buf[148 * 32:].view(4, TS, 8).cpu().numpy()
iss, epi, chk, qua = tr
t0 = iss[0, 0]
nP = 38
print("step | start aempty q<=P+1 di6pre simdone di6post end | retire(afull seen) accfree epidone | step_len retire_gap")
prev_ret = None
for s in range(2 * nP, 5 * nP):  # items 2..4 of CTA 0 (steady state)
    r = iss[s] - t0
    e = epi[s] - t0
    gap = (e[0] - prev_ret) if prev_ret is not None else 0
    prev_ret = e[0]
    print(f"{s:4d} P={s % nP:2d} | {r[0]:8d} +{r[1]-r[0]:5d} +{r[2]-r[1]:5d} +{r[3]-r[2]:5d} +{r[4]-r[3]:5d} +{r[5]-r[4]:5d} +{r[6]-r[5]:5d} | "
          f"{e[0]:8d} +{e[1]-e[0]:5d} +{e[2]-e[1]:5d} | len {iss[s+1,0]-iss[s,0]:5d} gap {gap:5d}")
print("chunk | sim issuer: start, +sempty0 wait, +ofull wait, last stage issued(+) | converter: first seen, pulled(+) | quanta: slot_free stored(+)")
for c in range(20, 32):
    print(f"chunk {c}: sim {chk[c,2]-t0:8d} +{chk[c,3]-chk[c,2]:5d} +{chk[c,4]-chk[c,3]:5d} +{chk[c,5]-chk[c,4]:5d} | conv {chk[c,0]-t0:8d} +{chk[c,1]-chk[c,0]:5d}   " + "  ".join(f"q{4*c+j}: {qua[4*c+j,0]-t0:8d} +{qua[4*c+j,1]-qua[4*c+j,0]:4d}" for j in range(4)))

print("stage | producer: slot free seen, +loads issued | sim issuer: operands seen (TMA latency from issue), +MMAs issued | slot round trip")
for i in range(300, 340):
    pf, pi, so, sm = qua[i, 2] - t0, qua[i, 3] - t0, qua[i, 4] - t0, qua[i, 5] - t0
    nxt = qua[i + 3, 2] - t0
    print(f"stage {i}: prod {pf:8d} +{pi - pf:4d} | sim {so:8d} (tma {so - pi:5d}) +{sm - so:4d} | commit->slot free seen {nxt - sm:5d} | round trip {qua[i+3,4]-qua[i,4]:5d}")
