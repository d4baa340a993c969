"""Config #4 throughput (original CB-Whisper classifier path), whisper-medium shape: ragged keywords x segments ->
stem activation (the in-scope end point).  Two paths:
  fused   : kws_interp_rows (width map on the utterance frames) -> kws_sim_operand (native similarity as an fp16
            operand) + kws_resize_row_weights (height map) -> kws_sim_stem(KWS_PAIRS_PER_KEYWORD); the resized image
            is never built.  The keyword bank is packed once (resident), as in serving.
  unfused : kws_sim (fp32 images) -> kws_resize_bilinear -> kws_stem
    python tools/cfg4_bench.py [--K 1000] [--S 16] [--unfused]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enhance_cb_whisper_b200 import cbw, ops, Resnet

ap = argparse.ArgumentParser()
ap.add_argument("--K", type=int, default=1000)
ap.add_argument("--S", type=int, default=16)
ap.add_argument("--max-pairs", type=int, default=1184)
ap.add_argument("--unfused", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(4)
C, D, Tu = 12, 1024, 1500
lens = torch.randint(10, 61, (a.K,), generator=g, device=dev).tolist()
kwd_list = [torch.nn.functional.normalize(torch.randn(C, t, D, generator=g, device=dev), dim=-1) for t in lens]
utt = torch.nn.functional.normalize(torch.randn(a.S, C, Tu, D, generator=g, device=dev), dim=-1)
torch.manual_seed(0)
sp = cbw.CBWKeywordSpotterB200(Resnet(C, 2).to(dev), size=(150, 750), body_dtype="bfloat16")
emb = sp.resnet.feature_extractor.embedder.embedder
wp, bias = ops.pack_stem_weights(emb.convolution.weight, emb.normalization.weight, emb.normalization.bias,
                                 emb.normalization.running_mean, emb.normalization.running_var)
kwd_n, lens_t = cbw.pack_keywords(kwd_list, dev, multiple=64)  # resident bank, built once per vocabulary
ev = lambda: torch.cuda.Event(enable_timing=True)
phases = {}


def run_fused():
    e = [ev() for _ in range(3)]
    e[0].record()
    utt_i = ops.interp_rows(utt, list(range(C)), 750)
    e[1].record()
    sp.stem_fused(kwd_n, lens_t, utt_i, ops.STEM_OUT_NHWC_BF16, max_pairs=a.max_pairs)
    e[2].record()
    return e


def run_unfused():
    e = [ev() for _ in range(2)]
    e[0].record()
    _, f16 = cbw.similarity_images(kwd_list, utt, (150, 750), want_f32=False, want_f16=True)
    flat = f16.view(a.K * a.S, *f16.shape[2:])
    for p0 in range(0, a.K * a.S, 256):
        ops.stem(flat[p0:p0 + 256], 750, wp, bias, ops.STEM_OUT_NHWC_BF16)
    e[1].record()
    return e


run = run_unfused if a.unfused else run_fused
run()
torch.cuda.synchronize()
e0, e1 = ev(), ev()
e0.record()
for _ in range(3):
    e = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
pairs = a.K * a.S
fl = sum(2.0 * C * t * Tu * D for t in lens) * a.S + 2.0 * 64 * 49 * C * 75 * 375 * pairs
extra = "" if a.unfused else f" (utterance resample {e[0].elapsed_time(e[1]):.2f} ms, similarity+resize+stem {e[1].elapsed_time(e[2]):.2f} ms)"
print(f"cfg4 {'unfused' if a.unfused else 'fused'} {a.K} ragged keywords (10..60 frames) x {a.S} segments: {ms:.2f} ms -> "
      f"{pairs / ms * 1e3:.0f} pairs/s, {fl / ms / 1e9:.1f} TFLOP/s algorithmic (reference-resolution similarity + stem){extra}")
