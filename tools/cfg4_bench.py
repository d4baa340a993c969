"""Config #4 throughput (original CB-Whisper classifier path): ragged keywords x segments ->
similarity -> bilinear resize (150x750) -> 12-channel stem (in-scope end point), whisper-medium shape.
    python tools/cfg4_bench.py [--K 200] [--S 8]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enhance_cb_whisper_b200 import cbw, ops, Resnet

ap = argparse.ArgumentParser()
ap.add_argument("--K", type=int, default=200)
ap.add_argument("--S", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(4)
C, D, Tu = 12, 1024, 1500
lens = torch.randint(10, 61, (a.K,), generator=g, device=dev).tolist()
kwd_list = [torch.nn.functional.normalize(torch.randn(C, t, D, generator=g, device=dev), dim=-1) for t in lens]
utt = torch.nn.functional.normalize(torch.randn(a.S, C, Tu, D, generator=g, device=dev), dim=-1)
torch.manual_seed(0)
emb = Resnet(C, 2).feature_extractor.embedder.embedder.to(dev)
wp, bias = ops.pack_stem_weights(emb.convolution.weight, emb.normalization.weight, emb.normalization.bias,
                                 emb.normalization.running_mean, emb.normalization.running_var)


def run():
    _, f16 = cbw.similarity_images(kwd_list, utt, (150, 750), want_f32=False, want_f16=True)
    flat = f16.view(a.K * a.S, *f16.shape[2:])
    for p0 in range(0, a.K * a.S, 256):
        ops.stem(flat[p0:p0 + 256], 750, wp, bias, ops.STEM_OUT_NHWC_BF16)


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
pairs = a.K * a.S
fl = sum(2.0 * C * t * Tu * D for t in lens) * a.S + 2.0 * 64 * 49 * C * 75 * 375 * pairs
print(f"cfg4 {a.K} ragged keywords (10..60 frames) x {a.S} segments: {ms:.2f} ms -> {pairs / ms * 1e3:.0f} pairs/s, "
      f"{fl / ms / 1e9:.1f} TFLOP/s algorithmic (similarity + stem)")
