"""Config #4 phase breakdown (development aid): pack keywords / normalise / similarity / resize / stem."""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enhance_cb_whisper_b200 import cbw, ops, Resnet

ap = argparse.ArgumentParser()
ap.add_argument("--K", type=int, default=1000)
ap.add_argument("--S", type=int, default=16)
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


def timed(fn, n=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


t_pack, (kwd_n, lens_t) = timed(lambda: cbw.pack_keywords(kwd_list, dev))
t_norm, utt_n = timed(lambda: ops.normalize_rows(utt.float().contiguous(), list(range(C)), None))
kb = 64
t_sim, (f32, _) = timed(lambda: ops.sim(kwd_n[:, :kb].contiguous(), utt_n, want_f32=True, want_f16=False))
t_rs, (_, f16) = timed(lambda: ops.resize_bilinear(f32, lens_t[:kb].contiguous(), (150, 750), want_f32=False, want_f16=True))
flat = f16.view(kb * a.S, *f16.shape[2:])
t_st, _ = timed(lambda: ops.stem(flat[:256], 750, wp, bias, ops.STEM_OUT_NHWC_BF16))
pairs = kb * a.S
print(f"pack {a.K} keywords {t_pack:.1f} ms | normalise {a.S} segments {t_norm:.2f} ms | per {pairs} pairs: sim {t_sim:.2f} ms, "
      f"resize {t_rs:.2f} ms, stem(256 pairs) {t_st:.2f} ms -> {t_st * pairs / 256:.2f} ms")
print(f"  per pair: sim {t_sim / pairs * 1e3:.2f} us, resize {t_rs / pairs * 1e3:.2f} us, stem {t_st / 256 * 1e3:.2f} us; "
      f"pack amortised over {a.S} segments {t_pack / (a.K * a.S) * 1e3:.2f} us")
