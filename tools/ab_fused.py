"""Development aid: time kws_sim_stem of two builds of the library on the same GPU (A/B).
    python tools/ab_fused.py enhance-cb-whisper_b200/libkws_b200_old.so enhance-cb-whisper_b200/libkws_b200.so"""
import ctypes as C, sys
import torch

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
Cc, K, U, Tk, Tu, P = 12, 592, 2, 150, 1500, 64
unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1)
kn, un = unit(Cc, K, Tk, P).half(), unit(Cc, U, Tu, P).half()
w = (torch.randn(64, Cc, 7, 7, generator=g, device=dev) * 0.05).contiguous()
one, zero = torch.ones(64, device=dev), torch.zeros(64, device=dev)
out = torch.empty(K * U, 75, 750, 64, dtype=torch.bfloat16, device=dev)
vp, i32 = C.c_void_p, C.c_int
libs = []
for path in sys.argv[1:]:
    lib = C.CDLL(path)
    lib.kws_stem_fused_weight_bytes.restype = C.c_size_t
    lib.kws_pack_stem_fused.argtypes = [vp] * 5 + [C.c_float, i32, vp, vp, vp]
    lib.kws_sim_stem.argtypes = [vp, vp] + [i32] * 7 + [vp, vp, i32, vp, vp]
    wf = torch.empty(lib.kws_stem_fused_weight_bytes(Cc) // 2, dtype=torch.float16, device=dev)
    bias = torch.empty(64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.kws_pack_stem_fused(w.data_ptr(), one.data_ptr(), zero.data_ptr(), zero.data_ptr(), one.data_ptr(), 1e-5, Cc,
                                   wf.data_ptr(), bias.data_ptr(), st) == 0
    libs.append((path, lib, wf, bias))


def run(lib, wf, bias):
    rc = lib.kws_sim_stem(kn.data_ptr(), un.data_ptr(), Cc, K, U, Tk, Tu, P, 0, wf.data_ptr(), bias.data_ptr(), 1,
                          out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0


for rnd in range(3):
    for path, lib, wf, bias in libs:
        run(lib, wf, bias)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            run(lib, wf, bias)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"round {rnd} {path.split('/')[-1]:24s}: {ms:.3f} ms -> {K * U / ms * 1e3:.0f} pairs/s")
