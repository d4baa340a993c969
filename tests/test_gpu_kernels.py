"""GPU: each sm_100a kernel, called through the C ABI, against torch arithmetic on the SAME
quantised operands (isolates kernel errors from 16-bit rounding) and against the fp32 oracle."""
import pytest
import torch

from oracle import kws_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_lib, cuda_dev):
    from enhance_cb_whisper_b200 import ops as _ops

    return _ops


def gen(dev, seed=1234):
    return torch.Generator(device=dev).manual_seed(seed)


def unit_rows(*shape, g, dev):
    x = torch.randn(*shape, generator=g, device=dev)
    return x / x.norm(dim=-1, keepdim=True)


def maxerr(a, b):
    return (a.float() - b.float()).abs().max().item()


# ---- prep ---------------------------------------------------------------------------------
def test_normalize_rows_selects_normalises_masks(ops, cuda_dev):
    g = gen(cuda_dev)
    B, Cin, T, D = 3, 5, 37, 384
    x = torch.randn(B, Cin, T, D, generator=g, device=cuda_dev)
    x[1, :, 30:] = 0  # zero-padded frames: norm clamps at eps, output stays 0
    lidx = [4, 1, 2]
    mask = (torch.rand(B, 3, T, generator=g, device=cuda_dev) > 0.2).float()
    out = ops.normalize_rows(x, lidx, mask)
    xs = x[:, lidx].permute(1, 0, 2, 3)
    exp = xs / xs.norm(dim=-1, keepdim=True).clamp(min=1e-6) * mask.permute(1, 0, 2)[..., None]
    assert out.shape == (3, B, T, D) and out.dtype == torch.float16
    assert maxerr(out, exp) <= 1e-3
    assert torch.equal(out[:, 1, 30:], torch.zeros_like(out[:, 1, 30:]))


@pytest.mark.parametrize("d16", ["F16", "BF16"])
def test_cast_rows16_is_bit_exact(ops, cuda_dev, d16):
    g = gen(cuda_dev)
    x = torch.randn(2, 4, 19, 128, generator=g, device=cuda_dev) * 3
    code = getattr(ops, d16)
    out = ops.cast_rows16(x, [3, 0], code)
    exp = x[:, [3, 0]].permute(1, 0, 2, 3).to(ops.TORCH16[code]).reshape(2, 2 * 19, 128)
    assert torch.equal(out, exp)


def test_cast16_saturates_fp16(ops, cuda_dev):
    x = torch.tensor([1e6, -1e6, 1.0, 65504.0], device=cuda_dev)
    out = ops.cast16(x, ops.F16)
    assert torch.equal(out.float(), torch.tensor([65504.0, -65504.0, 1.0, 65504.0], device=cuda_dev))


# ---- similarity GEMM ------------------------------------------------------------------------
@pytest.mark.parametrize("Cc,K,U,Tk,Tu,Dk", [(1, 1, 1, 16, 128, 64), (2, 3, 2, 22, 70, 64), (2, 2, 2, 150, 300, 128),
                                             (1, 2, 1, 150, 1500, 384), (1, 1, 1, 75, 750, 64), (1, 1, 1, 1, 1, 64)])
def test_sim_matches_matmul(ops, cuda_dev, Cc, K, U, Tk, Tu, Dk):
    g = gen(cuda_dev)
    kn = unit_rows(Cc, K, Tk, Dk, g=g, dev=cuda_dev).half()
    un = unit_rows(Cc, U, Tu, Dk, g=g, dev=cuda_dev).half()
    f32, f16 = ops.sim(kn, un, True, True)
    exp = torch.einsum("ckid,cujd->kucij", kn.float(), un.float())
    assert maxerr(f32, exp) <= 2e-4
    assert maxerr(f16[..., :Tu], exp) <= 1e-3


def test_sim_diag_pairing(ops, cuda_dev):
    g = gen(cuda_dev)
    Cc, K, Tk, Tu, Dk = 2, 3, 20, 70, 64
    kn = unit_rows(Cc, K, Tk, Dk, g=g, dev=cuda_dev).half()
    un = unit_rows(Cc, K, Tu, Dk, g=g, dev=cuda_dev).half()
    f32, _ = ops.sim(kn, un, True, False, diag=True)
    exp = torch.einsum("ckid,ckjd->kcij", kn.float(), un.float())
    assert maxerr(f32, exp) <= 2e-4


def test_sim_linearity_and_zero_rows(ops, cuda_dev):
    """size-independent properties: zero operand rows give exactly 0; scaling an operand row by 2
    scales its similarity row by exactly 2 (power-of-two scaling is exact in fp16/fp32)."""
    g = gen(cuda_dev)
    Cc, K, U, Tk, Tu, Dk = 1, 2, 2, 33, 140, 64
    kn = unit_rows(Cc, K, Tk, Dk, g=g, dev=cuda_dev).half()
    un = unit_rows(Cc, U, Tu, Dk, g=g, dev=cuda_dev).half()
    kn[0, 1, 10:] = 0
    a, _ = ops.sim(kn, un, True, False)
    assert torch.equal(a[1, :, 0, 10:], torch.zeros_like(a[1, :, 0, 10:]))
    kn2 = kn.clone()
    kn2[0, 0, 5] *= 0.5
    b, _ = ops.sim(kn2, un, True, False)
    assert torch.equal(b[0, :, 0, 5], a[0, :, 0, 5] * 0.5)


# ---- projector MLP ----------------------------------------------------------------------------
@pytest.mark.parametrize("Cc,B,T,D,P,d16", [(2, 3, 50, 128, 64, "BF16"), (3, 2, 150, 768, 64, "F16"),
                                            (1, 1, 300, 256, 32, "F16"), (1, 2, 7, 1280, 64, "F16")])
def test_mlp_matches_same_quantisation_reference(ops, cuda_dev, Cc, B, T, D, P, d16):
    g = gen(cuda_dev)
    code = getattr(ops, d16)
    tdt = ops.TORCH16[code]
    H = D // 2
    x = unit_rows(B, Cc, T, D, g=g, dev=cuda_dev)
    w1 = torch.randn(Cc, H, D, generator=g, device=cuda_dev) * (4.0 / D ** 0.5)
    b1 = torch.randn(Cc, H, generator=g, device=cuda_dev) * 0.05
    w2 = torch.randn(Cc, P, H, generator=g, device=cuda_dev) / H ** 0.5
    b2 = torch.randn(Cc, P, generator=g, device=cuda_dev) * 0.05
    mask = (torch.rand(B, Cc, T, generator=g, device=cuda_dev) > 0.1).float()
    xb = ops.cast_rows16(x, list(range(Cc)), code)
    w1b, w2b = ops.cast16(w1, code), ops.cast16(w2, code)
    raw = ops.mlp(xb, B, T, w1b, b1, w2b, b2, None, ops.MLP_OUT_RAW_F32)
    xq = xb.float().view(Cc, B, T, D)
    h = torch.relu(torch.einsum("cbtd,chd->cbth", xq, w1b.float()) + b1[:, None, None, :]).to(tdt).float()
    exp = torch.einsum("cbth,cph->cbtp", h, w2b.float()) + b2[:, None, None, :]
    assert maxerr(raw, exp) <= 2e-3
    nrm = ops.mlp(xb, B, T, w1b, b1, w2b, b2, mask, ops.MLP_OUT_NORM_F16)
    expn = exp / exp.norm(dim=-1, keepdim=True).clamp(min=1e-6) * mask.permute(1, 0, 2)[..., None]
    assert maxerr(nrm, expn) <= 2e-3
    # against the true fp32 MLP of the oracle (model.py:92-104): fp16 operands stay within 5e-3 of |out| ~ 1
    hf = torch.relu(torch.einsum("bctd,chd->cbth", x, w1) + b1[:, None, None, :])
    ef = torch.einsum("cbth,cph->cbtp", hf, w2) + b2[:, None, None, :]
    assert maxerr(raw, ef) <= (2e-2 if d16 == "BF16" else 5e-3)


@pytest.mark.parametrize("Cc,Cin,B,T,D,P,d16", [
    (2, 2, 3, 50, 128, 64, "F16"),      # H = 64: one 64-column hidden block
    (3, 5, 2, 150, 768, 64, "F16"),     # cfg2 shape: H = 384 in one pass, two MMAs per k-step; layer selection
    (1, 1, 2, 7, 1280, 64, "F16"),      # cfg3 shape: H = 640 as two passes of 320
    (2, 3, 1, 131, 1024, 64, "BF16"),   # whisper-medium: H = 512 as two passes of 256; bf16 operands
    (1, 1, 1, 300, 256, 32, "F16"),     # P = 32
    (12, 12, 20, 150, 256, 64, "F16"),  # 288 work items on 148 persistent CTAs, 12 layer changes (W2 reloads)
    (2, 2, 1, 1, 512, 16, "F16"),       # a single row, P = 16
])
def test_mlp_fused_matches_two_call_path_and_reference(ops, cuda_dev, Cc, Cin, B, T, D, P, d16):
    """kws_mlp_fused (raw fp32 rows -> cast -> GEMM1 -> ReLU -> GEMM2 -> normalise * mask in one kernel, hidden on chip)
    against kws_cast_rows16 + kws_mlp (same quantisation points: 16-bit x, 16-bit hidden) and against torch on the
    same quantised operands; all three output modes."""
    g = gen(cuda_dev)
    code = getattr(ops, d16)
    tdt = ops.TORCH16[code]
    H = D // 2
    assert ops.mlp_fused_supported(D, H, P)
    x = unit_rows(B, Cin, T, D, g=g, dev=cuda_dev)
    lidx = list(range(Cin))[::-1][:Cc]
    w1 = torch.randn(Cc, H, D, generator=g, device=cuda_dev) * (4.0 / D ** 0.5)
    b1 = torch.randn(Cc, H, generator=g, device=cuda_dev) * 0.05
    w2 = torch.randn(Cc, P, H, generator=g, device=cuda_dev) / H ** 0.5
    b2 = torch.randn(Cc, P, generator=g, device=cuda_dev) * 0.05
    mask = (torch.rand(B, Cc, T, generator=g, device=cuda_dev) > 0.1).float()
    w1b, w2b = ops.cast16(w1, code), ops.cast16(w2, code)
    xb = ops.cast_rows16(x, lidx, code)
    xq = xb.float().view(Cc, B, T, D)
    h = torch.relu(torch.einsum("cbtd,chd->cbth", xq, w1b.float()) + b1[:, None, None, :]).to(tdt).float()
    exp = torch.einsum("cbth,cph->cbtp", h, w2b.float()) + b2[:, None, None, :]
    expn = exp / exp.norm(dim=-1, keepdim=True).clamp(min=1e-6) * mask.permute(1, 0, 2)[..., None]
    raw = ops.mlp_fused(x, lidx, w1b, b1, w2b, b2, None, ops.MLP_OUT_RAW_F32)
    assert raw.shape == (Cc, B, T, P) and raw.dtype == torch.float32
    assert maxerr(raw, exp) <= 2e-3
    nrm = ops.mlp_fused(x, lidx, w1b, b1, w2b, b2, mask, ops.MLP_OUT_NORM_F16)
    assert nrm.dtype == torch.float16 and maxerr(nrm, expn) <= 2e-3
    r16 = ops.mlp_fused(x, lidx, w1b, b1, w2b, b2, None, ops.MLP_OUT_RAW_16)
    assert r16.dtype == tdt and maxerr(r16, exp) <= (3e-2 if d16 == "BF16" else 4e-3)
    # the two-call path quantises at the same points and accumulates in the same order
    raw2 = ops.mlp(xb, B, T, w1b, b1, w2b, b2, None, ops.MLP_OUT_RAW_F32)
    nrm2 = ops.mlp(xb, B, T, w1b, b1, w2b, b2, mask, ops.MLP_OUT_NORM_F16)
    assert maxerr(raw, raw2) <= 1e-5 and maxerr(nrm, nrm2) <= 1e-3
    with pytest.raises(Exception):  # a mask of another shape is refused, not mis-indexed
        ops.mlp_fused(x, lidx, w1b, b1, w2b, b2, mask[:, :, :-1].contiguous() if T > 1 else mask[:, :1, :0], ops.MLP_OUT_NORM_F16)


# ---- LEF temporal projector -------------------------------------------------------------------
@pytest.mark.parametrize("d16", ["F16", "BF16"])
@pytest.mark.parametrize("Cc,B,T,P", [(2, 3, 23, 64), (3, 2, 150, 64), (1, 1, 71, 32), (1, 2, 1, 64), (1, 1, 2, 64),
                                      (2, 5, 124, 64), (1, 3, 125, 128), (2, 40, 150, 64)])
def test_temporal_matches_oracle(ops, cuda_dev, Cc, B, T, P, d16):
    """Conv1d(3)+BN+MaxPool1d(3,2,1)+normalise+mask as a tcgen05 implicit GEMM over flat frames (tiles of 126 rows
    crossing item boundaries: T = 124, 125 put item edges on tile edges) vs the oracle's project_time on the same
    16-bit-rounded projections, and vs the oracle fed 16-bit-rounded folded weights (tight)."""
    g = gen(cuda_dev)
    dt = getattr(ops, d16)
    sd = O.make_weights("LEF", Cc, 128, P, seed=5)
    proj = torch.randn(Cc, B, T, P, generator=g, device=cuda_dev).to(ops.TORCH16[dt])
    st = lambda n: torch.stack([sd[f"time_projector.{i}.{n}"] for i in range(Cc)]).to(cuda_dev)
    wf, bf = ops.fold_temporal_weights(st("0.weight"), st("0.bias"), st("1.weight"), st("1.bias"),
                                       st("1.running_mean"), st("1.running_var"), dtype16=dt)
    assert wf.shape == (Cc, 3, P // 8, P, 8) and wf.dtype == ops.TORCH16[dt]
    T2 = (T + 1) // 2
    mask = (torch.rand(B, Cc, T2, generator=g, device=cuda_dev) > 0.1).float()
    out = ops.temporal(proj, wf, bf, mask)
    assert out.shape == (Cc, B, T2, P) and out.dtype == torch.float16
    norm = lambda y: y / y.norm(dim=-1, keepdim=True).clamp(min=1e-6) * mask.permute(1, 0, 2)[..., None]
    # (1) the reference arithmetic (fp32 weights) on the same rounded projections
    exp = norm(torch.stack([O.project_time(proj[i].float().cpu(), sd, i) for i in range(Cc)]).to(cuda_dev))
    assert maxerr(out, exp) <= (1e-2 if d16 == "BF16" else 2e-3)
    # (2) same quantisation: folded weights rounded to the operand type, fp64 accumulation
    s = st("1.weight") / torch.sqrt(st("1.running_var") + 1e-5)
    wq = (st("0.weight") * s[:, :, None, None]).to(ops.TORCH16[dt]).double()  # [C,P,P,3]
    bq = ((st("0.bias") - st("1.running_mean")) * s + st("1.bias")).double()
    # the packed weights are exactly that rounding, re-laid out as [C,3,P/8,P(out),8(in)]
    lay = wq.permute(0, 3, 2, 1).reshape(Cc, 3, P // 8, 8, P).permute(0, 1, 2, 4, 3)
    assert maxerr(wf, lay) <= 2e-3 * lay.abs().max().item()  # same values up to one rounding step of s
    wq = wf.double().permute(0, 1, 2, 4, 3).reshape(Cc, 3, P, P).permute(0, 3, 2, 1).contiguous()  # kernel's own
    y = torch.stack([torch.nn.functional.conv1d(proj[i].double().transpose(1, 2), wq[i], bq[i], padding=1)
                     for i in range(Cc)])
    y = torch.nn.functional.max_pool1d(y.flatten(0, 1), 3, 2, 1).view(Cc, B, P, T2).transpose(2, 3)
    assert maxerr(out, norm(y.float())) <= 6e-4  # fp16 output rounding (|v| <= 1) + fp32 accumulation order


# ---- stem ---------------------------------------------------------------------------------------
def stem_expect(f16, Tu, sd):
    """conv on the same fp16 features with the fp16-rounded BN-folded weights, fp64 accumulate."""
    w = sd[O.STEM + "convolution.weight"].double()
    gm, b = sd[O.STEM + "normalization.weight"].double(), sd[O.STEM + "normalization.bias"].double()
    m, v = sd[O.STEM + "normalization.running_mean"].double(), sd[O.STEM + "normalization.running_var"].double()
    s = gm / torch.sqrt(v + 1e-5)
    wq = (w * s[:, None, None, None]).float().half().double()
    x = f16[..., :Tu].double().flatten(0, -4)
    y = torch.nn.functional.conv2d(x, wq, (b - m * s), stride=2, padding=3)
    return torch.relu(y).float()


def pack_stem(ops, sd):
    return ops.pack_stem_weights(sd[O.STEM + "convolution.weight"], sd[O.STEM + "normalization.weight"],
                                 sd[O.STEM + "normalization.bias"], sd[O.STEM + "normalization.running_mean"],
                                 sd[O.STEM + "normalization.running_var"])


def pack_stem_fused(ops, sd):
    return ops.pack_stem_fused(sd[O.STEM + "convolution.weight"], sd[O.STEM + "normalization.weight"],
                               sd[O.STEM + "normalization.bias"], sd[O.STEM + "normalization.running_mean"],
                               sd[O.STEM + "normalization.running_var"])


@pytest.mark.parametrize("N,Cc,Tk,Tu", [(1, 3, 8, 40), (2, 3, 22, 70), (2, 12, 23, 301), (1, 16, 150, 1500),
                                        (1, 4, 1, 1), (1, 32, 21, 135), (2, 20, 9, 260)])
def test_stem_matches_conv(ops, cuda_dev, N, Cc, Tk, Tu):
    g = gen(cuda_dev)
    sd = {k: v.to(cuda_dev) for k, v in O.make_weights("L", Cc, 64, seed=9).items()}
    pitch = ops.pitch_for(Tu)
    f16 = (torch.rand(N, Cc, Tk, pitch, generator=g, device=cuda_dev) * 2 - 1).half()
    wp, bias = pack_stem(ops, sd)
    exp = stem_expect(f16, Tu, sd)
    out = ops.stem(f16, Tu, wp, bias, ops.STEM_OUT_NCHW_F32)
    assert out.shape == exp.shape
    assert maxerr(out, exp) <= 2e-4 * max(1.0, Cc / 16)
    o2 = ops.stem(f16, Tu, wp, bias, ops.STEM_OUT_NHWC_BF16)
    assert maxerr(o2, exp) <= 2e-2 * max(1.0, exp.abs().max().item() / 4)


# ---- fused similarity + stem ----------------------------------------------------------------------
@pytest.mark.parametrize("Cc,K,U,Tk,Tu,Dk", [(3, 2, 2, 22, 70, 64), (12, 2, 1, 150, 1500, 64), (4, 1, 2, 150, 300, 384),
                                             (12, 3, 2, 75, 750, 64), (1, 1, 1, 1, 1, 64), (5, 2, 3, 37, 251, 128),
                                             (8, 1, 1, 9, 123, 64), (2, 1, 1, 150, 122, 64)])
def test_sim_stem_fused_matches_unfused_and_conv(ops, cuda_dev, Cc, K, U, Tk, Tu, Dk):
    """kws_sim_stem == kws_stem(kws_sim(.)) up to fp32 summation order, and == conv2d on the fp16 similarity."""
    g = gen(cuda_dev)
    kn = unit_rows(Cc, K, Tk, Dk, g=g, dev=cuda_dev).half()
    un = unit_rows(Cc, U, Tu, Dk, g=g, dev=cuda_dev).half()
    kn[:, 0, Tk // 2:] = 0  # masked (zeroed) keyword frames
    un[:, -1, (2 * Tu) // 3:] = 0
    sd = {k: v.to(cuda_dev) for k, v in O.make_weights("L", Cc, 64, seed=9).items()}
    wp, bias = pack_stem(ops, sd)
    wf, bias_f = pack_stem_fused(ops, sd)
    assert torch.equal(bias, bias_f)
    assert ops.sim_stem_supported(Cc, Tk, Tu, Dk)
    _, f16 = ops.sim(kn, un, False, True)
    exp = stem_expect(f16, Tu, sd)
    ref = ops.stem(f16, Tu, wp, bias, ops.STEM_OUT_NCHW_F32)
    out = ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NCHW_F32)
    assert out.shape == exp.shape == ref.shape
    tol = 3e-4 * max(1.0, exp.abs().max().item())
    assert maxerr(out, exp) <= tol
    assert maxerr(out, ref) <= tol
    o2 = ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NHWC_BF16)
    assert maxerr(o2, exp) <= 2e-2 * max(1.0, exp.abs().max().item() / 4)


def test_sim_stem_fused_diag_and_many_items(ops, cuda_dev):
    """More items than SMs (persistent loop, barrier phases across items) and the DIAG pairing."""
    g = gen(cuda_dev)
    Cc, K, U, Tk, Tu, Dk = 12, 40, 5, 30, 260, 64  # 200 pairs x 3 column tiles = 600 items
    kn = unit_rows(Cc, K, Tk, Dk, g=g, dev=cuda_dev).half()
    un = unit_rows(Cc, U, Tu, Dk, g=g, dev=cuda_dev).half()
    sd = {k: v.to(cuda_dev) for k, v in O.make_weights("L", Cc, 64, seed=3).items()}
    wp, bias = pack_stem(ops, sd)
    wf, _ = pack_stem_fused(ops, sd)
    _, f16 = ops.sim(kn, un, False, True)
    ref = ops.stem(f16, Tu, wp, bias, ops.STEM_OUT_NCHW_F32)
    out = ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NCHW_F32)
    assert maxerr(out, ref) <= 3e-4 * max(1.0, ref.abs().max().item())
    un2 = unit_rows(Cc, K, Tu, Dk, g=g, dev=cuda_dev).half()
    _, f16d = ops.sim(kn, un2, False, True, diag=True)
    refd = ops.stem(f16d, Tu, wp, bias, ops.STEM_OUT_NCHW_F32)
    outd = ops.sim_stem(kn, un2, wf, bias, ops.STEM_OUT_NCHW_F32, diag=True)
    assert outd.shape == refd.shape
    assert maxerr(outd, refd) <= 3e-4 * max(1.0, refd.abs().max().item())


def test_sim_stem_range_is_a_block_of_the_full_job(ops, cuda_dev):
    """kws_sim_stem_range over keywords [k0,k1) x utterances [u0,u1) == that block of the full K x U job,
    bit for bit (same kernel, same operands), written at the front of a larger reused buffer."""
    g = gen(cuda_dev)
    Cc, K, U, Tk, Tu, Dk = 6, 7, 5, 20, 130, 64
    kn = unit_rows(Cc, K, Tk, Dk, g=g, dev=cuda_dev).half()
    un = unit_rows(Cc, U, Tu, Dk, g=g, dev=cuda_dev).half()
    sd = {k: v.to(cuda_dev) for k, v in O.make_weights("L", Cc, 64, seed=5).items()}
    wp, bias = pack_stem_fused(ops, sd)
    full = ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NCHW_F32)  # [K*U,64,Ho,Wo]
    full = full.view(K, U, *full.shape[1:])
    buf = torch.full((K * U * full[0, 0].numel(),), float("nan"), device=cuda_dev)
    for (k0, k1, u0, u1) in [(2, 6, 1, 4), (0, 1, 4, 5), (6, 7, 0, 5)]:
        blk = ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NCHW_F32, out=buf, k_range=(k0, k1), u_range=(u0, u1))
        assert blk.shape[0] == (k1 - k0) * (u1 - u0)
        assert torch.equal(blk.view(k1 - k0, u1 - u0, *blk.shape[1:]), full[k0:k1, u0:u1])
    with pytest.raises(Exception):
        ops.sim_stem(kn, un, wp, bias, ops.STEM_OUT_NCHW_F32, k_range=(5, 9))


@pytest.mark.parametrize("Cc,K,U,Tk,Tu,Dk,modes", [
    (12, 9, 3, 150, 300, 64, ("bf16", "f32")),   # the bench's instance (12 layers, Dk = 64), 16-row chunks
    (3, 7, 2, 37, 130, 64, ("bf16", "f32")),     # generic 16-row instance
    (4, 6, 2, 150, 200, 384, ("bf16", "f32")),   # L variant: 48-row chunks
    (5, 6, 2, 75, 251, 128, ("bf16",)),          # 32-row chunks
    (32, 5, 2, 75, 300, 64, ("bf16",)),          # multi-pass (12 + 12 + 8 layers)
    (16, 40, 5, 30, 260, 64, ("bf16",)),         # multi-pass, more items than SMs
])
def test_sim_stem_ragged_is_bit_identical(ops, cuda_dev, Cc, K, U, Tk, Tu, Dk, modes):
    """kws_sim_stem_ragged (keyword length table: rows beyond a keyword skipped and filled with relu(bias)) ==
    kws_sim_stem_range on the same operands, bit for bit, for every row of a pre-poisoned output buffer; lengths
    include 0 (ghost / empty), 1, odd / even, the full length and lengths above it."""
    g = gen(cuda_dev)
    kn = unit_rows(Cc, K, Tk, Dk, g=g, dev=cuda_dev).half()
    un = unit_rows(Cc, U, Tu, Dk, g=g, dev=cuda_dev).half()
    lens = torch.randint(0, Tk + 1, (K,), generator=g, device=cuda_dev, dtype=torch.int32)
    lens[0], lens[1], lens[2], lens[3] = 0, 1, Tk, min(Tk, 2)
    lens[4] = Tk - 1
    for k in range(K):
        kn[:, k, int(lens[k]):] = 0  # what the folded frame mask guarantees
    lens_arg = lens.clone()
    lens_arg[2] = Tk + 5  # clamped by the kernel
    sd = {k: v.to(cuda_dev) for k, v in O.make_weights("L", Cc, 64, seed=17).items()}
    wf, bias = pack_stem_fused(ops, sd)
    Ho, Wo = (Tk + 1) // 2, (Tu + 1) // 2
    for mode in modes:
        om = ops.STEM_OUT_NHWC_BF16 if mode == "bf16" else ops.STEM_OUT_NCHW_F32
        dt = torch.bfloat16 if mode == "bf16" else torch.float32
        dense = ops.sim_stem(kn, un, wf, bias, om).clone()
        buf = torch.full((K * U * 64 * Ho * Wo,), float("nan"), dtype=dt, device=cuda_dev)
        rag = ops.sim_stem(kn, un, wf, bias, om, out=buf, kwd_len=lens_arg)
        assert not torch.isnan(rag.float()).any(), "ragged kernel left output rows unwritten"
        assert torch.equal(rag, dense)
        # a sub-range of the pair grid (k0 > 0: the length table is indexed by the global keyword id)
        buf.fill_(float("nan"))
        blk = ops.sim_stem(kn, un, wf, bias, om, out=buf, k_range=(1, K - 1), u_range=(U - 1, U), kwd_len=lens_arg)
        exp = dense.view(K, U, *dense.shape[1:])[1:K - 1, U - 1:U].flatten(0, 1)
        assert torch.equal(blk, exp)
    # rows beyond a keyword really are the constant relu(bias) tile
    out = ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NCHW_F32 if "f32" in modes else ops.STEM_OUT_NHWC_BF16, kwd_len=lens_arg)
    o = out.view(K, U, *out.shape[1:]).float()
    rb = torch.relu(bias)
    if "f32" not in modes:
        rb = rb.to(torch.bfloat16).float()
    assert torch.equal(o[0], rb[None, :, None, None].expand_as(o[0]))  # length 0: the whole image


# ---- max-pool that follows the stem -----------------------------------------------------------------
@pytest.mark.parametrize("N,Cc,H,W", [(1, 64, 1, 1), (2, 64, 7, 9), (3, 64, 38, 375), (2, 64, 75, 750), (2, 8, 12, 5)])
def test_maxpool_nhwc_is_bit_identical_to_torch(ops, cuda_dev, N, Cc, H, W):
    """MaxPool2d(3,2,1) (HF ResNetEmbeddings.pooler) on the channels-last bf16 stem activation: exact."""
    g = gen(cuda_dev)
    x = torch.randn(N, H, W, Cc, generator=g, device=cuda_dev).to(torch.bfloat16).permute(0, 3, 1, 2)
    got = ops.maxpool_nhwc(x)
    exp = torch.nn.functional.max_pool2d(x.float(), 3, 2, 1)
    assert got.shape == exp.shape and got.dtype == torch.bfloat16
    assert got.permute(0, 2, 3, 1).is_contiguous()
    assert torch.equal(got.float(), exp)
    if H * W > 1:
        with pytest.raises(Exception):  # NCHW-contiguous input is refused, not silently re-laid out
            ops.maxpool_nhwc(x.contiguous())


# ---- scores + top-k -------------------------------------------------------------------------------
def test_scores_and_detections(ops, cuda_dev):
    g = gen(cuda_dev)
    logits = torch.randn(1000, 2, generator=g, device=cuda_dev) * 3
    hw = (torch.rand(1000, generator=g, device=cuda_dev) > 0.05).float()
    sc, det = ops.scores(logits, hw, 0.5)
    exp = logits.softmax(-1)[:, 1] * hw
    assert maxerr(sc, exp) <= 1e-6
    safe = (exp - 0.5).abs() > 1e-5
    assert torch.equal(det.bool()[safe], (exp >= 0.5)[safe])


@pytest.mark.parametrize("n,U,k", [(50, 3, 10), (1000, 4, 200), (7, 2, 7), (300, 1, 1), (2048, 5, 1024), (2049, 3, 10),
                                   (5000, 6, 200), (100000, 9, 200), (30000, 2, 1024)])
def test_topk_matches_torch_with_index_tiebreak(ops, cuda_dev, n, U, k):
    """Segmented selection (one level per factor-of-(2048/k) reduction) == a stable descending sort: exact values,
    exact ids, ties broken by the lower id -- also across segment boundaries."""
    g = gen(cuda_dev)
    sc = torch.randn(n, U, generator=g, device=cuda_dev)
    sc[n // 2] = sc[0]  # exact ties -> lower id first
    if n > 4096:
        sc[4097] = sc[1] = sc[:, :1].max() + 1.0  # the top value twice, in different segments
        sc[n - 1, :] = float("-inf")
    v, i = ops.topk(sc.contiguous(), k, None, 100)
    order = torch.argsort(-sc.double() - 0.0, dim=0, stable=True)[:k]  # stable: lower index first on ties
    assert torch.equal(v, torch.gather(sc, 0, order))
    assert torch.equal(i.long(), order + 100)
    # merge form: explicit (shuffled) ids, fewer candidates than k -> (-inf, -1) rows
    if n <= 1000:
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(n)).to(cuda_dev)
        ids = perm.to(torch.int32)[:, None].expand(n, U).contiguous()
        v2, i2 = ops.topk(sc.contiguous(), min(1024, n + 3), ids, 0)
        order2 = torch.stack([torch.tensor(sorted(range(n), key=lambda c: (-float(sc[c, u]), int(perm[c]))), device=cuda_dev)
                              for u in range(U)], dim=1)
        kk = min(1024, n + 3)
        assert torch.equal(v2[:n], torch.gather(sc, 0, order2)[:kk]) and torch.equal(i2[:n].long(), perm[order2][:kk])
        if kk > n:
            assert torch.isinf(v2[n:]).all() and (v2[n:] < 0).all() and (i2[n:] == -1).all()


# ---- config #4: similarity + bilinear resize (original CB-Whisper classifier) ---------------------
@pytest.mark.parametrize("K,U,Cc,Hs,Ws,size", [(3, 2, 2, 16, 40, (9, 20)), (2, 1, 3, 32, 150, (150, 75)),
                                               (4, 1, 1, 16, 31, (7, 50))])
def test_resize_bilinear_matches_interpolate(ops, cuda_dev, K, U, Cc, Hs, Ws, size):
    g = gen(cuda_dev)
    x = torch.randn(K, U, Cc, Hs, Ws, generator=g, device=cuda_dev)
    h = torch.randint(1, Hs + 1, (K,), generator=g, device=cuda_dev, dtype=torch.int32)
    o32, o16 = ops.resize_bilinear(x, h, size, want_f32=True, want_f16=True)
    for k in range(K):
        exp = torch.nn.functional.interpolate(x[k, :, :, : int(h[k])].flatten(0, 1)[None], size=size, mode="bilinear",
                                              align_corners=False, antialias=False)[0].view(U, Cc, *size)
        assert maxerr(o32[k], exp) <= 1e-5
        assert maxerr(o16[k][..., : size[1]], exp) <= 2e-3 * max(1.0, exp.abs().max().item())
    # no length table: every source row is valid
    o32b, _ = ops.resize_bilinear(x, None, size, want_f32=True, want_f16=False)
    expb = torch.nn.functional.interpolate(x.flatten(0, 2)[None], size=size, mode="bilinear", align_corners=False)[0]
    assert maxerr(o32b.flatten(0, 2), expb) <= 1e-5


@pytest.mark.parametrize("B,Cin,T,D,T_out", [(2, 3, 40, 64, 20), (1, 2, 1500, 1024, 750), (2, 2, 31, 128, 50), (1, 1, 7, 1280, 7)])
def test_interp_rows_matches_interpolate(ops, cuda_dev, B, Cin, T, D, T_out):
    """Width map of the config-#4 resize applied to the (normalised) utterance frames."""
    g = gen(cuda_dev)
    x = torch.randn(B, Cin, T, D, generator=g, device=cuda_dev) * 3
    idx = list(range(Cin))[::-1]
    got = ops.interp_rows(x, idx, T_out)
    xn = torch.nn.functional.normalize(x[:, idx], dim=-1)  # [B,C,T,D]
    exp = torch.nn.functional.interpolate(xn.flatten(0, 1).transpose(1, 2), size=T_out, mode="linear",
                                          align_corners=False).transpose(1, 2).view(B, Cin, T_out, D).transpose(0, 1)
    assert got.shape == (Cin, B, T_out, D) and got.dtype == torch.float16
    assert maxerr(got, exp) <= 6e-4


def test_resize_row_weights_are_the_bilinear_taps(ops, cuda_dev):
    lens = torch.tensor([1, 5, 17, 33, 64, 10], dtype=torch.int32, device=cuda_dev)
    K, Cc, Hp, Ho = lens.numel(), 2, 64, 150
    wy = ops.resize_row_weights(lens, K, Cc, Hp, Ho)
    assert wy.shape == (Cc, K, Ho, Hp)
    for k, h in enumerate(lens.tolist()):
        eye = torch.eye(h, device=cuda_dev)[None, None]  # resize of the identity = the map itself
        exp = torch.nn.functional.interpolate(eye, size=(Ho, h), mode="bilinear", align_corners=False)[0, 0]
        for c in range(Cc):
            assert maxerr(wy[c, k, :, :h], exp) <= 5e-4
            assert float(wy[c, k, :, h:].abs().max()) == 0.0 if h < Hp else True


def test_sim_operand_is_the_transposed_similarity(ops, cuda_dev):
    g = gen(cuda_dev)
    Cc, K, U, Tk, Tu, Dk = 3, 4, 2, 64, 200, 128
    kn = unit_rows(Cc, K, Tk, Dk, g=g, dev=cuda_dev).half()
    kn[:, 1, 20:] = 0
    un = unit_rows(Cc, U, Tu, Dk, g=g, dev=cuda_dev).half()
    got = ops.sim_operand(kn, un)  # [C, K*U, Tu, Tk]
    exp = torch.einsum("cutd,ckid->ckuti", un.float(), kn.float()).reshape(Cc, K * U, Tu, Tk)
    assert got.shape == exp.shape
    assert maxerr(got, exp) <= 1e-3
    assert float(got.view(Cc, K, U, Tu, Tk)[:, 1, :, :, 20:].abs().max()) == 0.0


@pytest.mark.parametrize("size,Tu,lens", [((30, 50), 100, (9, 21, 14, 64, 1)), ((150, 750), 1500, (10, 60))])
def test_cbw_fused_stem_never_builds_the_image(built_lib, cuda_dev, size, Tu, lens):
    """Config #4 scoring path: (Wy kwd) . (Wx utt)^T inside the fused kernel == stem of the restated
    cb_whisper.py:189-210 images (bf16 output; operand-side resize adds two fp16 roundings)."""
    from enhance_cb_whisper_b200 import Resnet, cbw, ops as _ops

    torch.manual_seed(5)
    g = torch.Generator().manual_seed(13)
    Cc, D, S = 12, 64, 2
    resnet = Resnet(Cc, 2, "resnet-18").eval()
    bn = resnet.feature_extractor.embedder.embedder.normalization
    with torch.no_grad():
        bn.weight.copy_(torch.rand(64, generator=g) + 0.5), bn.bias.copy_(torch.randn(64, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(64, generator=g) * 0.1), bn.running_var.copy_(torch.rand(64, generator=g) + 0.5)
    kwd_list = [torch.nn.functional.normalize(torch.randn(Cc, t, D, generator=g), dim=-1) for t in lens]
    utt = torch.nn.functional.normalize(torch.randn(S, Cc, Tu, D, generator=g), dim=-1)
    with torch.inference_mode():
        imgs = O.cbw_similarity_resized(kwd_list, utt, size=size)  # [K,S,C,h,w]
        exp = resnet.feature_extractor.embedder.embedder(imgs.flatten(0, 1))  # conv + BN + ReLU
    sp = cbw.CBWKeywordSpotterB200(resnet.to(cuda_dev), size=size)
    kwd_n, lens_t = cbw.pack_keywords([k.to(cuda_dev) for k in kwd_list], cuda_dev, multiple=64)
    utt_i = _ops.interp_rows(utt.to(cuda_dev), list(range(Cc)), size[1])
    got = []
    sp.stem_fused(kwd_n, lens_t, utt_i, _ops.STEM_OUT_NCHW_F32, max_pairs=4, consume=lambda k0, k1, st: got.append(st.clone()))
    got = torch.cat(got)
    assert got.shape == exp.shape
    assert maxerr(got.cpu(), exp) <= 3e-3 * max(1.0, exp.abs().max().item())
    got16 = []
    sp.stem_fused(kwd_n, lens_t, utt_i, _ops.STEM_OUT_NHWC_BF16, max_pairs=64, consume=lambda k0, k1, st: got16.append(st.float()))
    assert maxerr(torch.cat(got16).cpu(), exp) <= 2e-2 * max(1.0, exp.abs().max().item() / 4)


def test_cbw_similarity_images_match_oracle(built_lib, cuda_dev):
    """B200 config-#4 path == restated cb_whisper.py:189-210 (ragged keywords, matmul, bilinear resize)."""
    from enhance_cb_whisper_b200 import cbw

    g = torch.Generator().manual_seed(11)
    Cc, D, Tu, S = 3, 128, 140, 2
    kwd_list = [torch.nn.functional.normalize(torch.randn(Cc, t, D, generator=g), dim=-1) for t in (5, 17, 33, 1)]
    utt = torch.nn.functional.normalize(torch.randn(S, Cc, Tu, D, generator=g), dim=-1)
    exp = O.cbw_similarity_resized(kwd_list, utt, size=(30, 70))  # [K,S,C,30,70]
    got32, got16 = cbw.similarity_images([k.to(cuda_dev) for k in kwd_list], utt.to(cuda_dev), size=(30, 70),
                                         want_f32=True, want_f16=True)
    assert got32.shape == exp.shape
    assert maxerr(got32.cpu(), exp) <= 2e-3
    assert maxerr(got16[..., :70].cpu(), exp) <= 2e-3


def test_cbw_keyword_spotter_logits_and_detections(built_lib, cuda_dev):
    """12-channel classifier on the resized images: logits within 2e-3 of the reference arithmetic
    (src/model/model.py:78-93 on the images of cb_whisper.py:189-210), same argmax detections."""
    from enhance_cb_whisper_b200 import Resnet, cbw

    torch.manual_seed(5)
    g = torch.Generator().manual_seed(12)
    Cc, D, Tu, S, size = 12, 64, 100, 2, (30, 50)
    resnet = Resnet(Cc, 2, "resnet-18").eval()
    bn = resnet.feature_extractor.embedder.embedder.normalization
    with torch.no_grad():
        bn.weight.copy_(torch.rand(64, generator=g) + 0.5), bn.bias.copy_(torch.randn(64, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(64, generator=g) * 0.1), bn.running_var.copy_(torch.rand(64, generator=g) + 0.5)
    kwd_list = [torch.nn.functional.normalize(torch.randn(Cc, t, D, generator=g), dim=-1) for t in (9, 21, 14)]
    utt = torch.nn.functional.normalize(torch.randn(S, Cc, Tu, D, generator=g), dim=-1)
    with torch.inference_mode():
        imgs = O.cbw_similarity_resized(kwd_list, utt, size=size)  # [K,S,C,h,w]
        exp = resnet(imgs.flatten(0, 1)).view(len(kwd_list), S, 2)
    sp = cbw.CBWKeywordSpotterB200(resnet.to(cuda_dev), size=size)
    got = sp.logits([k.to(cuda_dev) for k in kwd_list], utt.to(cuda_dev), fused=False)  # images + un-fused stem
    assert maxerr(got.cpu(), exp) <= 2e-3
    assert sp.fused_ok(kwd_list, utt.to(cuda_dev))
    got = sp.logits([k.to(cuda_dev) for k in kwd_list], utt.to(cuda_dev))  # fused: the image is never built
    assert maxerr(got.cpu(), exp) <= 4e-3
    det = sp.detect([k.to(cuda_dev) for k in kwd_list], utt.to(cuda_dev))
    exp_hit = exp.argmax(-1) == 1
    margin = (exp[..., 1] - exp[..., 0]).abs().min().item()
    if margin > 1e-2:
        assert det == [torch.nonzero(exp_hit[:, s]).flatten().tolist() for s in range(S)]
    # bf16 body: the stem's max-pool inside the fused kernel (KWS_PAIRS_PER_KEYWORD + POOL) == stem -> kws_maxpool_nhwc
    sp16 = cbw.CBWKeywordSpotterB200(resnet.to(cuda_dev), size=size, body_dtype="bfloat16")
    a16 = sp16.logits([k.to(cuda_dev) for k in kwd_list], utt.to(cuda_dev))
    sp16.fused_pool = False
    b16 = sp16.logits([k.to(cuda_dev) for k in kwd_list], utt.to(cuda_dev))
    assert torch.equal(a16, b16)
    assert maxerr(a16.cpu(), exp) <= 0.1 * max(1.0, exp.abs().max().item())


def test_cbw_matches_reference_fixture(built_lib, cuda_dev):
    """Config #4 against the committed outputs of the UNMODIFIED reference pieces (tests/golden/cbw_small.npz:
    CBWhisper._calculate_cosine_similarity_matrices_, cb_whisper.py:189-210, and the 12-channel classifier
    src/model/resnet.py): images and logits within 2e-3 (un-fused path), fused path within 4e-3 (two more fp16
    roundings from the operand-side resize), same argmax detections, keyword_spotting-shaped entry."""
    from enhance_cb_whisper_b200 import cbw
    from oracle.make_golden import CBW_CASE, load_cbw_case

    kwd_list, utt, outs, net, same = load_cbw_case()
    assert same, "regenerated classifier differs from the one the fixture was made with"
    size = CBW_CASE["size"]
    kd, ud = [k.to(cuda_dev) for k in kwd_list], utt.to(cuda_dev)
    got32, got16 = cbw.similarity_images(kd, ud, size=size, want_f32=True, want_f16=True)
    assert maxerr(got32.cpu(), outs["images"]) <= 2e-3
    assert maxerr(got16[..., : size[1]].cpu(), outs["images"]) <= 2e-3
    sp = cbw.CBWKeywordSpotterB200(net.to(cuda_dev), size=size)
    lg_unfused = sp.logits(kd, ud, fused=False)
    assert maxerr(lg_unfused.cpu(), outs["logits"]) <= 2e-3
    assert sp.fused_ok(kwd_list, ud)
    lg_fused = sp.logits(kd, ud)
    assert maxerr(lg_fused.cpu(), outs["logits"]) <= 4e-3
    exp_hit = outs["logits"].argmax(-1) == 1  # cb_whisper.py:128
    clear = (outs["logits"][..., 1] - outs["logits"][..., 0]).abs() > 1e-2
    for lg in (lg_unfused, lg_fused):
        assert torch.equal((lg.cpu().argmax(-1) == 1)[clear], exp_hit[clear])
    # inputs that are NOT unit vectors: the reference does a plain matmul, so must this path (no re-normalisation)
    kd2 = [k * 0.5 for k in kd]
    h32, _ = cbw.similarity_images(kd2, ud, size=size, want_f32=True, want_f16=False)
    assert maxerr(h32.cpu(), outs["images"] * 0.5) <= 1e-3
    # keyword_spotting-shaped call (cb_whisper.py:110-131): groups of the keyword database -> retrieved keywords per segment
    names = [f"kw{i}" for i in range(len(kd))]
    groups = [{"hidden_states": kd[:3], "keywords": names[:3]}, {"hidden_states": kd[3:], "keywords": names[3:]},
              {"hidden_states": [], "keywords": []}]
    got = sp.keyword_spotting(ud, groups)
    hit = lg_fused.cpu().argmax(-1) == 1
    assert [sorted(g) for g in got] == [sorted(names[i] for i in range(len(kd)) if hit[i, s]) for s in range(utt.shape[0])]
    assert sp.keyword_spotting(None, groups, n_segments=2) == [[], []]  # failed feature extraction (cb_whisper.py:113-116)


@pytest.mark.parametrize("multi", [(12, 2, 2), (12, 2, 0), (8, 1, 1), (8, 2, 1), (12, 1, 2)])
@pytest.mark.parametrize("Cc,K,U,Tk,Tu", [(32, 2, 2, 75, 300), (16, 1, 2, 22, 130), (25, 2, 1, 9, 61), (13, 3, 50, 9, 130)])
def test_sim_stem_fused_channel_groups(ops, cuda_dev, Cc, K, U, Tk, Tu, multi):
    """More than 12 layers: one fused pass per group of layers, fp16 partial sums chained through the output
    buffer, bf16 channels-last result == conv2d on the fp16 similarity images (cfg3/cfg5 shape: C = 32).
    ``multi`` = (layers per pass, channels/16 per epilogue round trip, partial-sum prefetch: 0 none | 1 second
    staging set | 2 L2 only); (12, 2, 2) is the shipped default, the others are the development variants."""
    from enhance_cb_whisper_b200 import _lib

    lib = _lib.load()
    hooks = hasattr(lib, "kws_debug_set_fused_multi")  # only the -DKWS_DEBUG_HOOKS flavour (KWS_B200_LIB=...dbg.so)
    if not hooks and multi != (12, 2, 2):
        pytest.skip("development variant: needs the debug-hooks flavour of the library")
    if hooks:
        lib.kws_debug_set_fused_multi(*multi)
        lib.kws_debug_set_fused_reduce(0 if multi == (12, 2, 0) else 1)  # middle passes: TMA reduce-add | load+add
    try:
        g = gen(cuda_dev)
        kn = unit_rows(Cc, K, Tk, 64, g=g, dev=cuda_dev).half()
        un = unit_rows(Cc, U, Tu, 64, g=g, dev=cuda_dev).half()
        kn[:, 0, Tk // 2:] = 0
        sd = {k: v.to(cuda_dev) for k, v in O.make_weights("L", Cc, 64, seed=21).items()}
        wf, bias = pack_stem_fused(ops, sd)  # packed per group: after the variant is chosen
        assert ops.sim_stem_supported(Cc, Tk, Tu, 64, ops.STEM_OUT_NHWC_BF16)
        assert not ops.sim_stem_supported(Cc, Tk, Tu, 64, ops.STEM_OUT_NCHW_F32)
        _, f16 = ops.sim(kn, un, False, True)
        exp = stem_expect(f16, Tu, sd)
        out = ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NHWC_BF16)
        assert out.shape == exp.shape
        assert maxerr(out, exp) <= 2e-2 * max(1.0, exp.abs().max().item() / 4)
        with pytest.raises(Exception):
            ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NCHW_F32)
    finally:
        if hooks:
            lib.kws_debug_set_fused_multi(12, 2, 2)
            lib.kws_debug_set_fused_reduce(1)


# ---- fused similarity + stem + max-pool ------------------------------------------------
@pytest.mark.parametrize("Cc,K,U,Tk,Tu,Dk,ragged", [
    (12, 3, 2, 150, 1500, 64, False),  # the bench's instance (S12), 13 column tiles of 29 pooled columns
    (12, 9, 3, 150, 300, 64, True),    # S12 + keyword length table
    (3, 2, 2, 22, 70, 64, False),      # generic 16-row instance, one column tile
    (3, 7, 2, 37, 130, 64, True),      # odd Ho (19), odd Wo (65), ragged
    (4, 2, 2, 150, 300, 384, False),   # L variant: 48-row chunks
    (4, 6, 2, 150, 200, 384, True),
    (5, 2, 3, 75, 251, 128, False),    # 32-row chunks, Wo = 126 -> Wp = 63 = 2 full tiles + 5 columns
    (5, 6, 2, 75, 251, 128, True),
    (1, 1, 1, 1, 1, 64, False),        # 1 x 1 image
    (2, 1, 1, 150, 122, 64, False),
    (8, 1, 1, 9, 123, 64, False),
    (32, 3, 2, 75, 750, 64, False),    # multi-pass (12 + 12 + 8 layers): partial sums in the workspace, last pass pools
    (32, 5, 2, 75, 300, 64, True),
    (16, 40, 5, 30, 260, 64, True),    # multi-pass, more items than SMs
    (12, 40, 5, 30, 260, 64, False),   # more items than SMs (persistent loop, carried rows reset per item)
])
def test_sim_stem_pool_is_maxpool_of_the_fused_stem(ops, cuda_dev, Cc, K, U, Tk, Tu, Dk, ragged):
    """kws_sim_stem_pool == kws_maxpool_nhwc(kws_sim_stem_ragged(..., bf16 channels-last)), bit for bit (the maximum
    commutes with the monotone bf16 rounding and both run the same MMAs), == F.max_pool2d(3, 2, 1) of it."""
    g = gen(cuda_dev)
    kn = unit_rows(Cc, K, Tk, Dk, g=g, dev=cuda_dev).half()
    un = unit_rows(Cc, U, Tu, Dk, g=g, dev=cuda_dev).half()
    un[:, -1, (2 * Tu) // 3:] = 0
    lens = None
    if ragged:
        lens = torch.randint(0, Tk + 1, (K,), generator=g, device=cuda_dev, dtype=torch.int32)
        lens[0] = 0
        lens[1] = 1
        if K > 2:
            lens[2] = Tk
        if K > 3:
            lens[3] = min(Tk, 6)
        for k in range(K):
            kn[:, k, int(lens[k]):] = 0
    sd = {k: v.to(cuda_dev) for k, v in O.make_weights("L", Cc, 64, seed=23).items()}
    wf, bias = pack_stem_fused(ops, sd)
    full = ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NHWC_BF16, kwd_len=lens)
    exp = ops.maxpool_nhwc(full)
    assert torch.equal(exp.float(), torch.nn.functional.max_pool2d(full.float(), 3, 2, 1))
    Hp, Wp = exp.shape[2], exp.shape[3]
    buf = torch.full((K * U * 64 * Hp * Wp + 64,), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    got = ops.sim_stem_pool(kn, un, wf, bias, out=buf, kwd_len=lens)
    assert got.shape == exp.shape
    assert not torch.isnan(got.float()).any(), "pooled rows left unwritten"
    assert torch.isnan(buf[K * U * 64 * Hp * Wp:].float()).all(), "wrote beyond the pooled activation"
    assert torch.equal(got, exp)
    if K >= 3:  # a sub-range of the pair grid into the front of a reused buffer
        buf.fill_(float("nan"))
        blk = ops.sim_stem_pool(kn, un, wf, bias, out=buf, k_range=(1, K), u_range=(U - 1, U), kwd_len=lens)
        assert torch.equal(blk, exp.view(K, U, *exp.shape[1:])[1:K, U - 1:U].flatten(0, 1))
