"""GPU: parity at the BASELINE.json shapes, and tight checks of the kernel instances the benchmark runs.

(1) Module vs the CPU oracle at the cfg1 (L, 4 x 384-d), cfg2 (LE, 12 x 768-d) and cfg3 (LEF, 32 x 1280-d) shapes with
    K = 2 keywords x U = 1 utterance of 150 x 1500 frames: compressed operands, similarity features and stem activation
    within the north-star's 2e-3 absolute (the golden fixtures are tiny; these are the shapes the bench runs).
(2) The 12-layer specialised instance (kws_fused_kernel<1,16,0,2,1>) and the multi-pass instances (<1,16,1,2,*>) that the
    bench times write bf16: their output must be the bf16 ROUNDING of what the fp32 instances compute, to within one
    bf16 unit in the last place (plus, for the multi-pass chain, the fp16 rounding of the partial sums it stores).
"""
import pytest
import torch

from oracle import kws_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-3


@pytest.fixture(scope="module")
def ops(built_lib, cuda_dev):
    from enhance_cb_whisper_b200 import ops as _ops

    return _ops


SHAPES = {  # name: variant, C, D, P, Tk, Tu
    "cfg1": ("L", 4, 384, 64, 150, 1500),
    "cfg2": ("LE", 12, 768, 64, 150, 1500),
    "cfg3": ("LEF", 32, 1280, 64, 150, 1500),
}


def _model(variant, C, D, P, Tk, Tu, dev, sd):
    import enhance_cb_whisper_b200 as kb

    m = kb.KWSModelB200(n_layers=C, embedding_dim=D, proj_mlp_units=P, learn_features=variant != "L",
                        proj_mlp=variant != "L", frames_conv=variant == "LEF", resnet_version="resnet-18",
                        features_size=(Tk, Tu))
    full = dict(m.state_dict())
    full.update(sd)
    m.load_state_dict(full)
    return m.to(dev).eval()


def _bf16_ulp(x):
    """one unit in the last place of bf16 at |x| (8 significand bits), floor at the smallest normal fp16 step"""
    return torch.clamp(x.abs(), min=2.0 ** -14) * 2.0 ** -7


@pytest.mark.parametrize("name", list(SHAPES))
def test_module_matches_oracle_at_baseline_shapes(built_lib, cuda_dev, name):
    from enhance_cb_whisper_b200 import ops

    variant, C, D, P, Tk, Tu = SHAPES[name]
    K, U = 2, 1
    sd = O.make_weights(variant, C, D, P, seed=41)
    kwd, utt, km, um, _ = O.make_inputs(K, U, C, D, Tk, Tu, seed=42, ghost_frac=0.0, min_k=40)
    if variant == "LEF":
        km, um = O.pooled_mask(km).contiguous(), O.pooled_mask(um).contiguous()
    with torch.inference_mode():
        exp = O.forward_pairs(kwd, utt, km, um, sd, variant, upto="stem")
    m = _model(variant, C, D, P, Tk, Tu, cuda_dev, sd)
    eng = m.prepare(cuda_dev)
    kn = eng.compress(kwd.to(cuda_dev), km.to(cuda_dev))
    un = eng.compress(utt.to(cuda_dev), um.to(cuda_dev))
    # compressed operands: the oracle's rows, L2-normalised and masked (what the similarity contracts)
    for got, raw, mask in ((kn, exp["kwd_c"], km), (un, exp["utt_c"], um)):
        ref = (raw / raw.norm(dim=-1, keepdim=True).clamp_min(O.SIM_EPS) * mask[..., None]).transpose(0, 1)  # [C,B,T',Dk]
        assert got.shape == ref.shape
        assert (got.float().cpu() - ref).abs().max().item() <= TOL
    # similarity features [K,U,C,Tk',Tu'] (un-fused kernel, fp32 output = KWSOutput.features)
    r = m(kwd_features=kwd.to(cuda_dev), utt_features=utt.to(cuda_dev), kwd_mask=km.to(cuda_dev), utt_mask=um.to(cuda_dev))
    assert r.features.shape == exp["features"][:, 0].shape
    assert (r.features.cpu() - exp["features"][:, 0]).abs().max().item() <= TOL
    # stem activation, the in-scope end point: fp32 parity output where the fused kernel offers it (C <= 12), and the
    # bf16 channels-last output the bench uses
    e = exp["stem"].flatten(0, 1)
    scale = max(1.0, e.abs().max().item())
    got = {}
    if C <= 12:
        eng.hot_path(kn, un, ops.STEM_OUT_NCHW_F32, 4, lambda *a: got.__setitem__("f32", a[-1].float().cpu()))
        assert got["f32"].shape == e.shape
        assert (got["f32"] - e).abs().max().item() <= TOL * scale
    eng.hot_path(kn, un, ops.STEM_OUT_NHWC_BF16, 4, lambda *a: got.__setitem__("bf16", a[-1].float().cpu()))
    assert got["bf16"].shape == e.shape
    err = (got["bf16"] - e).abs()
    assert bool((err <= TOL * scale + _bf16_ulp(e)).all()), f"bf16 stem: max err {err.max().item():.3e}"


def _fused_inputs(Cc, K, U, Tk, Tu, Dk, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    unit = lambda *s: torch.nn.functional.normalize(torch.randn(*s, generator=g, device=dev), dim=-1).half()
    kn, un = unit(Cc, K, Tk, Dk), unit(Cc, U, Tu, Dk)
    kn[:, 0, (2 * Tk) // 3:] = 0
    un[:, -1, (3 * Tu) // 4:] = 0
    sd = {k: v.to(dev) for k, v in O.make_weights("L", Cc, 64, seed=seed + 1).items()}
    return kn, un, sd


def _pack(ops, sd, fused):
    f = ops.pack_stem_fused if fused else ops.pack_stem_weights
    return f(sd[O.STEM + "convolution.weight"], sd[O.STEM + "normalization.weight"], sd[O.STEM + "normalization.bias"],
             sd[O.STEM + "normalization.running_mean"], sd[O.STEM + "normalization.running_var"])


@pytest.mark.parametrize("K,U,Tk,Tu", [(3, 2, 150, 1500), (5, 3, 75, 750), (40, 5, 30, 260)])
def test_s12_instance_is_the_bf16_rounding_of_the_fp32_instance(ops, cuda_dev, K, U, Tk, Tu):
    """kws_fused_kernel<1,16,0,2,1> (12 layers, Dk = 64, bf16 channels-last: the bench's headline instance) against
    kws_fused_kernel<0,16,0> (fp32 NCHW, the instance the 3e-4 tests cover): same MMAs in the same order, so the bf16
    output is the rounding of the fp32 output -- at most one bf16 ulp apart (rounding boundary cases of the packed
    convert), and equal almost everywhere."""
    kn, un, sd = _fused_inputs(12, K, U, Tk, Tu, 64, cuda_dev, seed=51)
    wf, bias = _pack(ops, sd, fused=True)
    f32 = ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NCHW_F32)
    b16 = ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NHWC_BF16)
    assert b16.dtype == torch.bfloat16 and b16.shape == f32.shape
    d = (b16.float() - f32).abs()
    assert bool((d <= _bf16_ulp(f32)).all()), f"max |bf16 - fp32| = {d.max().item():.3e}"
    same = (b16 == f32.to(torch.bfloat16)).float().mean().item()
    assert same >= 0.999, f"only {same:.4f} of the outputs equal the round-to-nearest bf16 of the fp32 instance"


@pytest.mark.parametrize("Cc,K,U,Tk,Tu", [(32, 3, 2, 75, 750), (24, 2, 2, 75, 300), (32, 40, 4, 30, 260), (20, 2, 1, 150, 1500)])
def test_multipass_instances_match_the_fp32_stem(ops, cuda_dev, Cc, K, U, Tk, Tu):
    """C > 12 layers (cfg3 / cfg5): kws_fused_kernel<1,16,1,2,1> for the full 12-layer groups and <1,16,1,2,0> for the
    remainder chain fp16 partial sums through the output tiles.  Reference: the un-fused kws_stem (fp32 accumulation
    over all layers, fp32 workspace between its 16-layer groups) on the fp16 similarity of the same operands.  Bound per
    element: one bf16 ulp of the result + the fp16 rounding of each stored partial sum (2^-11 relative, one per group
    boundary, taken at the magnitude of the running sums)."""
    kn, un, sd = _fused_inputs(Cc, K, U, Tk, Tu, 64, cuda_dev, seed=61)
    wf, bias = _pack(ops, sd, fused=True)
    wp, bias_u = _pack(ops, sd, fused=False)
    assert torch.equal(bias, bias_u)
    _, f16 = ops.sim(kn, un, False, True)
    ref = ops.stem(f16, Tu, wp, bias, ops.STEM_OUT_NCHW_F32)
    got = ops.sim_stem(kn, un, wf, bias, ops.STEM_OUT_NHWC_BF16).float()
    assert got.shape == ref.shape
    # magnitude of the running partial sums after each 12-layer group (conv of the fp16 similarity, fp32, torch)
    w = sd[O.STEM + "convolution.weight"]
    s = sd[O.STEM + "normalization.weight"] / torch.sqrt(sd[O.STEM + "normalization.running_var"] + 1e-5)
    wq = (w * s[:, None, None, None]).half().float()
    x = f16[..., :Tu].float().flatten(0, 1)
    run = torch.zeros_like(ref)
    slack = torch.zeros_like(ref)
    groups = [(c0, min(c0 + 12, Cc)) for c0 in range(0, Cc, 12)]
    for c0, c1 in groups[:-1]:
        run = run + torch.nn.functional.conv2d(x[:, c0:c1], wq[:, c0:c1], None, stride=2, padding=3)
        slack = slack + run.abs() * 2.0 ** -11 + 2.0 ** -24
    d = (got - ref).abs()
    bound = _bf16_ulp(ref) + slack + 3e-4 * max(1.0, ref.abs().max().item())  # + fp32 summation-order term of the 3e-4 tests
    assert bool((d <= bound).all()), f"max |err| {d.max().item():.3e}, worst excess {(d - bound).max().item():.3e}"


@pytest.mark.parametrize("variant", ["L", "LE", "LEF"])
def test_streamed_bank_scores_match_the_oracle(built_lib, cuda_dev, variant):
    """build_keyword_bank (ragged *.bin-style items incl. a ghost) + score_bank against the CPU oracle on the padded
    batch (dataset.py:784-819 padding, model.py:783-795 scores): logits within 2e-3, ghost scores exactly 0."""
    import enhance_cb_whisper_b200 as kb
    from enhance_cb_whisper_b200 import bank
    from oracle.make_golden import build_body

    C, D, P, Tk, Tu = 3, 128, 64, 22, 70
    sd = O.make_weights(variant, C, D, P, seed=71)
    rv = "resnet-50" if variant == "L" else "resnet-18"  # model.py:74-76: the L variant ignores resnet_version
    fe, head = build_body(C, rv, 5)
    m = kb.KWSModelB200(n_layers=C, embedding_dim=D, proj_mlp_units=P, learn_features=variant != "L",
                        proj_mlp=variant != "L", frames_conv=variant == "LEF", resnet_version=rv,
                        features_size=(Tk, Tu))
    full = dict(m.state_dict())
    full.update({"model.feature_extractor." + k: t for k, t in fe.state_dict().items()})
    full.update({"model.classifier." + k: t for k, t in head.state_dict().items()})
    full.update(sd)
    m.load_state_dict(full)
    m = m.to(cuda_dev).eval()
    g = torch.Generator().manual_seed(4)
    lens = [5, 22, 30, 9, 1, 17, 12]
    items = [torch.nn.functional.normalize(torch.randn(12, t, D, generator=g), dim=-1) for t in lens]
    items[3] = None
    b = bank.build_keyword_bank(m, items, Tk, cuda_dev, chunk=3)
    padded = torch.stack([bank.pad_item(t if t is not None else torch.zeros(12, 1, D), Tk, C)[0] for t in items])
    valid = torch.tensor([0 if t is None else min(t.shape[1], Tk) for t in items])
    km = (torch.arange(Tk)[None] < valid[:, None]).float()[:, None, :].expand(-1, C, -1).contiguous()
    utt = torch.nn.functional.normalize(torch.randn(2, C, Tu, D, generator=g), dim=-1)
    um = torch.ones(2, C, Tu)
    if variant == "LEF":
        km, um = O.pooled_mask(km).contiguous(), O.pooled_mask(um).contiguous()
    body = lambda x: fe.pooler(fe.encoder(x).last_hidden_state).flatten(1)
    with torch.inference_mode():
        exp = O.forward_pairs(padded, utt, km, um, sd, variant, body=body, classifier=head)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sc, det, lg = bank.score_bank(m, b, utt.to(cuda_dev), um.to(cuda_dev))
    assert (lg.cpu() - exp["logits"]).abs().max().item() <= TOL
    hot = torch.tensor([0.0 if t is None else 1.0 for t in items])
    assert (sc.cpu() - exp["scores"] * hot[:, None]).abs().max().item() <= TOL
    assert float(sc[3].abs().max()) == 0.0
