"""GPU: the drop-in module (KWSModelB200.forward / .score, through the C ABI) against the committed golden
fixtures, i.e. against what the UNMODIFIED reference ``KWSModel.forward`` returned for the same seeded inputs
and weights (oracle/make_golden.py).  Tolerances are the north-star's: similarity and logits within 2e-3
absolute (fp32 accumulation, fp32 body, TF32 off), thresholded detections identical."""
import pytest
import torch

from oracle import kws_oracle as O
from oracle.make_golden import CASES, assemble_body, load_case

pytestmark = pytest.mark.gpu
TOL = 2e-3


def build(name, dev, **opts):
    import enhance_cb_whisper_b200 as kb

    meta, ins, sd, outs = load_case(name)
    fe, head, same = assemble_body(meta, sd)
    v = meta["variant"]
    m = kb.KWSModelB200(n_layers=meta["C"], embedding_dim=meta["D"], proj_mlp_units=meta["P"],
                        learn_features=(v != "L"), proj_mlp=(v != "L"), frames_conv=(v == "LEF"),
                        resnet_version=meta["resnet_version"], features_size=(meta["Tk"], meta["Tu"]), **opts)
    full = dict(m.state_dict())
    full.update({"model.feature_extractor." + k: t for k, t in fe.state_dict().items()})
    full.update({"model.classifier." + k: t for k, t in head.state_dict().items()})
    full.update(sd)
    m.load_state_dict(full)
    m = m.to(dev).eval()
    km, um = ins["kwd_mask"], ins["utt_mask"]
    if v == "LEF":  # the reference forward needs masks at pooled resolution (SURVEY 8c deviation ii)
        km, um = O.pooled_mask(km), O.pooled_mask(um)
    dev_ins = dict(kwd=ins["kwd"].to(dev), utt=ins["utt"].to(dev), km=km.contiguous().to(dev),
                   um=um.contiguous().to(dev), hot=ins["hotword_mask"].to(dev))
    return m, meta, dev_ins, outs, same


def err(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item()


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference_outputs(built_lib, cuda_dev, name):
    """forward() driven like test_step (model.py:756-780): one utterance at a time, batch dim 1."""
    m, meta, x, outs, same = build(name, cuda_dev)
    assert same, "regenerated ResNet body differs from the one the fixture was made with"
    for u in range(meta["U"]):
        r = m(kwd_features=x["kwd"], utt_features=x["utt"][u:u + 1], kwd_mask=x["km"], utt_mask=x["um"][u:u + 1])
        assert r.features.shape == outs["features"][:, u].shape and r.logits.shape == (meta["K"], 2)
        assert err(r.features, outs["features"][:, u]) <= TOL
        assert err(r.logits, outs["logits"][:, u]) <= TOL
        assert r.loss is None and r.logits_alt is None and r.loss_alt == {"loss_diag": None, "loss_resnet": None}


@pytest.mark.parametrize("name", list(CASES))
def test_fused_forward_skips_the_similarity_tensor(built_lib, cuda_dev, name):
    """b200_return_features=False: similarity + stem in one kernel, KWSOutput.features is None (no reference
    caller reads it), logits unchanged."""
    m, meta, x, outs, _ = build(name, cuda_dev, b200_return_features=False)
    labels = torch.zeros(meta["K"], dtype=torch.long, device=cuda_dev)
    for u in range(meta["U"]):
        r = m(kwd_features=x["kwd"], utt_features=x["utt"][u:u + 1], labels=labels, kwd_mask=x["km"],
              utt_mask=x["um"][u:u + 1])
        assert r.features is None
        assert err(r.logits, outs["logits"][:, u]) <= TOL
        exp_loss = torch.nn.functional.cross_entropy(outs["logits"][:, u], labels.cpu())
        assert abs(r.loss.item() - exp_loss.item()) <= TOL
        assert r.loss_alt["loss_resnet"] is r.loss


@pytest.mark.parametrize("name", list(CASES))
def test_stem_activation_matches_reference_hook(built_lib, cuda_dev, name):
    """The in-scope end point: stem activation of every pair vs the reference's embedder output."""
    from enhance_cb_whisper_b200 import ops

    m, meta, x, outs, _ = build(name, cuda_dev)
    eng = m.prepare(cuda_dev)
    kn = eng.compress(x["kwd"], x["km"])
    un = eng.compress(x["utt"], x["um"])
    got = {}
    eng.hot_path(kn, un, ops.STEM_OUT_NCHW_F32, max_pairs=4,
                 consume=lambda k0, k1, u0, u1, st: got.__setitem__((k0, k1, u0, u1), st.clone()))
    exp = outs["stem"]  # [K,U,64,Ho,Wo]
    for (k0, k1, u0, u1), st in got.items():
        e = exp[k0:k1, u0:u1].flatten(0, 1)
        assert st.shape == e.shape
        assert err(st, e) <= TOL * max(1.0, e.abs().max().item())


@pytest.mark.parametrize("name", list(CASES))
def test_score_detections_identical(built_lib, cuda_dev, name):
    """score(): all K x U pairs in one call == the reference's per-group loop; detections bit-identical."""
    m, meta, x, outs, _ = build(name, cuda_dev)
    gold = outs["scores"]  # softmax(logits)[:,1] * hotword_mask (model.py:783-795)
    # a threshold in the widest gap of the golden scores: every pair has a margin, so detections must be identical
    s = torch.sort(gold.flatten()).values
    gaps = s[1:] - s[:-1]
    i = int(torch.argmax(gaps))
    thr = float((s[i] + s[i + 1]) / 2)
    sc, det, logits = m.score(x["kwd"], x["utt"], x["km"], x["um"], hotword_mask=x["hot"], max_pairs=5, threshold=thr)
    assert err(logits, outs["logits"]) <= TOL
    assert err(sc, gold) <= TOL
    if float(gaps[i]) > 2.5 * TOL:  # margin gap/2 > |score error| (<= |logit error| / 2): nothing may flip
        assert torch.equal(det.cpu().bool(), gold >= thr)
    sure = (gold - thr).abs() > TOL
    assert torch.equal(det.cpu().bool()[sure], (gold >= thr)[sure])
    # default threshold (hparams.threshold = 0.5): identical wherever the reference score has a margin
    _, det05, _ = m.score(x["kwd"], x["utt"], x["km"], x["um"], hotword_mask=x["hot"])
    clear = (gold - 0.5).abs() > TOL
    assert torch.equal(det05.cpu().bool()[clear], (gold >= 0.5)[clear])
    ghosts = x["hot"].cpu() == 0
    if ghosts.any():
        assert float(sc.cpu()[ghosts].abs().max()) == 0.0


@pytest.mark.parametrize("name", ["LE_small", "LEF_odd", "L_small"])
def test_batched_steps_match_reference_outputs(built_lib, cuda_dev, name):
    """test_step / validation_step override (all groups of a DataLoader item in one pass) on the GPU against what
    the reference's steps produce from the golden logits: preds = softmax(logits)[:,1] * hotword_mask
    (model.py:783-795), loss = sum over groups of the mean cross-entropy (model.py:348)."""
    m, meta, x, outs, _ = build(name, cuda_dev, b200_return_features=False)
    K, groups = meta["K"], [[0, 1], [2]]
    labels = torch.tensor([1, 0, 1][:K], device=cuda_dev)
    m.test_step_outputs, m.validation_step_outputs = [], []
    for u in range(meta["U"]):
        batch = {"utt": x["utt"][u], "utt_mask": x["um"][u], "kwd": [[x["kwd"][i] for i in g] for g in groups],
                 "kwd_mask": [[x["km"][i] for i in g] for g in groups], "hotword_labels": [labels[g] for g in groups],
                 "hotword_mask": [x["hot"][g] for g in groups], "speaker": f"s{u}"}
        m.test_step(dict(batch), u)
        m.validation_step(dict(batch), u, dataloader_idx=1)
        t, v = m.test_step_outputs[-1], m.validation_step_outputs[1][-1]
        assert t["speaker"] == f"s{u}" and torch.equal(t["targets"], labels) and torch.equal(v["targets"], labels)
        assert err(t["preds"], outs["scores"][:, u]) <= TOL and torch.equal(t["preds"], v["preds"])
        gl = outs["logits"][:, u]
        exp_loss = sum(torch.nn.functional.cross_entropy(gl[g], labels.cpu()[g]) for g in groups)
        assert abs(float(v["loss"]) - float(exp_loss)) <= 2 * TOL
    assert len(m.validation_step_outputs) == 2


def test_forward_reuses_the_compressed_utterance_across_groups(built_lib, cuda_dev):
    """Callers that loop over keyword groups with the same utterance tensor (the reference test_step, model.py:769-780)
    compress the utterance once; an in-place change of the tensor invalidates the cache."""
    from enhance_cb_whisper_b200 import ops

    m, meta, x, outs, _ = build("LE_small", cuda_dev)
    utt, um = x["utt"][0], x["um"][0]
    r0 = m(kwd_features=x["kwd"][:2], utt_features=utt.unsqueeze(0), kwd_mask=x["km"][:2], utt_mask=um.unsqueeze(0))
    n0 = ops.LAUNCHES
    r1 = m(kwd_features=x["kwd"][2:], utt_features=utt.unsqueeze(0), kwd_mask=x["km"][2:], utt_mask=um.unsqueeze(0))
    n1 = ops.LAUNCHES
    assert err(torch.cat([r0.logits, r1.logits]), outs["logits"][:, 0]) <= TOL
    utt.mul_(1.0)  # bumps the version counter: same values, but the cache must not assume that
    m(kwd_features=x["kwd"][2:], utt_features=utt.unsqueeze(0), kwd_mask=x["km"][2:], utt_mask=um.unsqueeze(0))
    n2 = ops.LAUNCHES
    assert n2 - n1 > n1 - n0  # the third call compressed the utterance again, the second did not
    utt.mul_(-1.0)  # new contents at the same address: a stale entry would return r1's features
    r3 = m(kwd_features=x["kwd"][2:], utt_features=utt.unsqueeze(0), kwd_mask=x["km"][2:], utt_mask=um.unsqueeze(0))
    assert err(r3.features, r1.features) > 1e-2
    m._utt_cache = None
    r4 = m(kwd_features=x["kwd"][2:], utt_features=utt.unsqueeze(0), kwd_mask=x["km"][2:], utt_mask=um.unsqueeze(0))
    assert torch.equal(r3.features, r4.features) and torch.equal(r3.logits, r4.logits)


@pytest.mark.parametrize("name", ["LE_small", "LEF_odd", "L_small"])
@pytest.mark.parametrize("body", ["float32", "bfloat16"])
def test_ragged_scoring_is_bit_identical_to_dense(built_lib, cuda_dev, name, body):
    """b200_ragged (keyword lengths from the frame masks carried into the fused kernel) changes no bit of the logits,
    through score(), score_host() and the bank path."""
    m, meta, x, outs, _ = build(name, cuda_dev, b200_body_dtype=body)
    m.b200_ragged = False
    sc0, det0, lg0 = m.score(x["kwd"], x["utt"], x["km"], x["um"], hotword_mask=x["hot"], max_pairs=4)
    m.b200_ragged = True
    sc1, det1, lg1 = m.score(x["kwd"], x["utt"], x["km"], x["um"], hotword_mask=x["hot"], max_pairs=4)
    assert torch.equal(lg0, lg1) and torch.equal(sc0, sc1) and torch.equal(det0, det1)
    pin = lambda t: t.cpu().contiguous().pin_memory()
    sc2, det2, lg2 = m.score_host(pin(x["kwd"]), pin(x["utt"]), pin(x["km"]), pin(x["um"]), hotword_mask=pin(x["hot"]),
                                  max_pairs=4, kwd_slab=2, utt_slab=1, device=cuda_dev)
    assert torch.equal(lg0, lg2) and torch.equal(det0, det2)
    if body == "float32":
        assert err(lg1, outs["logits"]) <= TOL


def test_training_style_batch_is_diagonal(built_lib, cuda_dev):
    """utt batch == keyword batch: pair k with utterance k (model.py:171-173; training_step batches)."""
    m, meta, x, outs, _ = build("LE_small", cuda_dev)
    K, U = meta["K"], meta["U"]
    idx = [k % U for k in range(K)]
    utt = x["utt"][idx].contiguous()
    um = x["um"][idx].contiguous()
    r = m(kwd_features=x["kwd"], utt_features=utt, kwd_mask=x["km"], utt_mask=um)
    for k in range(K):
        assert err(r.features[k], outs["features"][k, idx[k]]) <= TOL
        assert err(r.logits[k], outs["logits"][k, idx[k]]) <= TOL


def test_interface_errors_mirror_the_reference(built_lib, cuda_dev):
    m, meta, x, _, _ = build("L_small", cuda_dev)
    with pytest.raises(AttributeError):  # reference: None.unsqueeze (model.py:187-191)
        m(kwd_features=x["kwd"], utt_features=x["utt"][:1])
    with pytest.raises(ValueError):  # HF channel check (modeling_resnet.py:71-75)
        m(kwd_features=x["kwd"][:, :2], utt_features=x["utt"][:1, :2], kwd_mask=x["km"][:, :2], utt_mask=x["um"][:1, :2])
    m.train()
    with pytest.raises(RuntimeError):  # inference only: BatchNorm is folded
        m(kwd_features=x["kwd"], utt_features=x["utt"][:1], kwd_mask=x["km"], utt_mask=x["um"][:1])
    m.eval()
    from enhance_cb_whisper_b200 import KWSError

    with pytest.raises(KWSError):  # no CPU fallback
        m.prepare(torch.device("cpu"))


def test_lef_many_layers_runs_fused_in_passes(built_lib, cuda_dev):
    """LEF with 16 layers (cfg3/cfg5 style): the throughput path runs the fused kernel in passes over groups of
    12 layers; its bf16 stem activation matches the oracle's fp32 stem."""
    import enhance_cb_whisper_b200 as kb
    from enhance_cb_whisper_b200 import ops

    C, D, P, Tk, Tu, K, U = 16, 128, 64, 22, 70, 3, 2
    torch.manual_seed(2)
    m = kb.KWSModelB200(n_layers=C, embedding_dim=D, proj_mlp_units=P, learn_features=True, proj_mlp=True,
                        frames_conv=True, resnet_version="resnet-18", features_size=(Tk, Tu))
    sd = O.make_weights("LEF", C, D, P, seed=77)
    full = dict(m.state_dict())
    full.update(sd)
    m.load_state_dict(full)
    m = m.to(cuda_dev).eval()
    kwd, utt, km, um, _ = O.make_inputs(K, U, C, D, Tk, Tu, seed=78, ghost_frac=0.0)
    km, um = O.pooled_mask(km).contiguous(), O.pooled_mask(um).contiguous()
    exp = O.forward_pairs(kwd, utt, km, um, sd, "LEF", upto="stem")["stem"].flatten(0, 1)
    eng = m.prepare(cuda_dev)
    kn = eng.compress(kwd.to(cuda_dev), km.to(cuda_dev))
    un = eng.compress(utt.to(cuda_dev), um.to(cuda_dev))
    assert eng.fused(kn.shape[2], un.shape[2], ops.STEM_OUT_NHWC_BF16)
    got = []
    eng.hot_path(kn, un, ops.STEM_OUT_NHWC_BF16, max_pairs=K * U, consume=lambda *a: got.append(a[-1].float().clone()))
    assert got[0].shape == exp.shape
    assert err(got[0], exp) <= 2e-2 * max(1.0, exp.abs().max().item() / 4)


def test_throughput_body_matches_unfused_modules(built_lib, cuda_dev):
    """b200_body_dtype="bfloat16": BatchNorms folded + cuDNN fused conv+bias(+residual)+ReLU + kws_maxpool_nhwc
    (body.py) against the unmodified HF modules run in fp32 on the same bf16 stem activation, and against the
    same modules cast to bf16 ("bfloat16_unfused"): the folded path must be no further from fp32 than the
    cast modules are (plus slack), and detections on clear margins agree."""
    import enhance_cb_whisper_b200 as kb
    from enhance_cb_whisper_b200 import ops
    from enhance_cb_whisper_b200.model import run_body

    m, meta, x, outs, _ = build("LE_small", cuda_dev, b200_body_dtype="bfloat16")
    eng = m.prepare(cuda_dev)
    kn, un = eng.compress(x["kwd"], x["km"]), eng.compress(x["utt"], x["um"])
    st = []
    eng.hot_path(kn, un, ops.STEM_OUT_NHWC_BF16, max_pairs=64, consume=lambda *a: st.append(a[-1].clone()))
    st = torch.cat(st)
    ref32 = run_body(m.model, st.float())
    fused = m._body(st)
    assert fused.dtype == torch.float32 and fused.shape == ref32.shape
    m.b200_body_dtype = "bfloat16_unfused"
    cast = m._body(st)
    m.b200_body_dtype = "bfloat16"
    scale = max(1.0, ref32.abs().max().item())
    e_fused, e_cast = err(fused, ref32) / scale, err(cast, ref32) / scale
    assert e_fused <= max(2.0 * e_cast, 2e-2), (e_fused, e_cast)
    # scores through the public call
    sc, det, logits = m.score(x["kwd"], x["utt"], x["km"], x["um"], hotword_mask=x["hot"])
    gold = outs["scores"]
    clear = (gold - 0.5).abs() > 0.1
    assert torch.equal(det.cpu().bool()[clear], (gold >= 0.5)[clear])
    with pytest.raises(ValueError):
        m.b200_body_dtype = "int8"
        m._body(st)


@pytest.mark.parametrize("name", ["LE_small", "LEF_odd"])
def test_score_host_streams_slabs_and_matches_score(built_lib, cuda_dev, name):
    """score_host(): pinned host inputs, keyword slabs uploaded on a copy stream while the previous slab is
    scored == score() on device-resident inputs (slab size not dividing K, twice to reuse the staging)."""
    m, meta, x, outs, _ = build(name, cuda_dev)
    pin = lambda t: t.cpu().contiguous().pin_memory()
    h = {k: pin(v) for k, v in x.items()}
    sc0, det0, lg0 = m.score(x["kwd"], x["utt"], x["km"], x["um"], hotword_mask=x["hot"], max_pairs=3)
    for slab in (3, 1, meta["K"] + 5):
        sc, det, lg = m.score_host(h["kwd"], h["utt"], h["km"], h["um"], hotword_mask=h["hot"], max_pairs=3,
                                   kwd_slab=slab, device=cuda_dev)
        assert sc.is_cuda and err(lg, lg0) <= 1e-4 and err(sc, sc0) <= 1e-5
        assert err(lg, outs["logits"]) <= TOL
        clear = (sc0 - 0.5).abs() > 1e-3
        assert torch.equal(det[clear], det0[clear])


@pytest.mark.parametrize("name", ["LE_small", "LEF_odd", "L_small"])
def test_fused_pool_scoring_is_bit_identical_to_the_separate_maxpool(built_lib, cuda_dev, name):
    """b200_fused_pool (kws_sim_stem_pool: MaxPool2d(3,2,1) inside the fused kernel) vs the stem activation in HBM +
    kws_maxpool_nhwc: the body sees the same bf16 tensor, so logits, scores and detections are equal bit for bit --
    through score(), the batched step and forward(return_features=False)."""
    m, meta, x, outs, _ = build(name, cuda_dev, b200_body_dtype="bfloat16", b200_return_features=False)
    assert m.b200_fused_pool
    a = m.score(x["kwd"], x["utt"], x["km"], x["um"], hotword_mask=x["hot"], max_pairs=5)
    ra = m(kwd_features=x["kwd"], utt_features=x["utt"][:1], kwd_mask=x["km"], utt_mask=x["um"][:1]).logits
    m.b200_fused_pool = False
    b = m.score(x["kwd"], x["utt"], x["km"], x["um"], hotword_mask=x["hot"], max_pairs=5)
    rb = m(kwd_features=x["kwd"], utt_features=x["utt"][:1], kwd_mask=x["km"], utt_mask=x["um"][:1]).logits
    for t, u in zip(a, b):
        assert torch.equal(t, u)
    assert torch.equal(ra, rb)


def test_hot_path_pooled_output_fused_and_unfused_agree(built_lib, cuda_dev):
    """engine.hot_path(STEM_OUT_POOL_NHWC_BF16): the fused kernel's pooled activation vs the fallback for shapes the fused
    kernel does not cover (kws_sim -> kws_stem -> kws_maxpool_nhwc, forced here): same shape, same values up to the fp32
    summation order of the two stem kernels before the bf16 rounding."""
    from enhance_cb_whisper_b200 import ops

    m, meta, x, outs, _ = build("LE_small", cuda_dev, b200_body_dtype="bfloat16", b200_return_features=False)
    eng = m.prepare(cuda_dev)
    kn, un = eng.compress(x["kwd"], x["km"]), eng.compress(x["utt"], x["um"])
    got = {}
    eng.hot_path(kn, un, ops.STEM_OUT_POOL_NHWC_BF16, 4, lambda k0, k1, u0, u1, st: got.__setitem__(("f", k0, u0), st.float().clone()))
    eng.fused = lambda *a, **k: False
    eng.hot_path(kn, un, ops.STEM_OUT_POOL_NHWC_BF16, 4, lambda k0, k1, u0, u1, st: got.__setitem__(("u", k0, u0), st.float().clone()))
    keys = [k for k in got if k[0] == "f"]
    assert keys
    Ho, Wo = (meta["Tk"] + 1) // 2, (meta["Tu"] + 1) // 2
    for _, k0, u0 in keys:
        a, b = got[("f", k0, u0)], got[("u", k0, u0)]
        assert a.shape == b.shape and a.shape[2:] == ((Ho + 1) // 2, (Wo + 1) // 2)
        assert (a - b).abs().max().item() <= 2e-2 * max(1.0, b.abs().max().item() / 4)
