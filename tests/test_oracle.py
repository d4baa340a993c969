"""CPU: the restated oracle against (a) the committed golden fixtures produced by the
unmodified reference forward and (b) the live reference when /root/reference exists."""
import pytest
import torch

from oracle import kws_oracle as O
from oracle import ref_stub
from oracle.make_golden import CASES, assemble_body, load_case


def _body_fns(fe, head):
    body = lambda x: fe.pooler(fe.encoder(x).last_hidden_state).flatten(1)
    return body, head


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_golden(name):
    meta, ins, sd, outs = load_case(name)
    v = meta["variant"]
    km, um = ins["kwd_mask"], ins["utt_mask"]
    if v == "LEF":
        km, um = O.pooled_mask(km), O.pooled_mask(um)
    fe, head, same_body = assemble_body(meta, sd)
    assert same_body, "seeded body weights differ from the ones the fixture was generated with"
    body, clf = _body_fns(fe, head)
    with torch.inference_mode():
        out = O.forward_pairs(ins["kwd"], ins["utt"], km, um, sd, v, body=body, classifier=clf)
    assert torch.allclose(out["features"], outs["features"], atol=1e-6, rtol=0)
    assert torch.allclose(out["stem"], outs["stem"], atol=2e-5, rtol=1e-5)
    assert torch.allclose(out["pool"], outs["pool"], atol=2e-5, rtol=1e-5)
    if same_body:
        assert torch.allclose(out["logits"], outs["logits"], atol=1e-4, rtol=1e-4)
        sc = out["scores"] * ins["hotword_mask"][:, None]
        assert torch.allclose(sc, outs["scores"], atol=1e-5)


def test_mask_folding_identity():
    """sim * m_k * m_u == <m_k a^, m_u b^> for arbitrary (not only 0/1) masks."""
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(5, 7, 16, generator=g), torch.randn(5, 9, 16, generator=g)
    ma, mb = torch.rand(5, 7, generator=g), (torch.rand(5, 9, generator=g) > 0.5).float()
    lhs = O.sim_matrix(a, b) * ma[:, :, None] * mb[:, None, :]
    an = a / a.norm(dim=-1, keepdim=True).clamp(min=1e-6) * ma[..., None]
    bn = b / b.norm(dim=-1, keepdim=True).clamp(min=1e-6) * mb[..., None]
    assert torch.allclose(lhs, an @ bn.transpose(1, 2), atol=1e-6)


def test_zero_rows_give_zero_similarity():
    a = torch.zeros(1, 4, 8)
    b = torch.randn(1, 3, 8)
    assert torch.equal(O.sim_matrix(a, b), torch.zeros(1, 4, 3))


def test_lef_output_length_and_padding_semantics():
    sd = O.make_weights("LEF", 1, 32, 8, seed=3)
    for T in (1, 2, 7, 10):
        y = O.project_time(torch.randn(2, T, 8), sd, 0)
        assert y.shape == (2, (T + 1) // 2, 8)


@pytest.mark.skipif(not ref_stub.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("variant", ["L", "LE", "LEF"])
def test_oracle_matches_live_reference(variant):
    C, D, P, Tk, Tu, K, U = 3, 64, 16, 21, 53, 3, 2
    m = ref_stub.build_reference_model(variant, C, D, P, resnet_version="resnet-18")
    sd = dict(m.state_dict())
    sd.update(O.make_weights(variant, C, D, P, seed=11))
    m.load_state_dict(sd)
    kwd, utt, km, um, _ = O.make_inputs(K, U, C, D, Tk, Tu, seed=12, ghost_frac=0.3)
    if variant == "LEF":
        km, um = O.pooled_mask(km), O.pooled_mask(um)
    fe = m.model.feature_extractor
    body = lambda x: fe.pooler(fe.encoder(x).last_hidden_state).flatten(1)
    with torch.inference_mode():
        out = O.forward_pairs(kwd, utt, km, um, sd, variant, body=body, classifier=m.model.classifier)
        for u in range(U):
            r = m(kwd_features=kwd, utt_features=utt[u:u + 1], kwd_mask=km, utt_mask=um[u:u + 1])
            assert torch.allclose(r.features, out["features"][:, u], atol=1e-6, rtol=0)
            assert torch.allclose(r.logits, out["logits"][:, u], atol=1e-5, rtol=1e-5)


@pytest.mark.skipif(not ref_stub.available(), reason="reference tree not present (GPU box)")
def test_reference_training_style_batch_is_diagonal():
    """utt batch == keyword batch pairs keyword k with utterance k (model.py:171-178)."""
    C, D, Tk, Tu, K = 2, 32, 9, 30, 3
    m = ref_stub.build_reference_model("L", C, D)
    kwd, utt, km, um, _ = O.make_inputs(K, K, C, D, Tk, Tu, seed=5, ghost_frac=0.0)
    with torch.inference_mode():
        r = m(kwd_features=kwd, utt_features=utt, kwd_mask=km, utt_mask=um)
        full = O.sim_features(kwd, utt, km, um)
    diag = torch.stack([full[k, k] for k in range(K)])
    assert torch.allclose(r.features, diag, atol=1e-6)


@pytest.mark.skipif(not ref_stub.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("size", [(15, 75), (150, 750), None])
def test_cbw_oracle_matches_live_reference_method(size):
    """Config #4: the restated similarity + resize == the UNMODIFIED CBWhisper._calculate_cosine_similarity_matrices_
    (src/model/cb_whisper.py:189-210, source exec'd where it lies; torchvision resize(antialias=False))."""
    g = torch.Generator().manual_seed(1)
    nrm = lambda t: t / torch.linalg.norm(t, dim=-1, keepdim=True)
    kws = [nrm(torch.randn(12, t, 32, generator=g)) for t in (10, 37, 1, 64, 60)]
    utt = nrm(torch.randn(2, 12, 300, 32, generator=g))
    ref = ref_stub.reference_cbw_similarity(kws, utt, size)
    got = O.cbw_similarity_resized(kws, utt, size=size)
    assert ref.shape == got.shape
    assert torch.equal(ref, got)
    # not unit-norm inputs: the reference does a plain matmul on whatever it is given, so does the oracle
    kws2 = [k * 3.0 for k in kws]
    assert torch.equal(ref_stub.reference_cbw_similarity(kws2, utt, size), O.cbw_similarity_resized(kws2, utt, size=size))


def test_cbw_oracle_matches_golden():
    """The committed fixture (outputs of the unmodified reference method + src/model/resnet.py classifier)."""
    from oracle.make_golden import CBW_CASE, load_cbw_case

    kwd_list, utt, outs, net, same = load_cbw_case()
    imgs = O.cbw_similarity_resized(kwd_list, utt, size=CBW_CASE["size"])
    assert torch.equal(imgs, outs["images"])
    native = O.cbw_similarity_resized(kwd_list, utt, size=None)
    assert torch.equal(native[:, :, :2], outs["images_native_head"])
    assert same, "regenerated classifier differs from the one the fixture was made with"
    with torch.inference_mode():
        st = O.stem(imgs.flatten(0, 1), {"model." + k: v for k, v in net.state_dict().items()})
        assert torch.allclose(st, outs["stem"], atol=2e-5, rtol=1e-5)
        logits = net.classifier(net.feature_extractor(imgs.flatten(0, 1)).pooler_output)
    assert torch.allclose(logits.view(outs["logits"].shape), outs["logits"], atol=1e-4, rtol=1e-4)


def test_cbw_similarity_resize_shapes():
    g = torch.Generator().manual_seed(1)
    kws = [torch.randn(12, t, 32, generator=g) for t in (10, 37)]
    utt = torch.randn(2, 12, 100, 32, generator=g)
    out = O.cbw_similarity_resized(kws, utt, size=(15, 75))
    assert out.shape == (2, 2, 12, 15, 75)
