"""CPU (build container only: needs /root/reference): the LightningCLI drop-in
``enhance_cb_whisper_b200.lightning.KWSModelB200`` under the stub modules of oracle/ref_stub -- construction
from YAML-style init args, state_dict keys, the legacy-checkpoint remap (model.py:931-952) and the host logic
of the batched test_step / validation_step override against the UNMODIFIED reference steps
(model.py:748-802, :304-385).  The B200 arithmetic itself is covered by the -m gpu tests
(tests/test_gpu_model.py::test_batched_steps_match_reference_outputs)."""
import copy

import pytest
import torch

from oracle import kws_oracle as O
from oracle import ref_stub

pytestmark = pytest.mark.skipif(not ref_stub.available(), reason="reference tree not present (GPU box)")

YAML_ARGS = dict(  # src/efficient_kws/configs/eval-LE-comp-acl.yaml:123-164, sizes reduced
    num_domains=72, sampling="utterance-examples", kw_type="tts", kw_p=0.5, features_size=(22, 70), learn_features=True,
    load_embeddings=True, n_layers=3, embedding_dim=64, frames_conv=False, proj_mlp=True, proj_mlp_units=16,
    resnet_version="resnet-18", threshold=0.37, sru_hidden_size=128, sru_num_layers=2)  # sru_*: dead keys via **kwargs


@pytest.fixture(scope="module")
def classes(built_lib):
    ref = ref_stub.load_reference()
    from enhance_cb_whisper_b200 import lightning

    return ref.KWSModel, lightning.KWSModelB200


def _quiet(fn, *a, **k):
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_constructs_from_yaml_args_with_reference_keys(classes):
    Ref, B200 = classes
    from enhance_cb_whisper_b200.model import B200ForwardMixin

    m = _quiet(B200, **YAML_ARGS, b200_body_dtype="bfloat16", b200_return_features=False)
    r = _quiet(Ref, **YAML_ARGS)
    assert [c.__name__ for c in type(m).__mro__[:3]] == ["KWSModelB200", "B200ForwardMixin", "KWSModel"]
    assert isinstance(m, Ref) and isinstance(m, B200ForwardMixin)
    assert list(m.state_dict().keys()) == list(r.state_dict().keys())
    assert m.hparams.threshold == 0.37 and m.hparams.sru_hidden_size == 128
    assert m.b200_body_dtype == "bfloat16" and m.b200_return_features is False and m.variant == "LE"
    # only forward and the two eval steps are replaced; every other hook is the reference's own function
    for name in ("on_test_epoch_end", "on_validation_epoch_end", "training_step", "configure_optimizers",
                 "on_load_checkpoint", "sim_matrix"):
        assert getattr(B200, name) is getattr(Ref, name), name
    for name in ("forward", "test_step", "validation_step"):
        assert getattr(B200, name) is getattr(B200ForwardMixin, name), name
    m.load_state_dict(r.state_dict())  # a reference checkpoint loads unchanged


def test_shipped_L_yaml_flags_behave_like_the_reference(classes):
    """learn_features: true + proj_mlp: false (the shipped L YAMLs) builds no classifier in the reference
    (model.py:71-85): both constructors succeed and the first use raises AttributeError."""
    Ref, B200 = classes
    import enhance_cb_whisper_b200 as kb

    args = dict(YAML_ARGS, proj_mlp=False)
    for cls in (Ref, B200, kb.KWSModelB200):
        m = _quiet(cls, **args)
        assert not hasattr(m, "model")
    with pytest.raises(AttributeError):
        _quiet(B200, **args).prepare(torch.device("cuda"))


def test_legacy_checkpoint_remap_is_the_references(classes):
    Ref, B200 = classes
    import enhance_cb_whisper_b200 as kb

    m = _quiet(B200, **YAML_ARGS)
    new = m.state_dict()
    legacy = {}
    for k, v in new.items():  # early checkpoints: model.resnet.{embedder,encoder}.* and model.resnet.classifier.*
        if k.startswith("model.feature_extractor."):
            legacy["model.resnet." + k[len("model.feature_extractor."):]] = v
        elif k.startswith("model."):
            legacy["model.resnet." + k[len("model."):]] = v
        else:
            legacy[k] = v
    ckpt = {"state_dict": dict(legacy)}
    m.on_load_checkpoint(ckpt)  # inherited from the reference (model.py:931-952)
    assert sorted(ckpt["state_dict"]) == sorted(new)
    ours = kb.KWSModelB200.remap_legacy_state_dict(legacy)
    assert sorted(ours) == sorted(new)
    for k in new:
        assert torch.equal(ours[k], ckpt["state_dict"][k])
    m.load_state_dict(ckpt["state_dict"])


def _batch(K_groups=(2, 1), C=3, D=64, Tk=22, Tu=70, seed=5):
    K = sum(K_groups)
    kwd, utt, km, um, hot = O.make_inputs(K, 1, C, D, Tk, Tu, seed=seed, ghost_frac=0.34)
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, 2, (K,), generator=g)
    it = iter(range(K))
    idx = [[next(it) for _ in range(n)] for n in K_groups]
    return {
        "utt": utt[0], "utt_mask": um[0],
        "kwd": [[kwd[i] for i in grp] for grp in idx],
        "kwd_mask": [[km[i] for i in grp] for grp in idx],
        "hotword_labels": [labels[grp] for grp in idx],
        "hotword_mask": [hot[grp] for grp in idx],
        "speaker": "spk-1",
    }


@pytest.mark.parametrize("with_hotword_mask", [True, False])
def test_batched_steps_equal_reference_steps(classes, with_hotword_mask):
    """Host logic of the override: with the B200 arithmetic replaced by the reference's own forward on the stacked
    keywords, test_step / validation_step append exactly what the unmodified reference steps append."""
    Ref, B200 = classes
    torch.manual_seed(0)
    r = _quiet(Ref, **YAML_ARGS).eval()
    sd = dict(r.state_dict())
    sd.update(O.make_weights("LE", 3, 64, 16, seed=3))
    r.load_state_dict(sd)
    m = _quiet(B200, **YAML_ARGS).eval()
    m.load_state_dict(sd)
    calls = []

    def via_reference(kwd, utt, kwd_mask, utt_mask):
        calls.append(kwd.shape[0])
        return Ref.forward(m, kwd_features=kwd, utt_features=utt, kwd_mask=kwd_mask, utt_mask=utt_mask).logits

    m._b200_group_logits = via_reference
    batch = _batch()
    if not with_hotword_mask:
        batch.pop("hotword_mask")
    with torch.inference_mode():
        r.on_test_epoch_start(), m.on_test_epoch_start()
        r.test_step(copy.deepcopy(batch), 0)
        m.test_step(copy.deepcopy(batch), 0)
        r.on_validation_epoch_start(), m.on_validation_epoch_start()
        for dl in (0, 2):
            r.validation_step(copy.deepcopy(batch), 0, dataloader_idx=dl)
            m.validation_step(copy.deepcopy(batch), 0, dataloader_idx=dl)
    assert calls == [3, 3, 3]  # one scoring pass per step, all groups stacked
    a, b = r.test_step_outputs[0], m.test_step_outputs[0]
    assert a["speaker"] == b["speaker"] and torch.equal(a["targets"], b["targets"])
    assert torch.allclose(a["preds"], b["preds"], atol=1e-6, rtol=0)
    assert len(r.validation_step_outputs) == len(m.validation_step_outputs) == 3
    for dl in (0, 2):
        a, b = r.validation_step_outputs[dl][-1], m.validation_step_outputs[dl][-1]
        assert torch.equal(a["targets"], b["targets"]) and a["loss_alt"] is b["loss_alt"] is None
        assert torch.allclose(a["preds"], b["preds"], atol=1e-6, rtol=0)
        assert torch.allclose(a["loss"], b["loss"], atol=1e-6, rtol=0)


def test_training_mode_steps_fall_back_to_the_reference(classes):
    Ref, B200 = classes
    m = _quiet(B200, **YAML_ARGS)
    m.train()
    assert m._b200_reference_hook("test_step").__func__ is Ref.test_step
    with pytest.raises(RuntimeError):  # forward itself refuses training mode (BatchNorm folded)
        b = _batch()
        m(kwd_features=torch.stack(b["kwd"][0]), utt_features=b["utt"][None], kwd_mask=torch.stack(b["kwd_mask"][0]),
          utt_mask=b["utt_mask"][None])
