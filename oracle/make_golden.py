"""Generate tests/golden/*.npz from the UNMODIFIED reference forward
(TEST INFRASTRUCTURE; run in the build container where /root/reference exists).

    python -m oracle.make_golden            # writes tests/golden/kws_<case>.npz

Each fixture holds seeded inputs, the hot-path weights (projector /
time_projector / stem, reference state_dict key names), and what the reference
``KWSModel.forward`` (src/efficient_kws/model.py:129-221) returned for every
utterance: ``features`` (KWSOutput.features), the stem activation (forward hook
on ``model.feature_extractor.embedder.embedder``), the pooled stem activation,
``logits`` and the detection score ``softmax(logits)[:, 1] * hotword_mask``
(model.py:783-795).  ResNet body weights are too large to commit; they are
regenerated from ``body_seed`` by ``build_body`` (same image => same
torch RNG stream) and guarded by ``body_checksum``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import kws_oracle as O
from . import ref_stub

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name: (variant, K, U, C, D, P, Tk, Tu, resnet_version, ghost_frac)
CASES = {
    "L_small": ("L", 3, 2, 3, 128, 64, 22, 70, "resnet-50", 0.0),
    "LE_small": ("LE", 3, 2, 3, 128, 64, 22, 70, "resnet-18", 0.34),
    "LEF_odd": ("LEF", 3, 2, 3, 128, 64, 23, 71, "resnet-18", 0.0),
    "LE_wide": ("LE", 2, 1, 4, 256, 64, 40, 300, "resnet-34", 0.0),
}


def body_checksum(model) -> float:
    tot = 0.0
    for k, v in model.state_dict().items():
        if v.is_floating_point():
            tot += float(v.double().abs().sum())
    return tot


def build_body(C: int, resnet_version: str, body_seed: int):
    """The HF ResNet (+ Linear head) exactly as src/efficient_kws/resnet.py:7-58
    builds it, restated so that it can be rebuilt where the reference tree is
    absent; seeded so fixtures do not have to carry 25 M parameters."""
    from transformers import ResNetConfig, ResNetModel

    torch.manual_seed(body_seed)
    cfg = ResNetConfig()
    if resnet_version == "resnet-18":
        cfg.layer_type, cfg.hidden_sizes, cfg.depths = "basic", [64, 128, 256, 512], [2, 2, 2, 2]
    elif resnet_version == "resnet-34":
        cfg.layer_type, cfg.hidden_sizes, cfg.depths = "basic", [64, 128, 256, 512], [3, 4, 6, 3]
    cfg.num_channels = C
    cfg.num_labels = 2
    fe = ResNetModel(cfg)
    head = torch.nn.Sequential(torch.nn.Flatten(1, -1), torch.nn.Linear(cfg.hidden_sizes[-1], 2, bias=True))
    return fe.eval(), head.eval()


def assemble_body(meta, sd):
    """Rebuild the seeded body for a fixture and load the fixture's stem weights into it.
    Returns (feature_extractor, head, same) where ``same`` says whether the regenerated
    weights reproduce the checksum recorded when the fixture was made."""
    fe, head = build_body(meta["C"], meta["resnet_version"], meta["body_seed"])
    stem_sd = {k[len("model.feature_extractor."):]: v for k, v in sd.items() if k.startswith("model.feature_extractor.")}
    fe.load_state_dict(stem_sd, strict=False)
    tot = 0.0
    for mod in (fe, head):
        for v in mod.state_dict().values():
            if v.is_floating_point():
                tot += float(v.double().abs().sum())
    same = abs(tot - meta["body_checksum"]) <= 1e-9 * max(1.0, abs(meta["body_checksum"]))
    return fe, head, same


def make_case(name: str, body_seed: int = 7):
    variant, K, U, C, D, P, Tk, Tu, rv, ghost = CASES[name]
    if variant == "L":
        rv = "resnet-50"  # model.py:74-76 ignores resnet_version for L
    m = ref_stub.build_reference_model(variant, C, D, P, resnet_version=rv, features_size=(Tk, Tu))
    fe, head = build_body(C, rv, body_seed)
    sd = {"model.feature_extractor." + k: v for k, v in fe.state_dict().items()}
    sd.update({"model.classifier." + k: v for k, v in head.state_dict().items()})
    hot = O.make_weights(variant, C, D, P, seed=100 + len(name))
    full = dict(m.state_dict())
    full.update(sd)
    full.update(hot)
    m.load_state_dict(full)
    kwd, utt, km, um, hm = O.make_inputs(K, U, C, D, Tk, Tu, seed=200 + len(name), ghost_frac=ghost)
    km_s, um_s = (O.pooled_mask(km), O.pooled_mask(um)) if variant == "LEF" else (km, um)

    grabbed = {}
    emb = m.model.feature_extractor.embedder
    h1 = emb.embedder.register_forward_hook(lambda mod, i, o: grabbed.__setitem__("stem", o.detach().clone()))
    h2 = emb.pooler.register_forward_hook(lambda mod, i, o: grabbed.__setitem__("pool", o.detach().clone()))
    feats, stems, pools, logits = [], [], [], []
    with torch.inference_mode():
        for u in range(U):
            # driven like test_step (model.py:756-780): one utterance, batch dim 1
            r = m(kwd_features=kwd, utt_features=utt[u : u + 1], kwd_mask=km_s, utt_mask=um_s[u : u + 1])
            feats.append(r.features)
            stems.append(grabbed["stem"])
            pools.append(grabbed["pool"])
            logits.append(r.logits)
    h1.remove()
    h2.remove()
    features = torch.stack(feats, 1)  # [K,U,C,Tk',Tu']
    logits = torch.stack(logits, 1)  # [K,U,2]
    scores = logits.softmax(-1)[..., 1] * hm[:, None]
    out = {
        "meta_variant": np.array(variant),
        "meta_resnet_version": np.array(rv),
        "meta_dims": np.array([K, U, C, D, P, Tk, Tu], dtype=np.int64),
        "meta_body_seed": np.array(body_seed, dtype=np.int64),
        "meta_body_checksum": np.array(body_checksum(m.model), dtype=np.float64),
        "in_kwd": kwd.numpy(),
        "in_utt": utt.numpy(),
        "in_kwd_mask": km.numpy(),
        "in_utt_mask": um.numpy(),
        "in_hotword_mask": hm.numpy(),
        "out_features": features.numpy(),
        "out_stem": torch.stack(stems, 1).numpy().astype(np.float32),
        "out_pool": torch.stack(pools, 1).numpy().astype(np.float32),
        "out_logits": logits.numpy(),
        "out_scores": scores.numpy(),
    }
    for k, v in hot.items():
        out["w_" + k] = v.numpy()
    return out


def load_case(name: str):
    """Load a committed fixture -> (meta dict, inputs dict, weights sd, outputs dict)."""
    z = np.load(os.path.join(GOLDEN_DIR, f"kws_{name}.npz"))
    K, U, C, D, P, Tk, Tu = [int(x) for x in z["meta_dims"]]
    meta = dict(
        variant=str(z["meta_variant"]),
        resnet_version=str(z["meta_resnet_version"]),
        K=K, U=U, C=C, D=D, P=P, Tk=Tk, Tu=Tu,
        body_seed=int(z["meta_body_seed"]),
        body_checksum=float(z["meta_body_checksum"]),
    )
    ins = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")}
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_")}
    outs = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out_")}
    return meta, ins, sd, outs


# ----------------------------------------------------------------------------------------------------------------
# Config #4 (original CB-Whisper classifier): images from the UNMODIFIED CBWhisper._calculate_cosine_similarity_matrices_
# (src/model/cb_whisper.py:189-210) and logits from the UNMODIFIED 12-channel classifier src/model/resnet.py:5-38 (what
# src/model/model.py:78-93 runs on them).
# ----------------------------------------------------------------------------------------------------------------
CBW_CASE = dict(C=12, D=64, Tu=100, S=2, lens=(9, 21, 14, 64, 1), size=(30, 50), seed=41, body_seed=43)


def cbw_inputs(case=CBW_CASE):
    g = torch.Generator().manual_seed(case["seed"])
    nrm = lambda t: t / torch.linalg.norm(t, dim=-1, keepdim=True)  # cb_whisper.py:106 (no eps)
    kwd_list = [nrm(torch.randn(case["C"], t, case["D"], generator=g)) for t in case["lens"]]
    utt = nrm(torch.randn(case["S"], case["C"], case["Tu"], case["D"], generator=g))
    return kwd_list, utt


def reference_cbw_resnet(body_seed: int):
    """src/model/resnet.py (imports only torch + transformers), loaded where it lies; seeded init + non-trivial
    stem BatchNorm statistics."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_cbw_resnet", os.path.join(ref_stub.REFERENCE_SRC, "model", "resnet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(body_seed)
    net = mod.Resnet(num_channels=12, num_classes=2).eval()
    randomize_stem_bn(net, body_seed + 1)
    return net


def randomize_stem_bn(net, seed: int):
    g = torch.Generator().manual_seed(seed)
    bn = net.feature_extractor.embedder.embedder.normalization
    with torch.no_grad():
        bn.weight.copy_(torch.rand(64, generator=g) + 0.5), bn.bias.copy_(torch.randn(64, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(64, generator=g) * 0.1), bn.running_var.copy_(torch.rand(64, generator=g) + 0.5)


def make_cbw_case(case=CBW_CASE):
    kwd_list, utt = cbw_inputs(case)
    imgs = ref_stub.reference_cbw_similarity(kwd_list, utt, case["size"])  # [K,S,C,h,w]
    imgs_native = ref_stub.reference_cbw_similarity(kwd_list, utt, None)  # kws_features_size unset
    net = reference_cbw_resnet(case["body_seed"])
    grabbed = {}
    h = net.feature_extractor.embedder.embedder.register_forward_hook(lambda m, i, o: grabbed.__setitem__("stem", o.detach().clone()))
    with torch.inference_mode():
        logits = net(imgs.flatten(0, 1)).view(len(kwd_list), case["S"], 2)
    h.remove()
    return {
        "images": imgs.numpy(), "images_native_head": imgs_native[:, :, :2].numpy(),
        "stem": grabbed["stem"].numpy().astype(np.float32), "logits": logits.numpy(),
        "body_checksum": np.array(body_checksum(net), dtype=np.float64),
    }


def load_cbw_case(case=CBW_CASE):
    """-> (kwd_list, utt, outputs dict, classifier with the attribute names of the reference wrapper, same_body)."""
    z = np.load(os.path.join(GOLDEN_DIR, "cbw_small.npz"))
    kwd_list, utt = cbw_inputs(case)
    from transformers import ResNetConfig, ResNetModel

    class _Net(torch.nn.Module):  # restated src/model/resnet.py:5-38 (same construction order => same seeded weights)
        def __init__(self):
            super().__init__()
            self.config = ResNetConfig()
            self.config.num_channels = 12
            self.config.num_labels = 2
            self.feature_extractor = ResNetModel(self.config)
            self.classifier = torch.nn.Sequential(torch.nn.Flatten(1, -1), torch.nn.Linear(self.config.hidden_sizes[-1], 2, bias=True))

    torch.manual_seed(case["body_seed"])
    net = _Net().eval()
    randomize_stem_bn(net, case["body_seed"] + 1)
    ck = float(z["body_checksum"])
    same = abs(body_checksum(net) - ck) <= 1e-9 * max(1.0, abs(ck))
    return kwd_list, utt, {k: torch.from_numpy(z[k]) for k in z.files if k != "body_checksum"}, net, same


def main():
    if not ref_stub.available():
        sys.exit("reference tree not present; golden fixtures can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = set(sys.argv[1:])
    for name in CASES:
        if only and name not in only:
            continue
        d = make_case(name)
        path = os.path.join(GOLDEN_DIR, f"kws_{name}.npz")
        np.savez_compressed(path, **d)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
    path = os.path.join(GOLDEN_DIR, "cbw_small.npz")
    np.savez_compressed(path, **make_cbw_case())
    print(f"cbw_small: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
