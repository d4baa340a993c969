"""Import the UNMODIFIED reference ``efficient_kws`` model under stub modules
(TEST INFRASTRUCTURE; works only where /root/reference exists, i.e. in the
build container -- never on the GPU box).

``pytorch_lightning``, ``torchmetrics`` and ``confidence_intervals`` are absent
from this image and cannot be installed (no network); the reference's
``src/efficient_kws/model.py`` imports them at module scope but its ``forward``
/ ``sim_matrix`` / ``resnet_forward`` (model.py:129-221) use none of them, so
minimal stand-ins are enough to execute the reference arithmetic as written.
No reference source is copied: the files are imported where they lie.
"""
from __future__ import annotations

import inspect
import os
import sys
import types

import torch.nn as nn

REFERENCE_SRC = os.environ.get("KWS_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "efficient_kws", "model.py"))


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class _LightningModule(nn.Module):
    """Stand-in for pl.LightningModule: an nn.Module whose
    ``save_hyperparameters()`` collects the caller's __init__ arguments."""

    def save_hyperparameters(self, *args, **kwargs):
        frame = inspect.currentframe().f_back
        local = frame.f_locals
        hp = _AttrDict()
        for k, v in local.items():
            if k in ("self", "__class__"):
                continue
            if k == "kwargs" and isinstance(v, dict):
                hp.update(v)
            else:
                hp[k] = v
        object.__setattr__(self, "hparams", hp)

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = _LightningModule
        sys.modules["pytorch_lightning"] = pl
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class PrecisionRecallCurve(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

        class Accuracy(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

        tm.PrecisionRecallCurve = PrecisionRecallCurve
        tm.Accuracy = Accuracy
        sys.modules["torchmetrics"] = tm
    if "confidence_intervals" not in sys.modules:
        ci = types.ModuleType("confidence_intervals")
        ci.evaluate_with_conf_int = lambda *a, **k: None
        sys.modules["confidence_intervals"] = ci


def load_reference():
    """Returns the reference module ``efficient_kws.model`` (unmodified)."""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import efficient_kws.model as ref_model  # noqa: E402

    return ref_model


def build_reference_model(variant: str, C: int, D: int, P: int = 64, resnet_version: str = "resnet-50",
                          features_size=(150, 1500), threshold: float = 0.5):
    """Instantiate the reference KWSModel for a variant.

    L uses ``learn_features=False`` because the shipped L YAMLs
    (learn_features: true, proj_mlp: false) construct no ResNet and cannot run
    (SURVEY.md section 4 item 1); with learn_features=False the reference always
    builds a ResNet-50 (model.py:71-76).
    """
    ref = load_reference()
    kw = dict(
        n_layers=C,
        embedding_dim=D,
        proj_mlp_units=P,
        features_size=tuple(features_size),
        resnet_version=resnet_version,
        threshold=threshold,
    )
    if variant == "L":
        kw.update(learn_features=False, proj_mlp=False, frames_conv=False)
    elif variant == "LE":
        kw.update(learn_features=True, proj_mlp=True, frames_conv=False)
    elif variant == "LEF":
        kw.update(learn_features=True, proj_mlp=True, frames_conv=True)
    else:
        raise ValueError(variant)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # reference prints the threshold
        m = ref.KWSModel(**kw)
    return m.eval()


def load_cbw_similarity_method():
    """The UNMODIFIED ``CBWhisper._calculate_cosine_similarity_matrices_`` (src/model/cb_whisper.py:189-210) as a
    plain function ``f(self, utt_hs, kwd_hs)``.

    ``src/model/cb_whisper.py`` cannot be imported here (it needs ``whisper``, ``string2string``, ``pytorch_lightning``
    and a sys.path hack at :10-11), so the method's source text is cut out of the file where it lies with ``ast`` and
    exec'd with the three names it uses (``torch``, ``torchvision``, ``List``).  Nothing is copied into the repo.
    ``self`` only needs ``self.hparams.kws_features_size``."""
    import ast
    from typing import List

    import torch
    import torchvision

    path = os.path.join(REFERENCE_SRC, "model", "cb_whisper.py")
    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "_calculate_cosine_similarity_matrices_":
            fn_src = ast.get_source_segment(src, node)
            break
    else:  # pragma: no cover
        raise RuntimeError("_calculate_cosine_similarity_matrices_ not found in " + path)
    import textwrap

    ns = {"torch": torch, "torchvision": torchvision, "List": List}
    exec(compile(textwrap.dedent(fn_src), path, "exec"), ns)
    return ns["_calculate_cosine_similarity_matrices_"]


def reference_cbw_similarity(kwd_list, utt_hs, size=(150, 750)):
    """Run the unmodified method -> [K, S, C, size0, size1] (the reference returns a per-segment list of
    [K, C, size0, size1]; stacked here keyword-major like the oracle)."""
    import torch

    fn = load_cbw_similarity_method()
    self = types.SimpleNamespace(hparams=types.SimpleNamespace(kws_features_size=size))
    per_seg = fn(self, utt_hs=utt_hs, kwd_hs=list(kwd_list))
    return torch.stack(per_seg, dim=1)
