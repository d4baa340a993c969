"""LightningCLI drop-in: the reference ``KWSModel`` with the B200 forward.

Usable where ``pytorch_lightning`` and the reference package ``efficient_kws``
are importable (i.e. inside a checkout of the reference, ``src/`` on the path).
Switch one line of any ``src/efficient_kws/configs/*.yaml``::

    model:
      class_path: enhance_cb_whisper_b200.lightning.KWSModelB200

``run_efficient_kws.py`` (``subclass_mode_model=True``, :50) instantiates it with
the unchanged ``init_args``; ``test_step`` / ``validation_step`` and every
metric hook are inherited from the reference, only ``forward`` is replaced.
Optional extra init args: ``b200_body_dtype`` ("float32" | "bfloat16"),
``b200_return_features`` (bool), ``b200_layer_idx`` (list of int), ``b200_mlp_dtype``
("float16" | "bfloat16").
"""
from __future__ import annotations

try:
    from efficient_kws.model import KWSModel as _ReferenceKWSModel
except Exception as exc:  # pragma: no cover - depends on the deployment
    raise ImportError(
        "enhance_cb_whisper_b200.lightning needs the reference package `efficient_kws` "
        "(Enhance-CB-Whisper/src on sys.path) and pytorch_lightning; use "
        "enhance_cb_whisper_b200.KWSModelB200 for the standalone nn.Module"
    ) from exc

from .model import B200ForwardMixin


class KWSModelB200(B200ForwardMixin, _ReferenceKWSModel):
    def __init__(self, *args, b200_body_dtype: str = "float32", b200_return_features: bool = True,
                 b200_layer_idx=None, b200_mlp_dtype: str = "float16", **kwargs):
        super().__init__(*args, **kwargs)
        self.b200_body_dtype = b200_body_dtype
        self.b200_return_features = b200_return_features
        self.b200_layer_idx = b200_layer_idx
        self.b200_mlp_dtype = b200_mlp_dtype
        self._b200_init()
