"""LightningCLI drop-in: the reference ``KWSModel`` with the B200 forward.

Usable where ``pytorch_lightning`` and the reference package ``efficient_kws``
are importable (i.e. inside a checkout of the reference, ``src/`` on the path).
Switch one line of any ``src/efficient_kws/configs/*.yaml``::

    model:
      class_path: enhance_cb_whisper_b200.lightning.KWSModelB200

``run_efficient_kws.py`` (``subclass_mode_model=True``, :50) instantiates it with
the unchanged ``init_args``.  Replaced: ``forward`` (B200 kernels) and, in eval
mode, ``test_step`` / ``validation_step`` (model.py:748-802, :304-385): all keyword
groups of a DataLoader item are stacked and scored in ONE pass, so the utterance is
compressed once instead of once per group of 50 keywords; what they append to
``test_step_outputs`` / ``validation_step_outputs`` is unchanged.  Every other hook
(epoch-end metrics, ``on_load_checkpoint`` legacy remap, optimisers, ``training_step``)
is inherited from the reference.  Covered by tests/test_lightning.py under the stub
modules of oracle/ref_stub.
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
                 b200_layer_idx=None, b200_mlp_dtype: str = "float16", b200_fused_pool: bool = True,
                 b200_ragged: bool = True, **kwargs):
        super().__init__(*args, **kwargs)
        self.b200_fused_pool = b200_fused_pool
        self.b200_ragged = b200_ragged
        self.b200_body_dtype = b200_body_dtype
        self.b200_return_features = b200_return_features
        self.b200_layer_idx = b200_layer_idx
        self.b200_mlp_dtype = b200_mlp_dtype
        self._b200_init()
