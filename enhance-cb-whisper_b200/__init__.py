"""B200-native (sm_100a) scoring path of Enhance-CB-Whisper's ``efficient_kws``:
per-layer compression of Whisper embeddings, layer-wise cosine-similarity
images and the ResNet stem, behind the reference's own module interface.

Import name: ``enhance_cb_whisper_b200`` (the directory is
``enhance-cb-whisper_b200``; ``enhance_cb_whisper_b200.py`` at the repository
root maps one onto the other).
"""
from ._lib import KWSError, LIB_PATH  # noqa: F401
from .model import KWSModelB200, KWSOutput, Resnet  # noqa: F401
from .engine import KWSEngine, PackedWeights, pack_weights  # noqa: F401
from . import bank, cbw  # noqa: F401

__all__ = ["KWSModelB200", "KWSOutput", "Resnet", "KWSEngine", "PackedWeights", "pack_weights", "KWSError", "bank", "cbw"]
