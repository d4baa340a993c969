"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/kws_b200.h
declares, and rejects bad arguments before touching the device (no compute calls without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kws_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kws_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("kws_abi_version", "kws_last_error", "kws_normalize_rows", "kws_mlp", "kws_temporal", "kws_sim",
              "kws_stem", "kws_sim_stem", "kws_maxpool_nhwc", "kws_scores", "kws_topk"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/kws_b200.h but not exported"


def test_python_binding_mirrors_header(built_lib):
    from enhance_cb_whisper_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.kws_abi_version() == _lib.ABI_VERSION


def test_library_is_sm100a_tcgen05(built_lib):
    """SASS evidence of the Blackwell-native path: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA)."""
    out = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnem in out.stdout, f"{mnem} not found in SASS"


def test_bad_arguments_are_rejected_without_a_device(built_lib):
    from enhance_cb_whisper_b200 import _lib

    lib = _lib.load()
    # null pointers
    assert lib.kws_sim(None, None, 1, 1, 1, 8, 8, 64, 0, None, None, 8, None) == -1
    assert b"null" in lib.kws_last_error()
    # Dk not a multiple of 64
    one = C.c_void_p(16)
    assert lib.kws_sim(one, one, 1, 1, 1, 8, 8, 48, 0, one, None, 8, None) == -1
    assert b"Dk" in lib.kws_last_error()
    # DIAG pairing needs U == K
    assert lib.kws_sim(one, one, 1, 2, 3, 8, 8, 64, 1, one, None, 8, None) == -1
    # MLP shape rules
    assert lib.kws_mlp(one, 1, 1, 1, 100, 50, 64, 0, one, one, one, one, one, None, 1e-6, 0, one, None) == -1
    assert b"multiple of 64" in lib.kws_last_error()
    # max-pool: channel chunks of 16 bytes
    assert lib.kws_maxpool_nhwc(one, 1, 4, 4, 12, one, None) == -1
    assert b"multiple of 8" in lib.kws_last_error()
    # top-k limits
    assert lib.kws_topk(one, None, 10, 1, 0, 2000, one, one, None, None) == -1
    assert lib.kws_topk(one, None, 5000, 1, 0, 10, one, one, None, None) == -1  # more than one segment: workspace
    assert lib.kws_topk_workspace_bytes(100, 4, 10) == 0 and lib.kws_topk_workspace_bytes(100000, 512, 200) > 0
    # fused similarity + stem (+ pool): pooled output only through kws_sim_stem_pool; > 12 layers need the workspace
    args = [one, one, None, 12, 2, 2, 16, 128, 64, 0, 0, 2, 0, 2, one, one]
    assert lib.kws_sim_stem_ragged(*args, 2, one, None) == -1  # KWS_STEM_OUT_POOL_NHWC_BF16 is not an out_mode here
    assert b"kws_sim_stem_pool" in lib.kws_last_error()
    args32 = list(args)
    args32[3] = 32
    assert lib.kws_sim_stem_pool(*args32, one, None, None) == -1
    assert b"workspace" in lib.kws_last_error()
    assert lib.kws_sim_stem_pool_workspace_bytes(12, 1000, 150, 1500) == 0
    assert lib.kws_sim_stem_pool_workspace_bytes(32, 1000, 75, 750) == 1000 * 38 * 375 * 64 * 2
    assert lib.kws_sim_stem_ragged(one, one, one, *args[3:9], 2, *args[10:], 1, one, None) == -1  # keyword lengths x KWS_PAIRS_PER_KEYWORD
    assert b"length table" in lib.kws_last_error()
    # fused projector: shapes it does not cover are reported, not mis-computed
    assert lib.kws_mlp_fused_supported(768, 384, 64) == 1 and lib.kws_mlp_fused_supported(1280, 640, 64) == 1
    assert lib.kws_mlp_fused_supported(100, 50, 64) == 0 and lib.kws_mlp_fused_supported(768, 384, 24) == 0
    with pytest.raises(_lib.KWSError):
        _lib.check(-1, "kws_topk")


def test_cpu_tensors_are_refused(built_lib):
    import torch

    from enhance_cb_whisper_b200 import ops

    x = torch.zeros(1, 1, 4, 64)
    with pytest.raises(ops.KWSError):
        ops.normalize_rows(x, [0], None)


def test_missing_library_fails_loudly(built_lib, monkeypatch):
    from enhance_cb_whisper_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libkws_b200.so")
    with pytest.raises(_lib.KWSError):
        _lib.load()
