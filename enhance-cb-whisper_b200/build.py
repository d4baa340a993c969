"""Build libkws_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python enhance-cb-whisper_b200/build.py [--force] [--verbose] [--debug-hooks]

``--debug-hooks`` builds a second flavour, ``libkws_b200_dbg.so`` (-DKWS_DEBUG_HOOKS: exported ``kws_debug_*``
setters, KWS_FUSED_* environment knobs, optional role timers), for ``tools/`` and the variant tests; load it by
pointing ``KWS_B200_LIB`` at it.  The release library carries none of that.

Plain ``nvcc -shared``: the library has no ATen / libtorch / libcuda link
dependency (cudart is linked statically), so it cross-compiles on a box without
a GPU and is loaded with ctypes at run time.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libkws_b200.so")
LIB_DBG = os.path.join(HERE, "libkws_b200_dbg.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["kws_abi.cu", "kws_prep.cu", "kws_gemm.cu", "kws_mlp_fused.cu", "kws_temporal.cu", "kws_stem.cu", "kws_fused.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, debug_hooks: bool = False) -> str:
    obj_dir = OBJ_DIR + ("_dbg" if debug_hooks else "")
    lib = LIB_DBG if debug_hooks else LIB
    os.makedirs(obj_dir, exist_ok=True)
    flags = list(NVCC_FLAGS)
    if debug_hooks:
        flags.append("-DKWS_DEBUG_HOOKS")
        if os.environ.get("KWS_WAIT_HINT_NS"):  # tune the mbarrier suspend-time hint
            flags.append("-DKWS_WAIT_HINT_NS=" + os.environ["KWS_WAIT_HINT_NS"])
            force = True
        if os.environ.get("KWS_FUSED_TIMERS"):  # per-role cycle counters in the fused kernel
            flags.append("-DKWS_FUSED_TIMERS")
            force = True
    headers = [os.path.join(CSRC, "kws_common.cuh"), os.path.join(INCLUDE, "kws_b200.h"), os.path.abspath(__file__)]
    nvcc = _nvcc()

    def compile_one(src: str):
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + flags + ["-I", INCLUDE, "-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
            return o, True
        return o, False

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if force or any(ch for _, ch in results) or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-o", lib] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, debug_hooks="--debug-hooks" in sys.argv)
    print(path)
