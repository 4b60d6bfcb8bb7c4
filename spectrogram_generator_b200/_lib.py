"""ctypes binding of ``libb200stft.so`` (C ABI: ``include/b2s.h``).

The library is built in-tree by :func:`build` (``nvcc`` for sm_100a) and loaded
from this directory.  There is no fallback: if the library is missing or a
compute entry fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libb200stft.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(_ROOT, "include", "b2s.h")

B2S_OK, B2S_ERR_BAD_ARG, B2S_ERR_UNSUPPORTED, B2S_ERR_CUDA, B2S_ERR_TIMEOUT = 0, -1, -2, -3, -4

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
BUILD_DIR = os.path.join(_HERE, "build")


class B2SError(RuntimeError):
    """A libb200stft entry returned an error code."""


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a (cross-compiles without a GPU) and link libb200stft.so.
    The kernel instantiations are spread over several translation units (b2s_inst_*.cu), which
    are compiled in parallel."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise B2SError("nvcc not found: cannot build libb200stft.so")
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(BUILD_DIR, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        res = subprocess.run([nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj], capture_output=True, text=True)
        if res.returncode != 0:
            raise B2SError(f"nvcc failed on {os.path.basename(src)}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, srcs))
    if verbose:
        for _, err in results:
            print(err)
    res = subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + [o for o, _ in results] + ["-lcudart"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise B2SError("link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None

_STFT_ARGS = [c_void_p, c_longlong, c_longlong, c_longlong, c_int, c_int, c_void_p, c_int, c_double,
              c_int, c_float, c_int, c_int, c_longlong, c_longlong, c_void_p, c_longlong, c_void_p]

_BAND_ARGS = [c_void_p, c_longlong, c_longlong, c_longlong, c_int, c_int, c_void_p, c_int, c_double,
              c_int, c_int, c_longlong, c_longlong, c_void_p, c_longlong, c_void_p]

_SUM_ARGS = [c_void_p, c_longlong, c_longlong, c_longlong, c_int, c_int, c_void_p, c_int, c_double,
             c_longlong, c_longlong, c_void_p, c_longlong, c_void_p, c_float, c_void_p, c_void_p]

# name -> (restype, argtypes); must list every symbol include/b2s.h declares
SIGNATURES = {
    "b2s_version": (c_int, []),
    "b2s_set_reserved_sms": (c_int, [c_int]),
    "b2s_set_option": (c_int, [c_char_p, c_int]),
    "b2s_peer_allreduce_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_uint, c_longlong, c_void_p, c_float,
                                       c_void_p]),
    "b2s_peer_allreduce_ex_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_uint, c_longlong, c_void_p, c_float,
                                          c_int, c_void_p]),
    "b2s_peer_allreduce_status": (c_int, [c_void_p]),
    "b2s_last_error": (c_char_p, []),
    "b2s_last_kernel": (c_char_p, []),
    "b2s_nperseg_support": (c_int, [c_int]),
    "b2s_frame_count": (c_longlong, [c_longlong, c_int, c_int]),
    "b2s_stft_psd_f32": (c_int, _STFT_ARGS),
    "b2s_stft_psd_f64": (c_int, _STFT_ARGS),
    "b2s_stft_band_power_f32": (c_int, _BAND_ARGS),
    "b2s_stft_band_power_f64": (c_int, _BAND_ARGS),
    "b2s_stft_psd_sum_scratch_elems": (c_longlong, [c_longlong, c_longlong]),
    "b2s_stft_psd_sum_f32": (c_int, _SUM_ARGS),
    "b2s_stft_psd_sum_f64": (c_int, _SUM_ARGS),
    "b2s_batch_sum_scratch_elems": (c_longlong, [c_longlong, c_longlong]),
    "b2s_batch_sum_f32": (c_int, [c_void_p, c_longlong, c_longlong, c_longlong, c_void_p, c_void_p,
                                  c_float, c_void_p]),
    "b2s_band_sums_scratch_elems": (c_longlong, []),
    "b2s_band_sums_f32": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b2s_display_scale_f32": (c_int, [c_void_p, c_longlong, c_int, c_float, c_void_p, c_void_p, c_void_p]),
}


def load():
    """Load the shared library (building is the job of __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2SError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the spectrogram path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def set_option(name: str, value: int):
    """Diagnostic switch of the kernel selection (include/b2s.h: b2s_set_option)."""
    check(load().b2s_set_option(name.encode(), int(value)), "b2s_set_option")


def last_kernel() -> str:
    return load().b2s_last_kernel().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc == B2S_OK:
        return
    msg = load().b2s_last_error().decode("utf-8", "replace")
    if rc == B2S_ERR_BAD_ARG:
        raise ValueError(f"{what}: {msg}")
    if rc == B2S_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == B2S_ERR_TIMEOUT:
        raise TimeoutError(f"{what}: {msg}")
    raise B2SError(f"{what}: {msg} (code {rc})")
