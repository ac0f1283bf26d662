"""ctypes loader for ``csrc/libcoda_b200.so`` -- the only native code of the package.

There is deliberately NO fallback: if the library is missing or a symbol the
headers declare is absent, importing an op raises.  Calls pass raw device
pointers (``tensor.data_ptr()``) and the current CUDA stream handle; the C-ABI is
declared in ``include/*.h``.
"""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "csrc" / "libcoda_b200.so"
INCLUDE_DIR = _PKG.parent / "include"

_lib = None
LAUNCHES = 0  # C-ABI kernel-launching calls made so far (bench.py reports the delta)


class CodaError(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every ``coda_*`` function the public headers declare."""
    names: list[str] = []
    for h in sorted(INCLUDE_DIR.glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names += re.findall(r"\b(coda_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise CodaError(
                f"{LIB_PATH} is missing: build it with `python -m coda_neurips2023_b200.build` "
                "(there is no CPU or PyTorch fallback for these ops)"
            )
        _lib = ctypes.CDLL(str(LIB_PATH))
        _lib.coda_status_string.restype = ctypes.c_char_p
        _lib.coda_status_string.argtypes = [ctypes.c_int]
        if hasattr(_lib, "coda_layer_norm_bwd_scratch"):
            _lib.coda_layer_norm_bwd_scratch.restype = ctypes.c_longlong
        missing = [s for s in declared_symbols() if not hasattr(_lib, s)]
        if missing:
            raise CodaError(f"{LIB_PATH} does not export: {missing}")
    return _lib


def check(status: int, what: str) -> None:
    global LAUNCHES
    LAUNCHES += 1
    if status != 0:
        msg = lib().coda_status_string(int(status))
        raise CodaError(f"{what} failed: {msg.decode() if msg else status} (status {status})")


def ptr(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def stream_of(t) -> ctypes.c_void_p:
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
