"""ctypes binding of libcoverb200.so (the C ABI in include/coverb200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libcoverb200.so"

_lib = None


class CvbError(RuntimeError):
    pass


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CvbError(
            f"{LIB_PATH} is missing - build it with `python -m cover_vla_b200.build` "
            "(there is no CPU / PyTorch fallback for this path)"
        )
    lib = C.CDLL(str(LIB_PATH))
    lib.cvb_last_error.restype = C.c_char_p
    lib.cvb_abi_version.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().cvb_last_error()
        raise CvbError(f"coverb200 call failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> C.c_void_p:
    """Device (or host) pointer of a torch tensor, or NULL."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr(stream=None) -> C.c_void_p:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
