"""Python handles on the operator-level C ABI (cvb_op_*).  Thin: pointer/stride marshalling only."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

EPI_STORE, EPI_GELU, EPI_RESID, EPI_GEGLU, EPI_F32 = range(5)


def _i64(v):
    return C.c_int64(int(v))


def gemm_bf16(a, w, *, epilogue=EPI_STORE, bias=None, resid=None, out=None, n_out=0, m_dev=None,
              force_bn=0):
    """out[M, N] = epilogue(a[M, K] @ w[N, K]^T).  a, w: bf16 CUDA, row-major (stride(-1) == 1)."""
    lib = _lib.load()
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and w.stride(-1) == 1
    M, K = a.shape
    N = w.shape[0]
    cols = n_out if epilogue == EPI_GEGLU else N
    if out is None:
        out = torch.empty(M, cols, device=a.device,
                          dtype=torch.float32 if epilogue == EPI_F32 else torch.bfloat16)
    rc = lib.cvb_op_gemm_bf16(
        _lib.ptr(a), _i64(a.stride(0)), _lib.ptr(w), _i64(w.stride(0)), M, N, K, int(epilogue),
        _lib.ptr(out), _i64(out.stride(0)), _lib.ptr(bias),
        int(bias is not None and bias.dtype == torch.float32),
        _lib.ptr(resid), int(resid is not None and resid.dtype == torch.float32),
        _i64(resid.stride(0) if resid is not None else 0), int(n_out), _lib.ptr(m_dev),
        int(force_bn), _lib.stream_ptr())
    _lib.check(rc)
    return out
