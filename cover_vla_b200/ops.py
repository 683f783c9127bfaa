"""Python handles on the operator-level C ABI (cvb_op_*).  Thin: pointer/stride marshalling only."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

EPI_STORE, EPI_GELU, EPI_RESID, EPI_GEGLU, EPI_F32, EPI_GEGLU64, EPI_PARTIAL = range(7)
SKINNY_AUTO = -100  # force_bn value: skinny (swap-AB, cluster split-K) GEMM with the split heuristic; -S forces S splits


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
    cols = n_out if epilogue in (EPI_GEGLU, EPI_GEGLU64) else N
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


def split3_f32(x, weight_layout=False, relu=False):
    """fp32 [rows, K] -> bf16 [rows, 3K]: [hi | hi | lo] (activations) or [hi | lo | hi] (weights), hi = bf16(x), lo = bf16(x - hi).
    gemm_bf16(split3(a), split3(w, weight_layout=True), epilogue=EPI_F32) is an fp32-accurate a @ w^T on the tensor cores."""
    lib = _lib.load()
    assert x.dtype == torch.float32 and x.is_cuda and x.stride(-1) == 1 and x.dim() == 2
    rows, K = x.shape
    out = torch.empty(rows, 3 * K, device=x.device, dtype=torch.bfloat16)
    lib.cvb_op_split3_f32.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
    _lib.check(lib.cvb_op_split3_f32(_lib.ptr(x), x.stride(0), _lib.ptr(out), rows, K, int(weight_layout), int(relu),
                                     _lib.stream_ptr()))
    return out


def gemm_splitk_partial(a, w, splits):
    """fp32 partial products [S, M, N] of a[M, K] @ w[N, K]^T, K split into S balanced runs of 64-wide blocks."""
    M, K = a.shape
    N = w.shape[0]
    S = max(1, min(int(splits), (K + 63) // 64))
    out = torch.empty(S, M, N, device=a.device, dtype=torch.float32)
    lib = _lib.load()
    rc = lib.cvb_op_gemm_bf16(
        _lib.ptr(a), _i64(a.stride(0)), _lib.ptr(w), _i64(w.stride(0)), M, N, K, EPI_PARTIAL,
        _lib.ptr(out), _i64(N), None, 0, None, 0, _i64(0), 0, None, -S, _lib.stream_ptr())
    _lib.check(rc)
    return out


def rmsnorm(x, w, eps=1e-6):
    """y = bf16(x * rsqrt(mean(x^2) + eps) * (1 + w)): GemmaRMSNorm; x bf16 or fp32 [rows, width], w bf16 or fp32."""
    lib = _lib.load()
    M, N = x.shape
    assert x.stride(1) == 1 and w.is_contiguous()
    y = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    lib.cvb_op_rmsnorm.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                   C.c_int, C.c_float, C.c_void_p]
    _lib.check(lib.cvb_op_rmsnorm(_lib.ptr(x), int(x.dtype == torch.float32), x.stride(0), _lib.ptr(w),
                                  int(w.dtype == torch.float32), _lib.ptr(y), N, M, N, float(eps), _lib.stream_ptr()))
    return y


def rmsnorm_reduce(partials, resid, w, eps=1e-6):
    """(h, y): h = bf16(bf16(sum_s partials[s]) + resid), y = GemmaRMSNorm(h) with weight w (bf16 or fp32)."""
    lib = _lib.load()
    S, M, N = partials.shape
    assert partials.dtype == torch.float32 and partials.is_contiguous()
    h = torch.empty(M, N, device=partials.device, dtype=torch.bfloat16)
    y = torch.empty_like(h)
    lib.cvb_op_rmsnorm_reduce.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int64,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_int, C.c_int, C.c_float, C.c_void_p]
    rc = lib.cvb_op_rmsnorm_reduce(_lib.ptr(partials), S, M * N, N, _lib.ptr(resid),
                                   int(resid.dtype == torch.float32), resid.stride(0), _lib.ptr(w),
                                   int(w.dtype == torch.float32), _lib.ptr(h), N, _lib.ptr(y), N, M, N, float(eps),
                                   _lib.stream_ptr())
    _lib.check(rc)
    return h, y


def layernorm_reduce(partials, bias, resid, w, b, eps=1e-6):
    """(h, y): h = bf16(bf16(sum_s partials[s] + bias) + resid), y = LayerNorm(h) * w + b (all bf16 but the partials)."""
    lib = _lib.load()
    S, M, N = partials.shape
    assert partials.dtype == torch.float32 and partials.is_contiguous()
    h = torch.empty(M, N, device=partials.device, dtype=torch.bfloat16)
    y = torch.empty_like(h)
    lib.cvb_op_layernorm_reduce.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                            C.c_int, C.c_int, C.c_float, C.c_void_p]
    rc = lib.cvb_op_layernorm_reduce(_lib.ptr(partials), S, M * N, N, _lib.ptr(bias), _lib.ptr(resid), resid.stride(0),
                                     _lib.ptr(w), _lib.ptr(b), _lib.ptr(h), N, _lib.ptr(y), N, M, N, float(eps),
                                     _lib.stream_ptr())
    _lib.check(rc)
    return h, y


def attention(q, k0, v0, *, heads, kv_heads, head_dim, kv0_len=None, kv0_len_dev=None, q_per_kv_batch=1,
              k1=None, v1=None, suffix_mask=False, scale=None, force_two_pass=False, rope=None, algo=0, vt0=None):
    """q [B, Tq, heads*hd]; k0/v0 [Bkv, T0, kv_heads*hd]; optional k1/v1 [B, T1, kv_heads*hd] (bf16, CUDA)."""
    lib = _lib.load()
    B, Tq, _ = q.shape
    out = torch.empty(B, Tq, heads * head_dim, device=q.device, dtype=torch.bfloat16)
    scale = float(scale if scale is not None else head_dim ** -0.5)
    T0 = k0.shape[1]
    lib.cvb_op_attention_tc.argtypes = ([C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                         C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64] + [C.c_int] * 5 +
                                        [C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p])
    rc = lib.cvb_op_attention_tc(
        _lib.ptr(q), q.stride(0), q.stride(1), _lib.ptr(k0), _lib.ptr(v0), k0.stride(0), k0.stride(1),
        _lib.ptr(kv0_len_dev), int(kv0_len if kv0_len is not None else T0), T0, q_per_kv_batch,
        _lib.ptr(k1), _lib.ptr(v1), k1.stride(0) if k1 is not None else 0, k1.stride(1) if k1 is not None else 0,
        k1.shape[1] if k1 is not None else 0, int(suffix_mask), _lib.ptr(out), out.stride(0), out.stride(1),
        B, heads, kv_heads, Tq, head_dim, scale, (int(algo) + 1 if algo else int(force_two_pass)), _lib.ptr(rope),
        _lib.ptr(vt0), (vt0.stride(1) if vt0 is not None else 0), _lib.stream_ptr())
    _lib.check(rc)
    return out


def transpose_values(v, pad_to=64):
    """[B, T, hd] -> V^T [B, hd, round_up(T, pad_to)] (zero padded): the layout the tcgen05 attention kernels read."""
    B, T, hd = v.shape
    vt = torch.zeros(B, hd, (T + pad_to - 1) // pad_to * pad_to, device=v.device, dtype=v.dtype)
    vt[:, :, :T] = v.transpose(1, 2)
    return vt


def attention_umma(q, k, v, *, lens=None, kmax=None, scale=None):
    """tcgen05 prefix attention: q [B, Tq, 8*256], k / v [B, Tk, 256] (bf16, CUDA); lens int32 [B] valid keys."""
    lib = _lib.load()
    B, Tq, qw = q.shape
    Tk = k.shape[1]
    hd, heads = 256, qw // 256
    assert q.is_contiguous() and k.is_contiguous()
    vt_ld = (Tk + 63) // 64 * 64
    vt = torch.zeros(B, hd, vt_ld, device=q.device, dtype=torch.bfloat16)
    vt[:, :, :Tk] = v.transpose(1, 2)
    out = torch.empty(B, Tq, heads * hd, device=q.device, dtype=torch.bfloat16)
    lib.cvb_op_attention_umma.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                          C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
    rc = lib.cvb_op_attention_umma(_lib.ptr(q), qw, B * Tq, Tq, _lib.ptr(k), B * Tk, Tk, _lib.ptr(vt), vt_ld,
                                   _lib.ptr(lens), Tk, int(kmax if kmax is not None else Tk), _lib.ptr(out),
                                   out.stride(0), out.stride(1), B, Tq, heads, hd,
                                   float(scale if scale is not None else hd ** -0.5), _lib.stream_ptr())
    _lib.check(rc)
    return out


SACT_NONE, SACT_RELU, SACT_GELU_ERF, SACT_SILU = range(4)


def sgemm_f32(a, w, *, bias=None, row_bias=None, resid=None, act=SACT_NONE, out=None):
    """out[M, N] = act(a[M, K] @ w[N, K]^T + bias + row_bias) + resid, all float32 CUDA (true-fp32 accumulation)."""
    lib = _lib.load()
    assert a.dtype == torch.float32 and w.dtype == torch.float32
    assert a.stride(-1) == 1 and w.stride(-1) == 1
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    lib.cvb_op_sgemm_f32.argtypes = ([C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                      C.c_void_p])
    rc = lib.cvb_op_sgemm_f32(_lib.ptr(a), _i64(a.stride(0)), _lib.ptr(w), _i64(w.stride(0)), M, N, K, _lib.ptr(out),
                              _i64(out.stride(0)), _lib.ptr(bias), _lib.ptr(row_bias), _lib.ptr(resid),
                              _i64(resid.stride(0) if resid is not None else 0), int(act), _lib.stream_ptr())
    _lib.check(rc)
    return out
