"""Shared test helpers: build an Engine from synthetic weights, compare tensors."""
from __future__ import annotations

from cover_vla_b200 import synthetic as S


def build_pi0_engine(d, w, max_R, max_K, **kw):
    return S.build_engine(d, w, None, None, max_R, max_K, **kw)


def build_full_engine(d, w, v, vw, max_R, max_K, **kw):
    """pi0 + verifier in one handle."""
    return S.build_engine(d, w, v, vw, max_R, max_K, **kw)


def rel_l2(x, y):
    x, y = x.float().cpu(), y.float().cpu()
    return ((x - y).norm() / (y.norm() + 1e-12)).item()


def max_abs(x, y):
    return (x.float().cpu() - y.float().cpu()).abs().max().item()
