"""Host-only pieces of the C ABI (no GPU needed): symbol table, denoise schedule, time embedding."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from oracle import pi0_oracle as O

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from cover_vla_b200 import build, _lib
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = (ROOT / "include" / "coverb200.h").read_text()
    names = re.findall(r"CVB_API\s+[\w\s\*]+?\b(cvb_\w+)\s*\(", hdr)
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/coverb200.h but not exported"
    assert lib.cvb_abi_version() == 3


def test_denoise_schedule_matches_reference_loop(lib):
    for steps in (10, 5, 7, 16):
        buf = (C.c_float * 64)()
        dt = C.c_float()
        n = lib.cvb_denoise_times_host(steps, buf, 64, C.byref(dt))
        times, dt_ref = O.denoise_times(steps)
        assert n == len(times)
        assert list(buf[:n]) == times
        assert dt.value == dt_ref


def test_time_embedding_bit_exact(lib):
    lib.cvb_time_embedding_host.argtypes = [C.c_float, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_uint16)]
    lib.cvb_time_embedding_host.restype = None
    times, _ = O.denoise_times(10)
    for dim in (1024, 64, 256):
        for t in times:
            ref = O.sinusoidal_time_embedding(torch.tensor([t], dtype=torch.float32), dim).to(torch.bfloat16)[0]
            out = (C.c_uint16 * dim)()
            lib.cvb_time_embedding_host(t, dim, 4e-3, 4.0, out)
            got = torch.tensor(list(out), dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
            mism = (got.view(torch.int16) != ref.view(torch.int16)).sum().item()
            # allow a 1-ulp double-precision libm difference to flip at most a couple of bf16 roundings
            assert mism <= 2, (dim, t, mism)
            assert (got.float() - ref.float()).abs().max().item() <= 2 ** -7


def test_product_path_never_touches_the_oracle_or_the_reference_tree():
    """The oracle is test infrastructure: nothing under cover_vla_b200/ may import it or read /root/reference, bench.py may
    only reach it from its CPU legs (cpu_baseline / --impl reference), and the product must fail loudly - not fall back -
    when the CUDA library is missing."""
    import ast
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    for f in sorted((root / "cover_vla_b200").rglob("*.py")):
        src = f.read_text()
        assert "/root/reference" not in src, f
        for node in ast.walk(ast.parse(src)):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), (f, mods)
    # bench.py: every oracle import sits inside a CPU-leg function
    tree = ast.parse((root / "bench.py").read_text())
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        uses = any((isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle") or
                   (isinstance(n, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in n.names)) for n in ast.walk(fn))
        if uses:
            assert fn.name in ("_cpu_setup", "cpu_decision", "cpu_baseline_sample", "run_reference"), fn.name
    assert not any(isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n) for n in tree.body)
    # no silent fallback: a missing library is an error
    from cover_vla_b200 import _lib
    saved_path, saved_handle = _lib.LIB_PATH, _lib._lib
    try:
        _lib.LIB_PATH, _lib._lib = root / "cover_vla_b200" / "does_not_exist.so", None
        with pytest.raises(_lib.CvbError):
            _lib.load()
    finally:
        _lib.LIB_PATH, _lib._lib = saved_path, saved_handle
