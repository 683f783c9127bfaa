"""The boundary is a C ABI: include/coverb200.h must be valid C (not only C++), and a host written in plain C - no torch,
no Python - must be able to bind the library.  tests/c_abi/host_check.c is compiled with -std=c11 -pedantic -Werror and
run; it needs no GPU (version, error reporting, configuration guards, the weight manifest, the denoise constants)."""
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def libdir():
    from cover_vla_b200 import build, _lib
    build.build()
    assert Path(_lib.LIB_PATH).exists()
    return Path(_lib.LIB_PATH).parent


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_header_is_valid_c_and_links_from_c(libdir, tmp_path):
    exe = tmp_path / "host_check"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{ROOT / 'include'}",
           str(ROOT / "tests" / "c_abi" / "host_check.c"), "-o", str(exe), f"-L{libdir}", "-l:libcoverb200.so",
           f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ABI 3 OK" in r.stdout
    assert "missing weight" in r.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="no C++ compiler")
def test_header_is_valid_cpp_too(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "coverb200.h"\nint main() { return CVB_ABI_VERSION == 3 ? 0 : 1; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", f"-I{ROOT / 'include'}", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
