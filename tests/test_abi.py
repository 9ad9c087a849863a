"""The C-ABI shared library: loads, exports every symbol include/gjkepa.h declares, and refuses to
work without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    txt = (ROOT / "include" / "gjkepa.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gjkepa_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from collision_detect_gjk_epa_b200 import _lib
    _lib.build()
    L = C.CDLL(str(_lib.SO_PATH))
    names = header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.SYMBOLS) == names


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from collision_detect_gjk_epa_b200 import Engine, GjkEpaError, GJKEPA, workloads, _lib
    with pytest.raises(GjkEpaError):
        Engine(0)
    b = workloads.unit_cubes()
    with pytest.raises(GjkEpaError):
        GJKEPA(1, 1.0, b.hull(0), b.hull(1))
    h = C.c_void_p()
    assert _lib.lib().gjkepa_create(C.byref(h), 0) == _lib.ERR_CUDA and not h


def test_product_does_not_touch_the_oracle():
    # the product path may not import, link or execute anything under oracle/
    pkg = ROOT / "collision_detect_gjk_epa_b200"
    bad = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|gjkepa_oracle|orc_gjkepa|liborc)")
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("Makefile")):
        assert not bad.search(p.read_text()), p
    # developer tools outside tests/ do not use it either (gen_dirtable.py only WRITES the direction table both sides include)
    imp = re.compile(r"(from\s+oracle|import\s+oracle)")
    for p in (ROOT / "tools").glob("*.py"):
        assert not imp.search(p.read_text()), p
    out = __import__("subprocess").run(["ldd", str(pkg / "libgjkepa_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def c_arg_counts():
    """name -> number of parameters of every function include/gjkepa.h declares."""
    txt = (ROOT / "include" / "gjkepa.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(gjkepa_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_fortran_shim_binds_only_declared_functions_with_matching_argument_counts():
    """No Fortran compiler exists here, so the ISO_C_BINDING shim is at least checked as text: every BIND(C, NAME=...) names
    a function of include/gjkepa.h, its dummy list has that function's number of parameters, no line exceeds the free-form
    limit of 132 columns, and every PUBLIC procedure is defined."""
    src = (ROOT / "fortran" / "gclib_gjkepa.f90").read_text()
    assert max(len(l) for l in src.splitlines()) <= 132
    counts = c_arg_counts()
    joined = re.sub(r"&\s*\n\s*&?", " ", src)                  # continuation lines
    binds = re.findall(r"FUNCTION\s+(\w+)\s*\(([^)]*)\)\s*BIND\(C,\s*NAME=\"(\w+)\"\)", joined)
    assert len(binds) >= 15
    for fname, dummies, cname in binds:
        assert fname == cname, (fname, cname)
        assert cname in counts, f"{cname} is not declared in include/gjkepa.h"
        n = len([d for d in dummies.split(",") if d.strip()])
        assert n == counts[cname], f"{cname}: {n} dummies in the shim, {counts[cname]} parameters in the header"
    pub = re.search(r"PUBLIC\s*::(.*?)\n\s*\n", joined, flags=re.S).group(1)
    for name in re.findall(r"\b(GJKEPA\w*)\b", pub):
        assert re.search(rf"(SUBROUTINE|FUNCTION)\s+{name}\b", src), f"PUBLIC {name} has no definition"


def test_header_and_example_are_plain_c99(tmp_path):
    """The boundary is a C ABI: include/gjkepa.h and the example that uses it must compile as C99, pedantically, without the
    C++ front end."""
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"), "-c",
                        str(ROOT / "examples" / "gjkepa_example.c"), "-o", str(tmp_path / "ex.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
