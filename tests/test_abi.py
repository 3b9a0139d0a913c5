"""CPU-only: the C-ABI library builds, loads and exports every symbol include/egogen_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "egogen_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(eg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from egogen_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_python_binding_covers_header():
    from egogen_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _header_symbols()
    l = _lib.lib()
    assert l.eg_version() >= 100
    assert l.eg_launch_count() >= 0


def test_product_never_imports_oracle():
    """The product package must not route through the CPU oracle."""
    pkg = os.path.join(ROOT, "egogen_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
