"""CPU: the C-ABI library loads and exports exactly what include/mar.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from multimodalaggressionrecognition_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mar.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"^\s*(?:const\s+)?(?:int64_t|int|void|char\*|const char\*)\s*\**\s*(mar_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.M | re.S)
    return {name: args for name, args in decls}


def test_header_declares_functions():
    fns = header_functions()
    assert len(fns) >= 25
    for must in ("mar_linear_fwd", "mar_attention_fwd", "mar_gru_fwd", "mar_layernorm_fwd", "mar_cross_entropy_fwd"):
        assert must in fns


def test_library_exports_every_header_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"libmar.so does not export {name}"


def test_python_prototypes_match_header():
    fns = header_functions()
    assert set(fns) == set(_lib.PROTOTYPES), set(fns) ^ set(_lib.PROTOTYPES)
    for name, args in fns.items():
        args = args.strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        assert n == len(_lib.PROTOTYPES[name][1]), f"{name}: header has {n} args, binding has {len(_lib.PROTOTYPES[name][1])}"


def test_load_and_version():
    lib = _lib.load()
    assert lib.mar_version() == 101
    assert lib.mar_launch_count() == 0 or lib.mar_launch_count() > 0
    assert isinstance(_lib.last_error(), str)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmar.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()
