"""no-GPU tier: libmidas_b200.so builds for sm_100a, loads, and exports every function that
include/midas_b200.h declares (no compute calls here); the ctypes table covers the same set."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    src = open(os.path.join(ROOT, "include", "midas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from midastouch_b200 import _lib

    path = _lib.build()
    lib = ctypes.CDLL(path)
    names = declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    lib.mt_version.restype = ctypes.c_int
    assert lib.mt_version() >= 100


def test_sass_is_sm100a():
    import subprocess
    from midastouch_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "midastouch_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
