"""The C-ABI library loads without a GPU and exports every symbol include/b200r.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200r.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200r_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_resolve():
    from robustart_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), "missing symbol " + n
    # the Python binding covers the whole header too
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().b200r_version() == 100


def test_argument_validation_without_gpu():
    """Pure host-side checks return B200R_EINVAL before any CUDA call."""
    from robustart_b200 import _lib
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    assert lib.b200r_corrupt_workspace_bytes(99, 1, 1, 224, 224, ctypes.byref(n)) == -1
    assert b"unknown corruption" in lib.b200r_last_error()
    assert lib.b200r_corrupt_workspace_bytes(0, 0, 1, 224, 224, ctypes.byref(n)) == -1
    assert lib.b200r_corrupt_workspace_bytes(11, 3, 256, 224, 224, ctypes.byref(n)) == 0 and n.value == 256 * 16
    assert lib.b200r_corrupt_ext_noise_count(0, 1, 2, 224, 224, ctypes.byref(n)) == 0 and n.value == 2 * 150528
    assert lib.b200r_corrupt_ext_noise_count(9, 1, 2, 224, 224, ctypes.byref(n)) == 0 and n.value == 2 * 65535


def test_product_never_imports_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/."""
    bad = []
    for base in ("robustart_b200", "RobustART", "prototype"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(dp, f)).read(), re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
