"""The C-ABI shared library loads without a GPU and exports every symbol include/mscl_b200.h declares,
with the argument counts the ctypes binding (mscl_b200/_cabi.py) assumes.  No compute call is made."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mscl_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|const char \*)\s*(mscl_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return decls


@pytest.fixture(scope="module")
def lib():
    from mscl_b200 import _cabi, build
    build.build()                      # no-op when the library is newer than its sources
    return ctypes.CDLL(_cabi.LIB_PATH)


def test_header_declares_the_documented_entry_points():
    d = _declared()
    assert len(d) == 37, sorted(d)
    for name in ("mscl_enqueue", "mscl_ema_multi", "mscl_fra_apply", "mscl_lmcl", "mscl_infonce_partial", "mscl_gather_rows"):
        assert name in d


def test_library_exports_every_declared_symbol(lib):
    for name in _declared():
        assert getattr(lib, name) is not None, name


def test_ctypes_prototypes_match_header():
    from mscl_b200 import _cabi
    d = _declared()
    assert set(_cabi.EXPORTS) == set(d)
    for name, argtypes in _cabi.PROTOTYPES.items():
        assert len(argtypes) == d[name], (name, len(argtypes), d[name])


def test_version_and_error_text_without_a_gpu(lib):
    lib.mscl_abi_version.restype = ctypes.c_int
    lib.mscl_last_error.restype = ctypes.c_char_p
    assert lib.mscl_abi_version() == 1
    # argument validation happens before any CUDA call: a null pointer is rejected with a message
    lib.mscl_gather_rows.restype = ctypes.c_int
    rc = lib.mscl_gather_rows(None, None, None, ctypes.c_int32(1), ctypes.c_int64(4), None)
    assert rc == -1 and b"null" in lib.mscl_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from mscl_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.MsclError, match="no CPU fallback"):
        _cabi.load()


def test_launch_counts_are_integers_for_known_entry_points():
    """`gpu_launches` of bench.py is summed from this table: every key is a declared entry point, every value an int."""
    from mscl_b200 import _cabi
    for name, n in _cabi._LAUNCHES_PER_CALL.items():
        assert name in _cabi.PROTOTYPES and isinstance(n, int) and 0 <= n <= 4, (name, n)
