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
    assert len(d) == 46, sorted(d)
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


def test_argument_validation_precedes_any_cuda_call():
    """Error convention of the boundary: bad shapes / alignments / null pointers come back as
    MSCL_EINVAL (-1) with a message, before anything is launched -- so this runs on a machine without a GPU."""
    from mscl_b200 import _cabi
    lib = _cabi.load()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)                  # a valid (host) address: only ever inspected, never dereferenced
    odd = ctypes.c_void_p(p.value + 4)                     # 4-byte aligned only
    cases = [
        ("mscl_hw_mean_ndhwc_fwd", (p, p, 2, 48, 2, 9, None), b"multiple of 32"),            # C % 32 != 0
        ("mscl_hw_mean_ndhwc_fwd", (p, p, 70000, 32, 1, 9, None), b"grid"),                   # N*T beyond grid.y
        ("mscl_hw_mean_fwd", (None, p, 4, 9, None), b"null"),
        ("mscl_upsample_trilinear_ndhwc_fwd", (p, p, 2, 6, 1, 2, 2, 2, 4, 4, None), b"C % 4"),
        ("mscl_upsample_trilinear_fwd", (p, p, 0, 1, 2, 2, 2, 4, 4, None), b"bad shape"),
        ("mscl_upsample_trilinear_fwd", (odd, p, 1, 1, 2, 2, 2, 4, 4, None), b"aligned"),
        ("mscl_linear_axis_bwd", (p, p, 1, 0, 4, 8, None), b"bad shape"),
        ("mscl_lmcl", (p, p, 2, 128, 4, 2, 1.0, p, p, p, p, None), b"t2"),                    # fewer flow frames than RGB frames
        ("mscl_retrieval_rank", (p, 3, p, p, 2, 5, p, None), b"ld"),                           # row pitch shorter than n_train
        ("mscl_center_normalize", (p, 4, 8, p, 0, p, p, None), b"n_chunks"),
        ("mscl_infonce_partial", (p, 0, p, p, 1024, 0, p, 1, 1, None), b"bad M"),
        ("mscl_infonce_partial", (p, 8, p, p, 1024, 0, p, 99, 1, None), b"n_part"),           # more partial slabs than 128-key tiles
        ("mscl_flow_visualize", (None, None, None, p, 1, 1, 4, 4, None), None),
    ]
    for name, args, needle in cases:
        rc = getattr(lib, name)(*args)
        msg = lib.mscl_last_error() or b""
        assert rc == -1, (name, rc, msg)
        if needle is not None:
            assert needle in msg, (name, msg)
    with pytest.raises(_cabi.MsclError):                   # the binding turns the code into an exception carrying the text
        _cabi.call("mscl_hw_mean_ndhwc_fwd", p, p, 2, 48, 2, 9, None)
