"""CPU-side checks of the C ABI: the library builds for sm_100a, loads, and exports every symbol that
include/sfk.h declares.  No compute calls (there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "sfk.h")).read()
    return sorted(set(re.findall(r"\b(sfk_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_loads_and_exports_header_symbols():
    from sfattack import build, lib
    path = build.build()
    assert os.path.exists(path)
    dll = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/sfk.h but not exported"
    assert sorted(lib.EXPORTS) == names
    assert dll.sfk_version() >= 100


def test_igemm_desc_layout_matches_header():
    from sfattack import lib
    # 2 pointers-with-int blocks ...: spot-check a few offsets against the C layout rules
    d = lib.SfkIgemmDesc
    assert d.a.offset == 0 and d.n_img.offset == 8 and d.b.offset == 32
    assert d.taps.offset == d.num_taps.offset + 4
    assert ctypes.sizeof(lib.SfkTap) == 20


def test_argument_errors_are_reported_without_a_gpu():
    from sfattack import build
    dll = ctypes.CDLL(build.build())
    dll.sfk_last_error_string.restype = ctypes.c_char_p
    rc = dll.sfk_igemm(None, None)
    assert rc == -1 and b"null" in dll.sfk_last_error_string()
    rc = dll.sfk_conv_c3_fwd(None, None, None, None, 1, 8, 8, 64, 1, None)
    assert rc == -1
