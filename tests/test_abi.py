"""The C-ABI library loads and exports every symbol include/schwz_b200.h
declares; error convention.  No compute calls (no GPU needed)."""
import ctypes
import re


def test_every_declared_symbol_is_exported(sz):
    text = open(sz.HEADER_PATH).read()
    names = sorted(set(re.findall(r"\b(schwz_b200_[a-z0-9_]+)\s*\(", text)))
    assert len(names) > 80
    lib = sz.load()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_torch_or_cxx_types_in_the_header(sz):
    text = open(sz.HEADER_PATH).read()
    for bad in ("torch", "at::", "std::", "Tensor", "template"):
        assert bad not in text


def test_error_convention(sz):
    lib = sz.load()
    n = ctypes.c_int32(0)
    nnz = ctypes.c_int64(0)
    p = ctypes.c_void_p()
    rc = lib.schwz_b200_read_mtx(b"/nonexistent/file.mtx", ctypes.byref(n), ctypes.byref(nnz),
                                 ctypes.byref(p), ctypes.byref(p), ctypes.byref(p))
    assert rc != 0
    assert b"Could not find the file" in lib.schwz_b200_last_error()


def test_device_entry_points_fail_loudly_without_a_gpu(sz):
    if sz.device_count() > 0:
        return
    import pytest
    with pytest.raises(sz.SchwzError):
        sz.Context(0)
