"""The oracle is test infrastructure: nothing under schwarz-lib_b200/ (the product) or include/ may
import, link, call or even name anything under oracle/, and the shared library must not depend on
it; loading the product fails loudly when the library has not been built."""
import os
import subprocess

import pytest

from conftest import ROOT

PRODUCT = os.path.join(ROOT, "schwarz-lib_b200")


def test_product_sources_never_mention_the_oracle():
    offenders = []
    for base in (PRODUCT, os.path.join(ROOT, "include")):
        for dirpath, dirnames, files in os.walk(base):
            dirnames[:] = [d for d in dirnames if d not in ("build", "lib", "bin", "__pycache__")]
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) and f != "Makefile":
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("oracle/", "libschwz_oracle", "import oracle", "schwz_oracle", "_ref/",
                               "libschwz_ref"):
                    if needle in text:
                        offenders.append((os.path.relpath(os.path.join(dirpath, f), ROOT), needle))
    assert not offenders, offenders


def test_shared_library_does_not_link_the_oracle(sz):
    out = subprocess.run(["ldd", sz.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "schwz_ref" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", sz.LIB_PATH], capture_output=True, text=True).stdout
    assert " orc_" not in syms and " ref_run" not in syms


def test_missing_library_fails_loudly(sz, monkeypatch):
    import schwz_b200
    monkeypatch.setattr(schwz_b200, "_lib", None)
    monkeypatch.setattr(schwz_b200, "LIB_PATH", os.path.join(ROOT, "no", "such", "libschwz_b200.so"))
    with pytest.raises(schwz_b200.SchwzError, match="not built"):
        schwz_b200.load()
