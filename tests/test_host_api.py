"""The plugin API of the reference: code written against its headers - subclass
schwz::SolverRAS, override the five virtual extension points with the reference's signatures,
construct / initialize() / run(), read the public data members - compiles unchanged against
schwarz-lib_b200/host/ for all four (ValueType, IndexType, MixedValueType) instantiations the
reference provides.  Compile-only: running needs a GPU (tests/test_gpu_bench_ras.py)."""
import os
import subprocess

from conftest import ROOT


def test_reference_style_plugin_code_compiles_against_the_host_headers():
    src = os.path.join(ROOT, "tests", "host_api", "plugin_check.cpp")
    inc = os.path.join(ROOT, "schwarz-lib_b200", "host")
    p = subprocess.run(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-I", inc, src],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[:4000]
