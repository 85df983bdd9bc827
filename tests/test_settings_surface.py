"""schwz::Settings / schwz::Metadata of the host layer against include/settings.hpp of the reference
(only where /root/reference is mounted): every plain member of the reference exists here with the
same type, and every default the reference states is the default here.  (Members the reference
leaves uninitialised - most of Metadata - get defined values here.)"""
import os
import re

import pytest

from conftest import ROOT

REF = "/root/reference/include/settings.hpp"
OURS = os.path.join(ROOT, "schwarz-lib_b200", "host", "settings.hpp")
TYPES = r"bool|int|unsigned int|std::string|gko::int32|gko::size_type|ValueType|IndexType|double"


def members(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r"\b(%s)\s+([^;(){}]+);" % TYPES, text):
        for part in m.group(2).split(","):
            mm = re.match(r"^(\w+)\s*(?:=\s*(.+))?$", part.strip())
            if mm:
                out[mm.group(1)] = (m.group(1), (mm.group(2) or "").strip().rstrip("u"))
    return out


def test_every_reference_member_and_default_is_here():
    if not os.path.exists(REF):
        pytest.skip("/root/reference not present")
    ref, ours = members(open(REF).read()), members(open(OURS).read())
    assert len(ref) >= 55
    for name, (typ, default) in ref.items():
        assert name in ours, name
        assert ours[name][0] == typ, (name, ours[name])
        if default:
            assert ours[name][1] == default, (name, default, ours[name][1])
    # enumerator values (bench_ras and user code store them as integers)
    src = open(OURS).read()
    for enum in ("partition_regular = 0", "partition_metis = 1", "partition_zoltan = 2",
                 "partition_custom = 3", "partition_regular2d = 4", "direct_solver_cholmod = 0",
                 "direct_solver_ginkgo = 1", "iterative_solver_ginkgo = 2",
                 "iterative_solver_dealii = 3", "solver_custom = 4", "direct_solver_umfpack = 5"):
        assert enum in src, enum
    refsrc = open(REF).read()
    for name, val in re.findall(r"(partition_\w+|\w+_solver_\w+|solver_custom)\s*=\s*0x([0-9a-f])", refsrc):
        assert "%s = %d" % (name, int(val, 16)) in src, name
