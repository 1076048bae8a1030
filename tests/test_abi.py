"""CPU: the C-ABI library loads and exports every symbol include/maf.h declares (no compute)."""
import os
import re

from maximizingacquisitionfunctions_b200 import _native as nat
from maximizingacquisitionfunctions_b200 import build as mbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "maf.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(maf_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    mbuild.build()
    lib = nat.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(nat.SIGNATURES) == syms
    assert lib.maf_version().decode().startswith("maf_b200")


def test_struct_layout_matches_header():
    import ctypes as C
    assert C.sizeof(nat.AcqParams) == 32
    assert nat.AcqParams.level.offset == 8 and nat.AcqParams.beta.offset == 24
