"""The C-ABI library loads on a CPU-only box and exports every symbol include/oibvh_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "oibvh_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oibvh_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    import oibvh_b200 as ob
    lib = ctypes.CDLL(ob.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/oibvh_b200.h but not exported"


def test_python_binding_covers_header():
    import oibvh_b200 as ob
    assert sorted(ob._SIGNATURES) == declared_symbols()


def test_record_layouts():
    text = open(os.path.join(ROOT, "include", "oibvh_b200.h")).read()
    assert "float min[3]" in text and "float max[3]" in text
    assert "uint32_t bvh_index[2]" in text and "uint32_t tri_index[2]" in text


def test_no_cpu_fallback():
    """without a GPU every compute entry point must fail loudly instead of computing on the host"""
    import oibvh_b200 as ob
    if ob.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(ob.OibvhError) as e:
        ob.Context(0)
    assert e.value.code == -2


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under oibvh_b200/ or include/ may reference it"""
    for base in ("oibvh_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    src = open(os.path.join(dirpath, f), errors="replace").read()
                    assert "import oracle" not in src and "liboibvh_oracle" not in src and "oibvh_ref" not in src, \
                        f"{os.path.join(dirpath, f)} references the oracle"
