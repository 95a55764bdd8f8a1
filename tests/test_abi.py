"""The C-ABI library loads without a GPU, exports every symbol include/ne_b200.h declares, and its
struct layouts match the ctypes mirror.  No compute calls here."""
import ctypes
import os
import re

import ne_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "ne_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|int64_t|const char\*)\s+(ne_[a-z0-9_]+)\s*\(", src, flags=re.M)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = ne_b200.Library()        # raises if the extension has not been built: no CPU fallback
    names = _header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib.dll, n), f"{n} declared in include/ne_b200.h but not exported"
    assert sorted(names) == sorted(ne_b200.abi.all_entry_points())
    assert lib.dll.ne_version() == ne_b200.abi.NE_ABI_VERSION


def test_struct_layouts_match_ctypes_mirror():
    lib = ne_b200.Library()
    for name, cls in ne_b200.abi.STRUCTS.items():
        assert lib.dll.ne_struct_size(name.encode()) == ctypes.sizeof(cls), name
    assert lib.dll.ne_struct_size(b"NoSuchStruct") == -1


def test_header_structs_all_mirrored():
    src = open(os.path.join(ROOT, "include", "ne_b200.h")).read()
    declared = set(re.findall(r"typedef struct (Ne[A-Za-z0-9]+)", src))
    assert declared == set(ne_b200.abi.STRUCTS)


def test_invalid_descriptor_is_rejected_without_touching_a_device():
    """Validation happens before any launch: a null descriptor / bad grid returns NE_E_INVALID."""
    lib = ne_b200.Library()
    d = ne_b200.abi.NeAtmosOceanDesc()   # all zero: nx = 0
    try:
        lib.call("atmosphere_ocean_fluxes", "f64", d, 0)
    except ne_b200.NeError as e:
        assert e.code == ne_b200.abi.NE_E_INVALID
    else:
        raise AssertionError("zeroed descriptor was accepted")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "numericalearth.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "oracle/" not in text and "ne_oracle" not in text, f
