"""The Julia extension cannot run here (no Julia in the image), so its ABI half is checked statically:

  * julia/ext/ne_b200_abi.jl is what tools/gen_julia_abi.py generates from include/ne_b200.h today (not stale);
  * every generated struct has the field names, order, OFFSETS and size of the ctypes mirror (abi.py), which the
    compiled library vouches for (tests/test_abi.py) — computed here with C's layout rules for isbits Julia structs;
  * every enum / #define constant has the value abi.py uses;
  * the hand-written extension (julia/ext/NumericalEarthB200Ext.jl) only uses struct names, field names and constants
    that exist, constructs every descriptor the nine overridden entry points need, and its delimiters balance.
"""
import ctypes as C
import os
import re
import subprocess
import sys

import ne_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ABI_JL = os.path.join(ROOT, "julia", "ext", "ne_b200_abi.jl")
EXT_JL = os.path.join(ROOT, "julia", "ext", "NumericalEarthB200Ext.jl")
A = ne_b200.abi

PRIM = {"Int32": (4, 4), "Int64": (8, 8), "UInt64": (8, 8), "Cdouble": (8, 8), "Ptr{Cvoid}": (8, 8)}


def _parse_julia_structs():
    text = open(ABI_JL).read()
    structs = {}
    for m in re.finditer(r"Base\.@kwdef struct (\w+)\n(.*?)\nend", text, flags=re.S):
        fields = []
        for line in m.group(2).splitlines():
            fm = re.match(r"\s+(\w+)::(.+?) = ", line)
            assert fm, line
            fields.append((fm.group(1), fm.group(2)))
        structs[m.group(1)] = fields
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"^const (NE_\w+) = (?:Int32\()?(-?\d+)\)?$", text, flags=re.M)}
    return structs, consts


def _size_align(t, structs, cache):
    """(size, alignment) of a Julia isbits type under C layout rules."""
    if t in PRIM:
        return PRIM[t]
    m = re.match(r"NTuple\{(\d+), (.+)\}$", t)
    if m:
        s, a = _size_align(m.group(2), structs, cache)
        return int(m.group(1)) * s, a
    if t not in cache:
        off, al = 0, 1
        for _, ft in structs[t]:
            s, a = _size_align(ft, structs, cache)
            off = (off + a - 1) // a * a + s
            al = max(al, a)
        cache[t] = ((off + al - 1) // al * al, al)
    return cache[t]


def _offsets(name, structs, cache):
    off, out = 0, []
    for f, ft in structs[name]:
        s, a = _size_align(ft, structs, cache)
        off = (off + a - 1) // a * a
        out.append((f, off, s))
        off += s
    return out


def test_generated_file_is_not_stale():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_julia_abi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_every_struct_matches_the_ctypes_mirror_field_by_field():
    structs, _ = _parse_julia_structs()
    assert set(structs) == set(A.STRUCTS), set(structs) ^ set(A.STRUCTS)
    cache = {}
    for name, cls in A.STRUCTS.items():
        jl = _offsets(name, structs, cache)
        py = [(f[0], getattr(cls, f[0]).offset, getattr(cls, f[0]).size) for f in cls._fields_]
        assert [x[0] for x in jl] == [x[0] for x in py], f"{name}: field names / order differ"
        assert jl == py, f"{name}: offsets or sizes differ:\n{jl}\n{py}"
        assert _size_align(name, structs, cache)[0] == C.sizeof(cls), name


def test_constants_match():
    _, consts = _parse_julia_structs()
    for k, v in consts.items():
        assert hasattr(A, k), f"{k} is in the header but not in abi.py"
        assert getattr(A, k) == v, k
    for k in dir(A):
        if k.startswith("NE_") and isinstance(getattr(A, k), int):
            assert k in consts, f"{k} is in abi.py but not in the header"


def _ext_text():
    text = open(EXT_JL).read()
    return re.sub(r"#=.*?=#", "", text, flags=re.S)


def test_extension_uses_only_existing_structs_fields_and_constants():
    structs, consts = _parse_julia_structs()
    text = _ext_text()
    code = "\n".join(line.split("#")[0] for line in text.splitlines())   # (no '#' inside strings in this file)
    code = re.sub(r'"(?:\\.|[^"\\])*"', '""', code)                       # names inside string literals (ENV keys) are not code
    for name in set(re.findall(r"\b(Ne[A-Z]\w+)\b", code)):
        assert name in structs, f"extension mentions unknown struct {name}"
    for c in set(re.findall(r"\b(NE_[A-Z0-9_]+)\b", code)):
        assert c in consts or c in ("NE_STRUCTS", "NE_B200_LIB"), f"extension mentions unknown constant {c}"
    # keyword constructions NeX(; a = …, b = …) / NeX(a = …): every keyword is a field of NeX
    for m in re.finditer(r"\b(Ne[A-Z]\w+)\(", code):
        name, i = m.group(1), m.end()
        depth, j = 1, i
        while depth and j < len(code):
            depth += code[j] in "([{"
            depth -= code[j] in ")]}"
            j += 1
        args = code[i:j - 1]
        top, d, cur = [], 0, ""
        for ch in args:
            d += ch in "([{"
            d -= ch in ")]}"
            if ch in ",;" and d == 0:
                top.append(cur); cur = ""
            else:
                cur += ch
        top.append(cur)
        fields = {f for f, _ in structs[name]}
        for a in top:
            km = re.match(r"\s*(\w+)\s*=[^=]", a)
            if km:
                assert km.group(1) in fields, f"{name}({km.group(1)} = …): no such field"


def test_extension_overrides_every_entry_point_and_builds_every_descriptor():
    text = _ext_text()
    for fn in ("initialize!", "interpolate_state!", "compute_atmosphere_ocean_fluxes!", "compute_atmosphere_sea_ice_fluxes!",
               "compute_sea_ice_ocean_fluxes!", "update_net_fluxes!", "apply_air_sea_radiative_fluxes!",
               "apply_air_sea_ice_radiative_fluxes!", "correct_state!"):
        assert re.search(r"function\s+[\w\.]*" + re.escape(fn) + r"\(", text), f"no method of {fn}"
    for desc in ("NeFracIndexDesc", "NeInterpDesc", "NeAtmosOceanDesc", "NeAtmosSeaIceDesc", "NeSeaIceOceanDesc", "NeSeaIceOceanStressDesc",
                 "NeAssembleOceanDesc", "NeAssembleSeaIceDesc", "NeApplyRadiationDesc", "NeElevationCorrectionDesc", "NeDiagDesc",
                 "NeFluxFormulation", "NeRoughnessLength", "NeInterfaceProperties", "NeSurfaceRadiation", "NeThermoParams",
                 "NeSubgridVelocity", "NeStopCriteria", "NeMediumProperties"):
        assert re.search(r"\b" + desc + r"\(", text), f"{desc} is never constructed"
    for sym in ("ne_frac_indices", "ne_interp_state", "ne_atmosphere_ocean_fluxes", "ne_atmosphere_sea_ice_fluxes", "ne_sea_ice_ocean_fluxes",
                "ne_sea_ice_ocean_stress", "ne_assemble_net_ocean_fluxes", "ne_assemble_net_sea_ice_fluxes", "ne_apply_radiative_fluxes",
                "ne_correct_atmosphere_elevation", "ne_diag_reduce", "ne_struct_size", "ne_last_error"):
        assert sym in text, f"{sym} is never called"
    # every symbol the extension ccalls exists in the header
    exported = set(A.all_entry_points())
    for s in set(re.findall(r":(ne_[a-z0-9_]+)", text)):
        assert s in exported or s + "_f64" in exported, f"extension calls {s}, which the header does not declare"


def test_extension_delimiters_balance():
    text = _ext_text()
    code = "\n".join(line.split("#")[0] for line in text.splitlines())
    code = re.sub(r'"(?:\\.|[^"\\])*"', '""', code)
    for o, c in ("()", "[]", "{}"):
        assert code.count(o) == code.count(c), f"unbalanced {o}{c}: {code.count(o)} vs {code.count(c)}"
    opens = len(re.findall(r"^\s*(?:function|if|for|while|let|module|struct|mutable struct|begin|try|quote|macro)\b", code, flags=re.M))
    opens += len(re.findall(r"\bdo\b", code))                               # `f(args) do x … end`
    opens += len(re.findall(r"=\s*(?:if|begin|let|try)\b", code))
    ends = len(re.findall(r"\bend\b", code))
    assert opens == ends, f"block openers {opens} vs `end` {ends}"


def _c_prototypes():
    """name -> (return type, [parameter types]) of every entry point declared in include/ne_b200.h, in a canonical spelling."""
    import re
    text = open(os.path.join(ROOT, "include", "ne_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"^\s*(int|int64_t|const char\*)\s+(ne_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text, flags=re.M):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"\s+", " ", a.strip())
                if re.match(r"const (Ne\w+)\s*\*", a):
                    params.append("struct:" + re.match(r"const (Ne\w+)\s*\*", a).group(1))
                elif a.startswith("const char*") or a.startswith("const char *"):
                    params.append("cstring")
                elif "*" in a:
                    params.append("pointer")
                elif a.startswith("int32_t"):
                    params.append("i32")
                elif a.startswith("int64_t"):
                    params.append("i64")
                elif a.startswith("uint64_t"):
                    params.append("u64")
                else:
                    raise AssertionError(f"unparsed parameter {a!r} of {name}")
        protos[name] = ({"int": "i32", "int64_t": "i64", "const char*": "cstring"}[ret], params)
    return protos


_JULIA_TYPES = {"Cint": "i32", "Int32": "i32", "Int64": "i64", "UInt64": "u64", "Cstring": "cstring", "Ptr{Cvoid}": "pointer",
                "Ptr{Float64}": "pointer", "Ptr{Ptr{Cvoid}}": "pointer"}


def test_every_ccall_matches_the_c_prototype_it_binds():
    """Return type, arity and argument kinds of every `ccall` in the extension against the prototype in include/ne_b200.h:
    `Ref{NeX}` must face `const NeX*` of the SAME struct, integers must have the declared width, both precisions of an
    `entry("ne_…", grid)` call must exist with the same signature."""
    import re
    src = open(os.path.join(ROOT, "julia", "ext", "NumericalEarthB200Ext.jl")).read()
    protos = _c_prototypes()
    assert len(protos) >= 45
    calls = re.findall(r'ccall\(\((:ne_[a-z0-9_]+|entry\("ne_[a-z0-9_]+", grid\)), (?:path|libne\[\])\), (\w+),\s*\(([^)]*)\)', src)
    assert len(calls) == len(re.findall(r"\bccall\(", src)) and len(calls) >= 19, len(calls)     # every ccall was parsed
    seen = set()
    for target, ret, argt in calls:
        if target.startswith(":"):
            names = [target[1:]]
        else:
            stem = re.match(r'entry\("(ne_[a-z0-9_]+)"', target).group(1)
            names = [stem + "_f64", stem + "_f32"]
        jargs = [a.strip() for a in argt.split(",") if a.strip()]
        for name in names:
            assert name in protos, f"ccall of {name}: no such prototype in the header"
            cret, cparams = protos[name]
            assert _JULIA_TYPES[ret] == cret, f"{name}: Julia return type {ret} vs C {cret}"
            assert len(jargs) == len(cparams), f"{name}: {len(jargs)} ccall arguments vs {len(cparams)} C parameters"
            for ja, cp in zip(jargs, cparams):
                m = re.match(r"Ref\{(Ne\w+)\}", ja)
                if m:
                    assert cp == "struct:" + m.group(1), f"{name}: Julia passes {ja}, C expects {cp}"
                else:
                    assert ja in _JULIA_TYPES, f"{name}: unknown Julia argument type {ja}"
                    assert _JULIA_TYPES[ja] == cp, f"{name}: Julia passes {ja}, C expects {cp}"
            seen.add(name)
    # the path's entry points are all bound
    for stem in ("ne_frac_indices", "ne_interp_state", "ne_atmosphere_ocean_fluxes", "ne_atmosphere_sea_ice_fluxes", "ne_sea_ice_ocean_fluxes",
                 "ne_sea_ice_ocean_stress", "ne_assemble_net_ocean_fluxes", "ne_assemble_net_sea_ice_fluxes", "ne_apply_radiative_fluxes",
                 "ne_correct_atmosphere_elevation", "ne_diag_reduce"):
        assert stem + "_f64" in seen and stem + "_f32" in seen, stem
