"""CPU oracle — TEST INFRASTRUCTURE ONLY (see the header of oracle/ne_oracle.cpp).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, and by
nothing in the product package.  Parity status: unpinned at the last-ulp level (Julia is absent, so
the reference cannot be executed here); pinned by the reference's own known-answer tests
(tests/test_oracle_reference_kats.py).
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ne_oracle.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libne_oracle.so")
FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math"]


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    hdr = os.path.join(HERE, "..", "include", "ne_b200.h")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        return LIB
    cmd = ["g++"] + FLAGS + ["-o", LIB, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stderr[-4000:])
    return LIB


_lib = None


def load():
    """Return the oracle as a `Library` (product ABI loader class) bound to the neo_* symbols."""
    global _lib
    if _lib is None:
        import ne_b200
        path = build()
        _lib = ne_b200.Library(path, prefix="neo_", takes_stream=False, is_device=False)
        d = _lib.dll
        d.neo_land_interface_humidity.restype = C.c_double
        d.neo_land_interface_humidity.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 9
        d.neo_saturation_specific_humidity.restype = C.c_double
        d.neo_saturation_specific_humidity.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
        d.neo_stability_f64.restype = C.c_double
        d.neo_stability_f64.argtypes = [C.c_void_p, C.c_double]
        d.neo_vsgs2_f64.restype = C.c_double
        d.neo_vsgs2_f64.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        d.neo_polynomial_drag_f64.restype = C.c_double
        d.neo_polynomial_drag_f64.argtypes = [C.c_void_p, C.c_double]
        d.neo_momentum_roughness_f64.restype = C.c_double
        d.neo_momentum_roughness_f64.argtypes = [C.c_void_p, C.c_double, C.c_double]
        d.neo_scalar_roughness_f64.restype = C.c_double
        d.neo_scalar_roughness_f64.argtypes = [C.c_void_p, C.c_double, C.c_double]
        d.neo_saturation_vapor_pressure_f64.restype = C.c_double
        d.neo_saturation_vapor_pressure_f64.argtypes = [C.c_void_p, C.c_double, C.c_int]
        d.neo_surface_specific_humidity_f64.restype = C.c_double
        d.neo_surface_specific_humidity_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double]
        d.neo_interpolator_f64.argtypes = [C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        d.neo_interpolator_f32.argtypes = [C.c_float, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_float)]
        d.neo_set_threads.argtypes = [C.c_int]
        d.neo_series_slot_fill.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        d.neo_max_threads.restype = C.c_int
        d.neo_get_op_counts.argtypes = [C.POINTER(C.c_uint64)]
    return _lib
