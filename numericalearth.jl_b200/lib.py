"""Loader for libne_b200.so (the C-ABI library of sm_100a kernels) and array back-ends.

The product path has exactly one compute library: the CUDA extension.  If it has not been built
(`python __graft_entry__.py build`) or no CUDA device is visible, every compute call raises — there
is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import abi as A

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libne_b200.so")


class NeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libne_b200 error {code}: {message}")
        self.code = code


class Library:
    """Typed ctypes view of a shared library exporting the include/ne_b200.h entry points.

    `prefix` is "ne_" for the product library.  (The test-suite loads the CPU oracle, which exports
    the same descriptors under "neo_", through this same class — from tests/ only.)
    """

    def __init__(self, path=LIB_PATH, prefix="ne_", takes_stream=True, is_device=True):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no CPU fallback.")
        self.path = path
        self.prefix = prefix
        self.takes_stream = takes_stream
        self.is_device = is_device
        self.dll = C.CDLL(path)
        for base, desc in A.DESC_ENTRY_POINTS.items():
            for suffix in ("_f64", "_f32"):
                name = prefix + base[len("ne_"):] + suffix
                fn = getattr(self.dll, name, None)
                if fn is None:
                    continue
                fn.restype = C.c_int
                fn.argtypes = [C.POINTER(desc), C.c_void_p] if takes_stream else [C.POINTER(desc)]
        if prefix == "ne_":
            self.dll.ne_last_error.restype = C.c_char_p
            self.dll.ne_version.restype = C.c_int
            self.dll.ne_device_count.restype = C.c_int
            self.dll.ne_struct_size.restype = C.c_int64
            self.dll.ne_struct_size.argtypes = [C.c_char_p]
            self.dll.ne_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
            self.dll.ne_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
            self.dll.ne_stream_synchronize.argtypes = [C.c_void_p]
            self.dll.ne_measure_fp64_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
            self.dll.ne_count_solve_ops_f64.argtypes = [C.POINTER(A.NeAtmosOceanDesc), C.POINTER(C.c_uint64), C.c_void_p]
            self.dll.ne_host_pipeline_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
            self.dll.ne_host_pipeline_destroy.argtypes = [C.c_void_p]
            self.dll.ne_series_ring_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(A.NeSeriesRingDesc)]
            self.dll.ne_series_ring_destroy.argtypes = [C.c_void_p]
            self.dll.ne_series_ring_load.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
            self.dll.ne_series_ring_acquire.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
            self.dll.ne_series_ring_release.argtypes = [C.c_void_p, C.c_void_p]
            for suffix in ("_f64", "_f32"):
                fn = getattr(self.dll, "ne_host_pipelined_step" + suffix)
                fn.restype = C.c_int
                fn.argtypes = [C.c_void_p, C.POINTER(A.NeHostStepDesc), C.c_void_p]

    def last_error(self):
        if self.prefix != "ne_":
            return ""
        return (self.dll.ne_last_error() or b"").decode()

    def call(self, base, dtype, desc, stream=0):
        """base: e.g. 'atmosphere_ocean_fluxes'; dtype: 'f64'|'f32'."""
        fn = getattr(self.dll, f"{self.prefix}{base}_{dtype}")
        rc = fn(C.byref(desc), C.c_void_p(stream)) if self.takes_stream else fn(C.byref(desc))
        if rc != 0:
            msg = self.last_error()
            if rc == A.NE_E_NO_VARIANT:
                from .formulations import NoKernelVariantError
                raise NoKernelVariantError(msg)
            raise NeError(rc, msg)

    def device_count(self):
        return self.dll.ne_device_count()

    def diag_allreduce(self, nccl_comm, sums_ptr, n, stream=0):
        """ne_diag_allreduce_f64: in-place NCCL sum of `n` device doubles over the ranks of the caller's ncclComm_t
        (an integer handle), enqueued on `stream`."""
        self.dll.ne_diag_allreduce_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        rc = self.dll.ne_diag_allreduce_f64(C.c_void_p(nccl_comm), C.c_void_p(sums_ptr), C.c_int32(n), C.c_void_p(stream))
        if rc != 0:
            if rc == A.NE_E_NO_VARIANT:
                from .formulations import NoKernelVariantError
                raise NoKernelVariantError(self.last_error())
            raise NeError(rc, self.last_error())

    def count_solve_ops(self, desc, stream=0):
        """FP64 instructions one launch of the a–o solve executes on `desc` (ne_count_solve_ops_f64): dict of thread-level
        counts.  Synchronises."""
        out = (C.c_uint64 * 8)()
        rc = self.dll.ne_count_solve_ops_f64(C.byref(desc), out, C.c_void_p(stream))
        if rc != 0:
            if rc == A.NE_E_NO_VARIANT:
                from .formulations import NoKernelVariantError
                raise NoKernelVariantError(self.last_error())
            raise NeError(rc, self.last_error())
        fma, mul, add, other, trips, wtrips = (int(out[k]) for k in range(6))
        return {"dfma": fma, "dmul": mul, "dadd": add, "library_code_estimate": other, "thread_trips": trips,
                "warp_trips": wtrips, "flop": 2 * fma + mul + add}

    def measure_fp64_peak(self):
        tf, mhz = C.c_double(0), C.c_double(0)
        rc = self.dll.ne_measure_fp64_peak(C.byref(tf), C.byref(mhz))
        if rc != 0:
            raise NeError(rc, self.last_error())
        return tf.value, mhz.value


_default = None


def get_library():
    """The process-wide CUDA library; raises if it is missing."""
    global _default
    if _default is None:
        _default = Library()
    return _default


# -------------------------------------------------------------------------------------------------
# array back-ends: how fields are allocated and how their addresses are taken
# -------------------------------------------------------------------------------------------------
NP_DTYPES = {"f64": np.float64, "f32": np.float32}


class TorchCudaBackend:
    """Device arrays are torch CUDA tensors (PyTorch is the memory/stream plumbing, not the product)."""
    is_device = True

    def __init__(self, device="cuda:0"):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible: numericalearth.jl_b200 has no CPU fallback")
        self.device = torch.device(device)
        self._td = {"f64": torch.float64, "f32": torch.float32, "u8": torch.uint8, "i32": torch.int32}

    def zeros(self, shape, dtype):
        return self.torch.zeros(shape, dtype=self._td[dtype], device=self.device)

    def from_numpy(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def to_numpy(self, a):
        return a.detach().cpu().numpy()

    def ptr(self, a):
        return a.data_ptr()

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)


class NumpyHostBackend:
    """Host arrays.  Only meaningful together with a host library (the oracle, in tests)."""
    is_device = False
    _td = {"f64": np.float64, "f32": np.float32, "u8": np.uint8, "i32": np.int32}

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype=self._td[dtype])

    def from_numpy(self, a):
        return np.ascontiguousarray(a).copy()

    def to_numpy(self, a):
        return a

    def ptr(self, a):
        return a.ctypes.data

    def stream(self):
        return 0

    def synchronize(self):
        pass
