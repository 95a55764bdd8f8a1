"""numericalearth.jl_b200 — B200-native atmosphere–surface interface path of NumericalEarth.jl.

The directory name carries a dot, so import it through the root shim:

    import ne_b200                       # registers this package as `numericalearth_jl_b200`
    from ne_b200 import ComponentInterfaces, SimilarityTheoryFluxes, ...

Only the hot path lives here: `csrc/` (hand-written sm_100a kernels + the C ABI of
include/ne_b200.h), `abi.py` (ctypes mirror), `formulations.py` (the reference's plugin types ->
POD kernel variants), `interface.py` (the reference's interface-computation API), `sharding.py`
(latitude bands over torch.distributed), `pipeline.py` (host-buffer step with copy/compute overlap) and `synthetic.py` (JRA55/ECCO-shaped synthetic inputs).
"""
from . import abi  # noqa: F401
from .formulations import *  # noqa: F401,F403
from .formulations import NoKernelVariantError  # noqa: F401
from .interface import (ComponentInterfaces, ExchangeGrid, LatLonSourceGrid, PrescribedAtmosphere,  # noqa: F401
                        PrescribedLand, PrescribedRadiation, SlabLandState, interpolating_time_indices)
from .pipeline import HostPipelinedStep  # noqa: F401
from .series_window import SeriesWindow, WindowPolicy  # noqa: F401
from .lib import LIB_PATH, Library, NeError, NumpyHostBackend, TorchCudaBackend, get_library  # noqa: F401
