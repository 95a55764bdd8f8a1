"""a-o kernel time on C4 for a non-default stability-function pair (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402

backend = ne_b200.TorchCudaBackend("cuda:0")
lib = ne_b200.get_library()
flux = ne_b200.SimilarityTheoryFluxes(stability_functions=ne_b200.large_yeager_stability_functions())
ci = synthetic.build_case("C4", backend, FT="f64", atm_FT="f32", atmosphere_ocean_fluxes=flux)
ci.initialize()
ci.interpolate_state(0.37 * 10800.0)
d = ci.atmosphere_ocean_desc()
for env in ({}, {"NE_B200_FORCE_GENERIC": "1"}):
    os.environ.update(env)
    for _ in range(2):
        lib.call("atmosphere_ocean_fluxes", "f64", d, backend.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        lib.call("atmosphere_ocean_fluxes", "f64", d, backend.stream())
    e1.record()
    torch.cuda.synchronize()
    print("Large-Yeager stability functions, C4:", env, f"{e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
    for k in env:
        os.environ.pop(k)
