"""Fixed cost of one a-o solve call on tiny grids (development tool): where do the ~60 us of a small launch go?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402

backend = ne_b200.TorchCudaBackend("cuda:0")
lib = ne_b200.get_library()
stream = backend.stream()
for nx, ny in ((30, 6), (62, 14), (360, 150), (1440, 560)):
    cfg = dict(nx=nx, ny=ny, latitude=(-40.0, 40.0), src_nx=64, src_ny=32)
    ci = synthetic.build_case(cfg, backend, FT="f64", atm_FT="f32", with_iterations=True)
    ci.initialize()
    ci.interpolate_state(0.37 * 10800.0)
    d = ci.atmosphere_ocean_desc()
    for env in ({}, {"NE_B200_TAB2_NO_ORDER": "1"}, {"NE_B200_TAB_V1": "1"}):
        os.environ.update(env)
        for _ in range(5):
            lib.call("atmosphere_ocean_fluxes", "f64", d, stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            lib.call("atmosphere_ocean_fluxes", "f64", d, stream)
        e1.record()
        torch.cuda.synchronize()
        it = backend.to_numpy(ci.grid.interior(ci.ao_iterations))
        print(f"{nx}x{ny} {env} {e0.elapsed_time(e1) / 50 * 1000:.1f} us per call; max trips {it.max()} mean {it[it > 0].mean():.1f}", flush=True)
        for k in env:
            os.environ.pop(k)
