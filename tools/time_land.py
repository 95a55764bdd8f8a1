"""Atmosphere-land kernel time on a grid for the default land tree, fast path vs generic (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import formulations as F  # noqa: E402
import test_land_fluxes as tl  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
backend = ne_b200.TorchCudaBackend("cuda:0")
tl.CFG = ne_b200.synthetic.CONFIGS[cfg] if hasattr(ne_b200, "synthetic") else None
from numericalearth_jl_b200 import synthetic  # noqa: E402
tl.CFG = synthetic.CONFIGS[cfg]
for name in ("bulk", "skin"):
    ci = tl._case(backend, None, "f64", "f32", tl.HUMIDITIES[name]())
    for env in ({}, {"NE_B200_FORCE_GENERIC": "1"}):
        os.environ.update(env)
        for _ in range(2):
            ci.compute_atmosphere_land_fluxes()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            ci.compute_atmosphere_land_fluxes()
        e1.record()
        torch.cuda.synchronize()
        it = backend.to_numpy(ci.al_iterations)
        print(cfg, name, env, f"{e0.elapsed_time(e1) / 5:.3f} ms", "mean trips", float(it[it > 0].mean()), "max", int(it.max()), flush=True)
        for k in env:
            os.environ.pop(k)
