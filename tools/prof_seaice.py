"""Run the atmosphere-sea-ice kernel a few times on C3 (for ncu captures; development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402

backend = ne_b200.TorchCudaBackend("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
ci = synthetic.build_case(cfg, backend, FT="f64", atm_FT="f32", sea_ice=True, with_iterations=True)
ci.initialize()
ci.interpolate_state(0.37 * 10800.0)
for _ in range(3):
    ci.compute_atmosphere_sea_ice_fluxes()
torch.cuda.synchronize()
