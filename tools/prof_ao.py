"""Run the atmosphere-ocean kernel a few times on C4 (for ncu captures; development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402

backend = ne_b200.TorchCudaBackend("cuda:0")
lib = ne_b200.get_library()
cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
FT = sys.argv[2] if len(sys.argv) > 2 else "f64"
ci = synthetic.build_case(cfg, backend, FT=FT, atm_FT="f32", with_iterations=True)
ci.initialize()
ci.interpolate_state(0.37 * 10800.0)
d = ci.atmosphere_ocean_desc()
for _ in range(4):
    lib.call("atmosphere_ocean_fluxes", FT, d, backend.stream())
torch.cuda.synchronize()
