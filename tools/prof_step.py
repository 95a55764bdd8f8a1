"""Run the fused interface step a few times on C4 (for ncu captures; development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402

backend = ne_b200.TorchCudaBackend("cuda:0")
lib = ne_b200.get_library()
cfg = os.environ.get("NE_CFG", "C4")
ci = synthetic.build_case(cfg, backend, FT="f64", atm_FT="f32", with_iterations=False)
ci.initialize()
d = ci.fused_step_desc(0.37 * 10800.0)
for _ in range(4):
    lib.call("fused_interface_step", "f64", d, backend.stream())
torch.cuda.synchronize()
