"""Time the interpolation launches of the C4 step (merged 9-series launch and the two component launches) for a list of
env-var settings, and check every setting bit for bit against the first one (development tool)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402


def main():
    cfg = sys.argv[1]
    settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in sys.argv[2:]] or [{}]
    backend = ne_b200.TorchCudaBackend("cuda:0")
    lib = ne_b200.get_library()
    ci = synthetic.build_case(cfg, backend, FT="f64", atm_FT="f32", with_iterations=True)
    ci.initialize()
    fused = ci.fused_step_desc(0.37 * 10800.0)
    stream = backend.stream()
    ref = None
    out = []
    for env in settings:
        os.environ.update(env)
        res = {}
        def step():
            lib.call("fused_interface_step", "f64", fused, stream)
        def atm():
            lib.call("interp_state", "f64", fused.atmosphere, stream)
        def rad():
            lib.call("interp_state", "f64", fused.radiation, stream)
        for name, fn in (("atmosphere(7)", atm), ("radiation(2)", rad), ("fused_step(merged 9)", step)):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 20
        # the step's own interpolation phase (merged launch when the two descriptors allow it)
        for _ in range(3):
            lib.call("fused_interface_step", "f64", fused, stream)
        torch.cuda.synchronize()
        fields = {n: backend.to_numpy(getattr(ci.atmos_state, n)).copy() for n in ci.atmos_state.names()}
        fields.update({"rad_" + n: backend.to_numpy(getattr(ci.rad_state, n)).copy() for n in ci.rad_state.names()})
        same = None
        if ref is None:
            ref = fields
        else:
            same = all(np.array_equal(ref[k], fields[k], equal_nan=True) for k in ref)
        print(cfg, env, {k: round(v, 4) for k, v in res.items()}, "bitwise equal to first:", same, flush=True)
        out.append({"env": env, "ms": res, "bitwise_equal_to_first": same})
        for k in env:
            os.environ.pop(k)
    json.dump(out, open(f"gpurun_out/time_interp_{cfg}.json", "w"))


if __name__ == "__main__":
    main()
