"""Time the atmosphere-ocean kernel alone on the C4 workload for a list of env-var settings
(development tool; numbers go to gpurun_out/, not to bench lines)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402


def main():
    settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in sys.argv[1:]] or [{}]
    backend = ne_b200.TorchCudaBackend("cuda:0")
    lib = ne_b200.get_library()
    ci = synthetic.build_case("C4", backend, FT="f64", atm_FT="f32", with_iterations=False)
    ci.initialize()
    ci.interpolate_state(0.37 * 10800.0)
    d = ci.atmosphere_ocean_desc()
    stream = backend.stream()
    out = []
    for env in settings:
        for k, v in env.items():
            os.environ[k] = v
        for _ in range(3):
            lib.call("atmosphere_ocean_fluxes", "f64", d, stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            lib.call("atmosphere_ocean_fluxes", "f64", d, stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out.append({"env": env, "ms": ms})
        print(env, f"{ms:.3f} ms")
        for k in env:
            os.environ.pop(k, None)
    json.dump(out, open("gpurun_out/time_ao.json", "w"))


if __name__ == "__main__":
    main()
