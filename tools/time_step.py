"""Time the fused interface step and its pieces on a workload for a list of env-var settings
(development tool; numbers go to gpurun_out/, not to bench lines)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import synthetic  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    cfg = os.environ.get("NE_CFG", "C4")
    settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in sys.argv[1:]] or [{}]
    backend = ne_b200.TorchCudaBackend("cuda:0")
    lib = ne_b200.get_library()
    ci = synthetic.build_case(cfg, backend, FT="f64", atm_FT="f32", with_iterations=False)
    ci.initialize()
    from numericalearth_jl_b200 import sharding
    f = ci.ao_fluxes
    diag = sharding.FluxDiagnostics(ci, [f.latent_heat, f.sensible_heat, f.water_vapor, f.x_momentum, f.y_momentum,
                                         ci.net_ocean.T, ci.net_ocean.eta], n_blocks=int(os.environ.get("NE_DIAG_BLOCKS", "1184")))
    d = ci.fused_step_desc(0.37 * 10800.0, diagnostics=diag)
    stream = backend.stream()
    out = []
    for env in settings:
        for k, v in env.items():
            os.environ[k] = v
        r = {"env": env}
        r["fused_step"] = timed(lambda: lib.call("fused_interface_step", "f64", d, stream))
        r["interp_and_ao"] = timed(lambda: lib.call("interp_and_ao_fluxes", "f64", d, stream))
        r["interp_atm"] = timed(lambda: lib.call("interp_state", "f64", d.atmosphere, stream))
        r["interp_rad"] = timed(lambda: lib.call("interp_state", "f64", d.radiation, stream))
        r["ao"] = timed(lambda: lib.call("atmosphere_ocean_fluxes", "f64", d.ao, stream))
        r["assemble"] = timed(lambda: lib.call("assemble_net_ocean_fluxes", "f64", d.assemble, stream))
        r["apply_rad"] = timed(lambda: lib.call("apply_radiative_fluxes", "f64", d.apply_radiation, stream))
        r["diag"] = timed(lambda: lib.call("diag_reduce", "f64", diag.desc, stream))
        r["post_solve(fused_step - interp_and_ao)"] = r["fused_step"] - r["interp_and_ao"]
        out.append(r)
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}))
        for k in env:
            os.environ.pop(k, None)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/time_step.json", "w"))


if __name__ == "__main__":
    main()
