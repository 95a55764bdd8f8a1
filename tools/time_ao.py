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
    # usage: time_ao.py [--config C4] [--dtype f64] [--out name] ENV1=a,ENV2=b ENV1=c ...
    argv = sys.argv[1:]
    opts = {"--config": "C4", "--dtype": "f64", "--out": "time_ao"}
    while argv and argv[0] in opts:
        opts[argv[0]] = argv[1]
        argv = argv[2:]
    settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in argv] or [{}]
    backend = ne_b200.TorchCudaBackend("cuda:0")
    lib = ne_b200.get_library()
    FT = opts["--dtype"]
    ci = synthetic.build_case(opts["--config"], backend, FT=FT, atm_FT="f32", with_iterations=True)   # iterations: the trip-count ordered launch
    ci.initialize()
    ci.interpolate_state(0.37 * 10800.0)
    d = ci.atmosphere_ocean_desc()
    stream = backend.stream()
    out = []
    for env in settings:
        for k, v in env.items():
            os.environ[k] = v
        for _ in range(3):
            lib.call("atmosphere_ocean_fluxes", FT, d, stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            lib.call("atmosphere_ocean_fluxes", FT, d, stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        rec = {"env": env, "ms": ms, "config": opts["--config"], "dtype": FT}
        if FT == "f64" and "NE_B200_TAB_V1" not in env and "NE_B200_TAB2_LIBM_PROLOGUE" not in env:
            try:
                rec["counts"] = lib.count_solve_ops(d, stream)
                rec["executed_tflops"] = rec["counts"]["flop"] / (ms * 1e-3) / 1e12
            except Exception as e:   # noqa: BLE001
                rec["counts_error"] = repr(e)[:200]
        out.append(rec)
        print(opts["--config"], FT, env, f"{ms:.3f} ms", rec.get("executed_tflops"), flush=True)
        for k in env:
            os.environ.pop(k, None)
    json.dump(out, open(f"gpurun_out/{opts['--out']}.json", "w"))


if __name__ == "__main__":
    main()
